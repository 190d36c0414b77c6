# round 2 session J: whole GPU suite after the IBM arm removal and the cross-slab son work
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02q_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02q_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
