# round 2 session AB: whole GPU suite at HEAD, smoke, the driver's two bench commands at N=1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02ab_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02ab_pytest.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02ab_ref.json 2> gpurun_out/err_ab_ref.txt; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02ab_bench.json 2> gpurun_out/err_ab_bench.txt; echo "bench rc=$?"
cut -c1-600 gpurun_out/r02ab_ref.json; cut -c1-1500 gpurun_out/r02ab_bench.json
