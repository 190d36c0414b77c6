# round 2 session L: CUDA path against all reference goldens (incl. plate in son, output files)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_golden.py -m gpu -q > gpurun_out/r02s_pytest_reference_golden.txt 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02s_pytest_reference_golden.txt | cut -c1-300
