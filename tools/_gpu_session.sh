# round 2 session S: GPU side after the structural fix: reference goldens (both routes), FSI and harness tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_golden.py tests/test_gpu_fsi.py tests/test_gpu_harness.py tests/test_gpu_io.py -m gpu -q > gpurun_out/r02aa_pytest.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/r02aa_pytest.txt | cut -c1-300
