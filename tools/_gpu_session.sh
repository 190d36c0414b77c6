# round 2 session R: compute-sanitizer memcheck over a representative subset (small cases)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r02_memcheck.log python -m pytest tests/test_gpu_golden.py tests/test_gpu_refine.py tests/test_gpu_ibm_exact.py tests/test_gpu_io.py -m gpu -x -q > gpurun_out/r02_memcheck_pytest.txt 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r02_memcheck_pytest.txt; grep -c "Invalid\|Error" gpurun_out/r02_memcheck.log; tail -5 gpurun_out/r02_memcheck.log | cut -c1-300
