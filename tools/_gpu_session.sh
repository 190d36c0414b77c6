# round 2 session C: CUDA path against the reference goldens; ncu --set full of the IBM=true collide instantiation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_golden.py -m gpu -q > gpurun_out/r02j_pytest_reference_golden.txt 2>&1; echo "refgolden rc=$?"; tail -40 gpurun_out/r02j_pytest_reference_golden.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:collide_push_kernel<1, true' -s 10 -c 2 -o gpurun_out/r02j_ncu_collide_ibm_true python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02j_under_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -5
