# round 2 session B: full-size parity tests; ncu --set full of the IBM=true collide instantiation
mkdir -p gpurun_out
nproc; free -g | head -2
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q --durations=5 > gpurun_out/r02i_pytest_fullsize.txt 2>&1; echo "fullsize rc=$?"; tail -12 gpurun_out/r02i_pytest_fullsize.txt
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:collide_push_kernel<1, 1' -s 10 -c 2 -o gpurun_out/r02i_ncu_collide_ibm_true python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02i_under_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out | tail -5
