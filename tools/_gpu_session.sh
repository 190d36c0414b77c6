# round 2 session D: whole GPU suite after the hygiene changes; ncu --set full of the IBM=true collide instantiation; default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02k_pytest_gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:collide_push_kernel<\(int\)1, \(bool\)1' -s 10 -c 2 -o gpurun_out/r02k_ncu_collide_ibm_true python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02k_under_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/r02k_bench_plate512_s400.json 2> gpurun_out/err_k1.txt; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02k_bench_plate512_s400.json
ls -la gpurun_out | tail -5
