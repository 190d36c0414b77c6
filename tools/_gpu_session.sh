mkdir -p gpurun_out
timeout 400 python bench.py --workload school2048r --steps 40 --warmup 10 --no-cpu-baseline --no-parity-check --trace-out gpurun_out/r02w_trace_school2048r > gpurun_out/r02w_bench.json 2> gpurun_out/err_w1.txt; echo "rc=$?"
FSILBM_IBM_PROFILE=1 timeout 400 python bench.py --workload school2048r --steps 120 --warmup 10 --no-cpu-baseline --no-parity-check > /dev/null 2> gpurun_out/r02w_ibm_profile.txt; echo "rc=$?"; grep "ibm" gpurun_out/r02w_ibm_profile.txt | head -8 | cut -c1-300
