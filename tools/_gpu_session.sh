# round 2 session O: hoisted son IBM -- refinement tests, plate-in-son golden, school2048r on one GPU with and without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_refine.py tests/test_gpu_reference_golden.py -m gpu -q -k "refine or son or plate" > gpurun_out/r02v_pytest.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/r02v_pytest.txt | cut -c1-300
timeout 400 python bench.py --workload school2048r --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02v_bench_school2048r_n1_hoist.json 2> gpurun_out/err_v1.txt; echo "rc=$?"
FSILBM_NO_HOIST=1 timeout 400 python bench.py --workload school2048r --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02v_bench_school2048r_n1_nohoist.json 2> gpurun_out/err_v2.txt; echo "rc=$?"
python - <<'P'
import json
for t in ('hoist','nohoist'):
    d=json.load(open(f'gpurun_out/r02v_bench_school2048r_n1_{t}.json')); r=d['roofline']
    print(t, round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), d['details']['structural_solver']['host_ms_per_step_all_bodies'], d['clocks']['sm_mhz'])
P
