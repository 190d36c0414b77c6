# round 2 session AJ: cell-list sort by 32 cells per warp; son-to-father faces in one launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ibm_exact.py tests/test_gpu_fsi.py tests/test_gpu_refine.py tests/test_gpu_reference_golden.py tests/test_gpu_harness.py -m gpu -q -x > gpurun_out/r02aj_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02aj_pytest.txt | cut -c1-300
timeout 400 python bench.py --workload school2048r --steps 100 --warmup 10 --no-cpu-baseline --no-parity-check > gpurun_out/r02aj_school2048r.json 2> gpurun_out/err_aj1.txt; echo "rc=$?"
timeout 400 python bench.py --workload school2048r --steps 100 --warmup 10 --no-cpu-baseline --no-parity-check --trace-out gpurun_out/r02aj_trace_school2048r > gpurun_out/r02aj_school2048r_tr.json 2> gpurun_out/err_aj2.txt; echo "rc=$?"
timeout 300 python bench.py --workload heave1024 --steps 100 --warmup 10 --no-cpu-baseline --no-parity-check > gpurun_out/r02aj_heave1024.json 2> gpurun_out/err_aj3.txt; echo "rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02aj_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), round(r['collide_alone']['kernel_ms'],4), d['clocks']['sm_mhz'], round(d['e2e']['value']), d['e2e']['seconds'], d['details']['structural_solver']['host_ms_per_step_all_bodies'])
    except Exception as e: print(f, 'ERR', e)
P
head -30 gpurun_out/r02aj_trace_school2048r.rank0.csv
