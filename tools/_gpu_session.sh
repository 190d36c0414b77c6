# round 2 session P: whole GPU suite + smoke + default bench after the last library changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02y_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02y_bench_plate512_s20.json 2>/dev/null; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r02y_bench_plate512_s20.json')); r=d['roofline']
print(d['config']['workload'], round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), r['traffic'], round(d['e2e']['value']), d['parity_check']['ok'], d['cpu_baseline']['value'], d['gpu_launches'])
P
