# round 2 session AO: whole GPU suite at HEAD, smoke, the driver's two N=1 commands, launch lists of the default bench and of school2048r
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02ao_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02ao_pytest.txt | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02ao_ref.json 2> gpurun_out/err_ao_ref.txt; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02ao_bench.json 2> gpurun_out/err_ao_bench.txt; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 220 --csv --log-file gpurun_out/r02ao_launches_school2048r.csv python bench.py --workload school2048r --steps 4 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02ao_under_ncu1.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r02ao_launches_plate512.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02ao_under_ncu2.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import json
for f in ('gpurun_out/r02ao_ref.json','gpurun_out/r02ao_bench.json'):
    d=json.load(open(f)); r=d.get('roofline') or {}
    print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],4), r.get('frac'), d.get('clocks'), d['e2e']['value'], (d.get('parity_check') or {}).get('ok'), d.get('gpu_launches'))
P
