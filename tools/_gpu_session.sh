# round 2 session AP (8 GPUs): the driver's N=8 commands, 200 steps, school2048r
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29551 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02ap_ref_n8.json 2> gpurun_out/err_ap0.txt; echo "ref rc=$?"
timeout 600 $T --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02ap_bench_n8_s20.json 2> gpurun_out/err_ap1.txt; echo "bench rc=$?"
timeout 600 $T --master-port 29553 bench.py --gpus 8 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r02ap_bench_n8_s200.json 2> gpurun_out/err_ap2.txt; echo "bench rc=$?"
timeout 600 $T --master-port 29554 bench.py --gpus 8 --workload school2048r --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r02ap_school2048r_n8.json 2> gpurun_out/err_ap3.txt; echo "bench rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ap_*.json')):
    try:
        d=json.load(open(f)); r=d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],4), r.get('frac'), d.get('clocks',{}) and d['clocks'].get('sm_mhz'), round(d['e2e']['value']), (d.get('parity_check') or {}).get('ok'))
    except Exception as e: print(f, 'ERR', e)
P
