# round 2 session AK (2 GPUs): multi-rank parity after the IBM / transfer kernel changes; the driver's N=2 commands
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multi_rank_case.py > gpurun_out/r02ak_multi_rank_parity_n2.txt 2>&1; echo "parity rc=$?"; grep -c "OK" gpurun_out/r02ak_multi_rank_parity_n2.txt; grep -v "OK$" gpurun_out/r02ak_multi_rank_parity_n2.txt | tail -8 | cut -c1-250
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02ak_ref_n2.json 2> gpurun_out/err_ak0.txt; echo "ref rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02ak_bench_n2_s20.json 2> gpurun_out/err_ak1.txt; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r02ak_bench_n2_s200.json 2> gpurun_out/err_ak2.txt; echo "bench rc=$?"
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ak_*.json')):
    try:
        d=json.load(open(f)); r=d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],4), r.get('frac'), d.get('clocks',{}).get('sm_mhz'), round(d['e2e']['value']), d.get('parity_check',{}) and d['parity_check'].get('ok'))
    except Exception as e: print(f, 'ERR', e)
P
