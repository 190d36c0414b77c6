TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/multi_rank_case.py > gpurun_out/r02_multi_rank_parity_n2.txt 2>&1; echo "parity rc=$?"
grep -c OK gpurun_out/r02_multi_rank_parity_n2.txt; grep FAIL gpurun_out/r02_multi_rank_parity_n2.txt | head -5; tail -3 gpurun_out/r02_multi_rank_parity_n2.txt
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload channel256 --trace-out gpurun_out/r02g_trace_channel256_n2 > gpurun_out/r02g_bench_channel256_n2_s20.json 2> gpurun_out/err_g1.txt
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --trace-out gpurun_out/r02g_trace_heave1024_n2 > gpurun_out/r02g_bench_heave1024_n2_s20.json 2> gpurun_out/err_g2.txt
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r02g_bench_heave1024_n2_s200.json 2> gpurun_out/err_g3.txt
for f in gpurun_out/r02g_bench_*.json; do python - $f <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],4), r.get('frac'), (r.get('collide_alone') or {}).get('kernel_ms'), d['parity_check'] and d['parity_check']['ok'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
tail -3 gpurun_out/err_g1.txt
