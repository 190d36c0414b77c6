set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/multi_rank_case.py > gpurun_out/r02_multi_rank_parity_n2.txt 2>&1; echo "parity rc=$?"
grep -c OK gpurun_out/r02_multi_rank_parity_n2.txt; grep -c FAIL gpurun_out/r02_multi_rank_parity_n2.txt; tail -4 gpurun_out/r02_multi_rank_parity_n2.txt
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 --workload channel256 > gpurun_out/r02d_bench_channel256_n2.json 2> gpurun_out/err_d1.txt
timeout 300 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r02d_bench_heave1024_n2.json 2> gpurun_out/err_d2.txt
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_heave1024_n2_s20.json 2> gpurun_out/err_d3.txt
timeout 300 $TR bench.py --gpus 2 --steps 100 --warmup 10 --workload school2048 > gpurun_out/r02d_bench_school2048_n2.json 2> gpurun_out/err_d4.txt
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02d_bench_ref_n2.json 2> gpurun_out/err_d5.txt
for f in gpurun_out/r02d_bench_*n2*.json; do python - $f <<'P'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get('roofline') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],4), r.get('frac'), (r.get('collide_alone') or {}).get('kernel_ms'), d.get('parity_check',{}) and d['parity_check'].get('ok'), d.get('cpu_baseline'))
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
tail -3 gpurun_out/err_d*.txt
