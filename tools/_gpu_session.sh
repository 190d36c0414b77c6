# round 2 session H (8 GPUs): multi-rank parity, heave1024 / channel256 / school2048 / school2048r at 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR tests/multi_rank_case.py > gpurun_out/r02o_multi_rank_parity_n8.txt 2>&1; echo "parity rc=$?"
grep -c " OK" gpurun_out/r02o_multi_rank_parity_n8.txt; grep "FAIL" gpurun_out/r02o_multi_rank_parity_n8.txt | head -5; tail -3 gpurun_out/r02o_multi_rank_parity_n8.txt | cut -c1-300
for w in heave1024 channel256 school2048 school2048r; do
  timeout 400 $TR bench.py --gpus 8 --workload $w --steps 200 --warmup 10 > gpurun_out/r02o_bench_${w}_n8_s200.json 2> gpurun_out/err_o_$w.txt; echo "$w rc=$?"
  python - $w <<'P'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r02o_bench_{sys.argv[1]}_n8_s200.json')); r=d['roofline']
    print(sys.argv[1], 'MLUPS', round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],4), 'alone', round(r['collide_alone']['kernel_ms'],4), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['parity_check'] and d['parity_check']['ok'])
except Exception as e: print(sys.argv[1], 'ERR', e)
P
done
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02o_bench_default_n8_s20.json 2> gpurun_out/err_o_def.txt; echo "default s20 rc=$?"
timeout 400 $TR bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02o_bench_reference_n8.json 2> gpurun_out/err_o_ref.txt; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02o_bench_reference_n8.json
