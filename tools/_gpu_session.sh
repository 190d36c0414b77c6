# round 2 session N (4 GPUs): the driver's own commands at N=4, both arms
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --impl reference --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02u_bench_reference_n4.json 2> gpurun_out/err_u0.txt; echo "ref rc=$?"; cut -c1-200 gpurun_out/r02u_bench_reference_n4.json
timeout 400 $TR bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02u_bench_default_n4_s20.json 2> gpurun_out/err_u1.txt; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r02u_bench_default_n4_s20.json')); r=d['roofline']
print(d['config']['workload'], round(d['value']), round(d['ms_per_step'],4), round(r['frac'],4), round(d['e2e']['value']), d['parity_check']['ok'], d['clocks'])
P
