# round 2 session AS (2 GPUs): the driver's N=2 command at the final HEAD (asynchronous read-back on its own copy stream on slabs)
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02as_bench_n2_s20.json 2> gpurun_out/err_as.txt; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02as_bench_n2_s20.json')); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity_check']['ok'])"
