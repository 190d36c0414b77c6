# round 2 session G (2 GPUs): e2e laps of heave1024 x2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r02n_bench_heave1024_n2_s200.json 2> gpurun_out/err_n1.txt; echo "bench n2 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r02n_bench_heave1024_n2_s200.json')); print(d['ms_per_step'], d['e2e'])
P
timeout 400 $TR bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02n_bench_heave1024_n2_s100.json 2> gpurun_out/err_n2.txt; echo "bench n2 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r02n_bench_heave1024_n2_s100.json')); print(d['ms_per_step'], d['e2e'])
P
