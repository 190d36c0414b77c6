# round 2 session AU: clocks sampled through NVML during the timed region (several samples even in a 30 ms region)
mkdir -p gpurun_out
timeout 60 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r02au_bench.json 2> gpurun_out/err_au.txt; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02au_bench.json')); print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/err_au.txt | cut -c1-200
