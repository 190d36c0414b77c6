# round 2 session Q: e2e stability after the read-back warm-up; ncu traffic of heave1024's collide launches (one GPU)
mkdir -p gpurun_out
for i in 1 2; do timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02z_bench_plate512_s20_$i.json 2>/dev/null; python - $i <<'P'
import json,sys
d=json.load(open(f'gpurun_out/r02z_bench_plate512_s20_{sys.argv[1]}.json')); e=d['e2e']
print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), 'e2e', round(e['value']), round(e['seconds'],4), round(e['upload_seconds'],4))
P
done
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:collide_push_kernel' -s 12 -c 2 -o gpurun_out/r02z_ncu_collide_step_heave1024 python bench.py --workload heave1024 --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02z_under_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02z_ncu_collide_step_heave1024.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum 2>/dev/null | cut -c1-400 | tail -3
