# round 2 session K (2 GPUs): multi-rank parity after the IBM arm removal; heave1024 x2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/multi_rank_case.py > gpurun_out/r02r_multi_rank_parity_n2.txt 2>&1; echo "parity rc=$?"
grep -c " OK" gpurun_out/r02r_multi_rank_parity_n2.txt; grep "FAIL\|Error\|error" gpurun_out/r02r_multi_rank_parity_n2.txt | head -8 | cut -c1-400; tail -3 gpurun_out/r02r_multi_rank_parity_n2.txt | cut -c1-300
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02r_bench_heave1024_n2_s20.json 2> gpurun_out/err_r1.txt; echo "bench rc=$?"; cut -c1-250 gpurun_out/r02r_bench_heave1024_n2_s20.json
