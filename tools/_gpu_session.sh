# round 2 session I (2 GPUs): multi-rank parity with sons across the interface
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/multi_rank_case.py > gpurun_out/r02p_multi_rank_parity_n2.txt 2>&1; echo "parity rc=$?"
grep -c " OK" gpurun_out/r02p_multi_rank_parity_n2.txt; grep "FAIL\|Error\|error\|refinement" gpurun_out/r02p_multi_rank_parity_n2.txt | head -8 | cut -c1-400; tail -4 gpurun_out/r02p_multi_rank_parity_n2.txt | cut -c1-300
