# round 2 session AQ: ncu --set full of the cooperative IBM kernel and the merged father-to-son launch (school2048r, one GPU)
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"ibm_loop_kernel|pair_f2s_kernel" -s 4 -c 3 --csv --page raw --log-file gpurun_out/r02aq_ncu_full_ibm_loop_f2s_school2048r.csv python bench.py --workload school2048r --steps 3 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02aq_under_ncu.log 2>&1; echo "ncu rc=$?"
wc -c gpurun_out/r02aq_ncu_full_ibm_loop_f2s_school2048r.csv
