# round 2 session A: tests, default bench (plate512), launch list, ncu --set full of the IBM=true collide and the IBM loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02h_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r02h_bench_default.json 2> gpurun_out/err_h1.txt; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02h_bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02h_bench_reference.json 2> gpurun_out/err_h2.txt; echo "ref rc=$?"; cut -c1-600 gpurun_out/r02h_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r02h_launches_plate512.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r02h_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:collide_push -s 30 -c 2 -o gpurun_out/r02h_ncu_collide_ibm python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02h_under_ncu2.log 2>&1; echo "ncu collide rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ibm_loop -s 12 -c 2 -o gpurun_out/r02h_ncu_ibm_loop python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check > gpurun_out/r02h_under_ncu3.log 2>&1; echo "ncu ibm rc=$?"
ls -la gpurun_out | tail -12
