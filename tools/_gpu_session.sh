# round 2 session AR: after the removal of the 48-register IBM build and the separate copy stream of the asynchronous read-backs
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_io.py tests/test_gpu_ibm_exact.py -m gpu -q -x > gpurun_out/r02ar_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02ar_pytest.txt | cut -c1-200
timeout 90 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02ar_bench.json 2> gpurun_out/err_ar.txt; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02ar_bench.json')); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity_check']['ok'])"
