# round 2 session AT: the mirror's staging-size check in front of the asynchronous flow window (used by bench.py's e2e leg)
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_io.py -m gpu -q -x > gpurun_out/r02at_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02at_pytest.txt | cut -c1-200
timeout 60 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-parity-check > gpurun_out/r02at_bench.json 2> gpurun_out/err_at.txt; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02at_bench.json')); print(round(d['value']), d['ms_per_step'], d['e2e']['value'])"
