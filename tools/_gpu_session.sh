# round 2 session AV: the two new pin cases (sons periodic in y and z) through both GPU routes
mkdir -p gpurun_out
timeout 30 python -m pytest tests/test_gpu_reference_golden.py -m gpu -q -x -k "periodic_son_yz" > gpurun_out/r02av_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02av_pytest.txt | cut -c1-300
