# round 2 session M: the Fortran driver over the shim, executed by the interpreter against the CUDA library
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fortran_shim.py -q -rs > gpurun_out/r02t_pytest_fortran_shim.txt 2>&1; echo "rc=$?"; tail -30 gpurun_out/r02t_pytest_fortran_shim.txt | cut -c1-400
