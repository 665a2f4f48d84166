set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_target.py solve c2 10000 500 2 > gpurun_out/t_solve_c2.log 2>&1; tail -4 gpurun_out/t_solve_c2.log
timeout 300 python tools/profile_target.py solve c3 100000 1000 2 > gpurun_out/t_solve_c3.log 2>&1; tail -4 gpurun_out/t_solve_c3.log
