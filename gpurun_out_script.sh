set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --workload c4 --steps 3 --warmup 1 > gpurun_out/bench_c4.log 2>&1; tail -1 gpurun_out/bench_c4.log
