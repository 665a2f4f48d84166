set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --large > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
timeout 900 python bench.py --workload c5 --steps 5 --warmup 2 > gpurun_out/bench_c5_n1.log 2>&1; tail -1 gpurun_out/bench_c5_n1.log
