mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu.py -m gpu -q -x -k "fit or config2 or ragged or fuzz or config3" > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
echo c5; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fit_dense_vec -c 3 --csv python tools/profile_target.py fit c5 1000000 2000 3 2>/dev/null | grep fit_dense | awk -F'","' '{print substr($5,1,34), $NF}'
echo c3; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fit_dense_vec -c 3 --csv python tools/profile_target.py fit c3 100000 1000 3 2>/dev/null | grep fit_dense | awk -F'","' '{print substr($5,1,34), $NF}'
