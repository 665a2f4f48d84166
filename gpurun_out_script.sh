set -x
mkdir -p gpurun_out
python tools/profile_target.py solve c2 10000 500 3 > gpurun_out/t_solve_c2.log 2>&1; tail -3 gpurun_out/t_solve_c2.log
python tools/profile_target.py fit c3 100000 1000 5 > gpurun_out/t_fit_c3.log 2>&1; tail -5 gpurun_out/t_fit_c3.log
python tools/profile_target.py fit c5 1000000 2000 5 > gpurun_out/t_fit_c5.log 2>&1; tail -5 gpurun_out/t_fit_c5.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pack_kernel -c 1 -o gpurun_out/pack_c2_r01 python tools/profile_target.py solve c2 10000 500 1 > gpurun_out/ncu_pack.log 2>&1; tail -2 gpurun_out/ncu_pack.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_matrix_kernel -s 1 -c 1 -o gpurun_out/fit_c3_r01 python tools/profile_target.py fit c3 100000 1000 2 > gpurun_out/ncu_fit.log 2>&1; tail -2 gpurun_out/ncu_fit.log
ls -la gpurun_out
