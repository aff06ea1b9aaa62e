set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_c.log 2>&1; tail -3 gpurun_out/r2_pytest_c.log
python tools/profile_update.py > gpurun_out/r2_plain_u.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_update.csv python tools/profile_update.py > gpurun_out/r2_ncu_u.log 2>&1
python tools/profile_update.py --with-refresh > gpurun_out/r2_plain_ur.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_update_with_refresh.csv python tools/profile_update.py --with-refresh > gpurun_out/r2_ncu_ur.log 2>&1
python tools/gemm_one.py syrk_conv2 > gpurun_out/r2_plain_syrk.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r2_prof_syrk_conv2 python tools/gemm_one.py syrk_conv2 > gpurun_out/r2_ncu_syrk.log 2>&1
python tools/inv_check.py 32 > gpurun_out/r2_plain_inv.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:inv_resident -s 2 -c 1 -o gpurun_out/r2_prof_inv_resident python tools/inv_check.py 32 > gpurun_out/r2_ncu_inv.log 2>&1
tail -3 gpurun_out/r2_plain_syrk.log gpurun_out/r2_ncu_syrk.log gpurun_out/r2_ncu_inv.log gpurun_out/r2_ncu_u.log
ls -la gpurun_out/r2_*
