# round-2 evidence run (one gpurun call): GPU tests, bench, launch lists, ncu captures.  Outputs under gpurun_out/r2f_*.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err; tail -c 400 gpurun_out/r2f_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_bench_ref.log 2> gpurun_out/r2f_bench_ref.err
python tools/profile_update.py > gpurun_out/r2f_plain_u.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_update.csv python tools/profile_update.py > gpurun_out/r2f_ncu_u.log 2>&1
python tools/profile_update.py --with-refresh > gpurun_out/r2f_plain_ur.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_update_with_refresh.csv python tools/profile_update.py --with-refresh > gpurun_out/r2f_ncu_ur.log 2>&1
python tools/gemm_one.py syrk_conv2_gather > gpurun_out/r2f_plain_syrk.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r2f_prof_syrk_conv2_gather python tools/gemm_one.py syrk_conv2_gather > gpurun_out/r2f_ncu_syrk.log 2>&1
python tools/conv_one.py f1 3 5 > gpurun_out/r2f_plain_f1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r2f_prof_conv1_fwd python tools/conv_one.py f1 3 5 > gpurun_out/r2f_ncu_f1.log 2>&1
tail -2 gpurun_out/r2f_plain_syrk.log gpurun_out/r2f_plain_f1.log
ls -la gpurun_out/r2f_*
