cd /root/repo
python bench.py --steps 50 --warmup 5 > gpurun_out/r2b_bench.log 2> gpurun_out/r2b_bench.err; tail -c 3000 gpurun_out/r2b_bench.err; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2b_bench.log') if x.startswith('{')]
d=json.loads(l[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','env_frames_per_sec')}); print(d['e2e']['ms_per_step']); r=d['roofline']
print(r['kernel'][:80], r['launch_ms'], r['frac']); print(r['hbm_kernel']['launch_ms'], r['hbm_kernel']['frac'], r['hbm_kernel']['conv1_factor_syrk']['launch_ms'])
print(r['stage_ms']); print(d['rollout']); print(r['conv_dgrad']['conv_dgrad_launches'])
PY
python tools/gemm_one.py syrk_conv2_gather
python tools/profile_update.py > gpurun_out/r2b_plain_u.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_update.csv python tools/profile_update.py > gpurun_out/r2b_ncu_u.log 2>&1
python tools/profile_update.py --with-refresh > gpurun_out/r2b_plain_ur.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_update_with_refresh.csv python tools/profile_update.py --with-refresh > gpurun_out/r2b_ncu_ur.log 2>&1
python tools/gemm_one.py syrk_conv2_gather > gpurun_out/r2b_plain_syrk.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -o gpurun_out/r2b_prof_syrk_conv2_gather python tools/gemm_one.py syrk_conv2_gather > gpurun_out/r2b_ncu_syrk.log 2>&1
ls -la gpurun_out/r2b_*
