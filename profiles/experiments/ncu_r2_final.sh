# Final-build captures of the headline command (graph-replayed PD loop): launch list + one full set of the PD kernel.
set -x
B="python bench.py --steps 20 --warmup 3 --no-families --no-cpu --no-copy-probe"
$B > gpurun_out/r2f_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_pd_bench.csv $B > gpurun_out/r2f_ncu_launches.log 2>&1
$B > gpurun_out/r2f_plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pd_torque_vec4 -s 60 -c 2 -o gpurun_out/r2f_pd_full $B > gpurun_out/r2f_ncu_pd.log 2>&1
ls -la gpurun_out/r2f_*
