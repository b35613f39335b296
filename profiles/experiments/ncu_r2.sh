set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
B="python bench.py --steps 20 --warmup 3 --no-families --no-cpu"
$B > gpurun_out/r2_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_pd_bench.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
$B > gpurun_out/r2_plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pd_torque_vec4 -s 60 -c 2 -o gpurun_out/r2_pd_full $B > gpurun_out/r2_ncu_pd.log 2>&1
python profiles/run_family.py osc --n 262144 --iters 4 > gpurun_out/r2_plain_osc262.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:osc_kernel -s 1 -c 2 -o gpurun_out/r2_osc_262k python profiles/run_family.py osc --n 262144 --iters 4 > gpurun_out/r2_ncu_osc262.log 2>&1
python profiles/run_family.py servo --n 65536 --iters 4 > gpurun_out/r2_plain_servo.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:servo_step -s 1 -c 2 -o gpurun_out/r2_servo_c2 python profiles/run_family.py servo --n 65536 --iters 4 > gpurun_out/r2_ncu_servo.log 2>&1
python profiles/run_family.py servo --stats --iters 4 > gpurun_out/r2_plain_servo_s.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:servo_step -s 1 -c 2 -o gpurun_out/r2_servo_stats python profiles/run_family.py servo --stats --iters 4 > gpurun_out/r2_ncu_servo_s.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
