python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 400 --warmup 20 > gpurun_out/r2_scale_n$NG.json 2> gpurun_out/r2_scale_n$NG.err || tail -20 gpurun_out/r2_scale_n$NG.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_scale_n$NG.json"))
print("N=$NG us/step", round(d["ms_per_step"]*1e3,2), "value %.4g"%d["value"], "e2e %.3g"%d["e2e"]["value"], "ceil %.3g"%d["e2e"]["host_link_ceiling"]["env_steps_per_s"])
print("strong", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["strong"].items() if k!="limiter"})
print("per_step", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["per_step_stats"].items() if k not in ("collective","side_stream_kernel")})
print(d["config"])
PY
