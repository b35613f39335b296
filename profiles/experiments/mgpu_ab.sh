set -x
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 400 --warmup 20 --no-cpu "$@" > gpurun_out/r2_mg_${NG}_$tag.json 2> gpurun_out/r2_mg_${NG}_$tag.err || tail -5 gpurun_out/r2_mg_${NG}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_mg_${NG}_$tag.json"))
    print("$tag", "us/step", round(d["ms_per_step"]*1e3,2), "value", "%.4g"%d["value"], "e2e", "%.3g"%d["e2e"]["value"], "ceil", "%.3g"%d["e2e"]["host_link_ceiling"]["env_steps_per_s"], "strong", {k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get("strong") or {}).items() if k in ("us_per_step","efficiency_vs_n1","envs_per_gpu")})
except Exception as e: print("$tag FAILED", e)
PY
}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py 2>&1 | tail -2
run peer_k16 --stats-every 16 --stats-collective peer --no-strong
run peer_k1 --stats-every 1 --stats-collective peer --no-strong
run peer_ov_k1 --stats-every 1 --stats-collective peer --stats-overlap --no-strong
run peerlag_ov_k1 --stats-every 1 --stats-collective peer-lagged --stats-overlap --no-strong
run peer_ov_k16 --stats-every 16 --stats-collective peer --stats-overlap --no-strong
