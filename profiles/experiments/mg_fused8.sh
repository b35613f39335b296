# N GPUs ($NG): slice / statistics check, then the default bench (statistics published by the PD kernel) at 400 and at the driver's 20 steps
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py 2>&1 | tail -2
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --no-cpu "$@" > gpurun_out/r2_mgf_${NG}_$tag.json 2> gpurun_out/r2_mgf_${NG}_$tag.err || tail -5 gpurun_out/r2_mgf_${NG}_$tag.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_mgf_${NG}_$tag.json"))
    print("$tag", "us/step", round(d["ms_per_step"]*1e3,2), "value", "%.4g"%d["value"], d.get("stats_check"), d.get("step_issue"), "e2e %.3g"%d["e2e"]["value"])
    if d.get("per_step_stats"): print("   forms", {k: round(v,2) for k,v in d["per_step_stats"]["us_per_step_by_exchange_form"].items()})
    if d.get("strong"): print("   strong", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["strong"].items() if k in ("us_per_step","efficiency_vs_n1","envs_per_gpu")})
except Exception as e: print("$tag FAILED", e)
PY
}
run fused_400 --steps 400 --warmup 20
run fused_20 --steps 20 --warmup 5 --no-strong
