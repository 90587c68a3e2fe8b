python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/bench_loader.py --clips 2048 --json gpurun_out/r02_loader_timing.json
python bench.py --workload vae --steps 3 --warmup 3 > gpurun_out/bench_vae_r2.json 2> gpurun_out/bench_vae_r2.err; tail -2 gpurun_out/bench_vae_r2.err | cut -c1-200
python bench.py --workload reach --steps 3 --warmup 3 > gpurun_out/bench_reach_r2.json 2> gpurun_out/bench_reach_r2.err; tail -2 gpurun_out/bench_reach_r2.err | cut -c1-200
python - <<PY
import json
for w in ("vae","reach"):
    d=json.load(open(f"gpurun_out/bench_{w}_r2.json")); print(w, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],1), "upd frac", round(d["roofline"]["frac"],3), d["phases_ms"], "cpu", d.get("cpu_baseline",{}).get("value"))
PY
python tools/profile_rollout.py --envs 2048 --time
