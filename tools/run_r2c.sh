set -x
mkdir -p gpurun_out
for w in cfg5 cfg3 cfg1 cfg2 cfg4; do
  timeout 900 python bench.py --workload $w --steps 20 > gpurun_out/r2c_bench_$w.json 2> gpurun_out/r2c_bench_$w.err || tail -20 gpurun_out/r2c_bench_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2c_bench_$w.json").read().strip().splitlines()[-1])
    print("$w", "ms/step %.3f value %.0f e2e %.0f launches %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["launch_mode"]))
    print("   roofline", {k: d["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "frac", "launch_ms")} if d.get("roofline") else None)
    for k in ("drop_in", "fp32_mode", "graph_check", "cpu_baseline", "eager"):
        if k in d: print("  ", k, d[k])
    for r in d.get("kernel_rooflines", [])[:14]: print("    ", r)
except Exception as e:
    print("$w: no line", e)
PY
done
