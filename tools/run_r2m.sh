# round 2, backward levels: GPU suite, then same-box A/B of the fused levels on the cfg5 step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -5
for lv in off on; do
  if [ $lv = off ]; then export PCADV_LEVEL=0; else export PCADV_LEVEL=1; fi
  timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extras --top-kernels 60 > gpurun_out/r2m_bench_$lv.json 2> gpurun_out/r2m_err.txt || tail -5 gpurun_out/r2m_err.txt
  python - "$lv" <<'PY'
import json, sys
lv = sys.argv[1]
d = json.loads(open("gpurun_out/r2m_bench_%s.json" % lv).read().strip().splitlines()[-1])
print("LEVELS %-28s ms/step %.3f value %.0f e2e %.0f launches %s clocks %s" % (lv, d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]["sm_mhz"]))
print(json.dumps(d["kernel_ms_per_step"]))
PY
done
