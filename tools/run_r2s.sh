# round 2: GPU suite with the one-hot level, then same-box A/B on the cfg5 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for oh in 0 1; do
  PCADV_ONEHOT_LEVEL=$oh timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extras > gpurun_out/r2s_bench_oh$oh.json 2> gpurun_out/r2s_err.txt || tail -5 gpurun_out/r2s_err.txt
  python - "$oh" <<'PY'
import json, sys
oh = sys.argv[1]
d = json.loads(open("gpurun_out/r2s_bench_oh%s.json" % oh).read().strip().splitlines()[-1])
print("ONEHOT %s ms/step %.3f e2e %.0f clocks %s graph_check %s" % (oh, d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["graph_check"]["max_rel_diff"]))
PY
done
