# round 2: GPU suite + same-box A/B of the CTA-pair weight multicast on the cfg5 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for pr in 0 1; do
  PCADV_ROWS_PAIR=$pr timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extras > gpurun_out/r2p_bench_pair$pr.json 2> gpurun_out/r2p_err.txt || tail -5 gpurun_out/r2p_err.txt
  python - "$pr" <<'PY'
import json, sys
pr = sys.argv[1]
d = json.loads(open("gpurun_out/r2p_bench_pair%s.json" % pr).read().strip().splitlines()[-1])
print("PAIR %s ms/step %.3f e2e %.0f clocks %s graph_check %s" % (pr, d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["graph_check"]["max_rel_diff"]))
PY
done
