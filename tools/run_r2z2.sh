# round 2: GPU suite with three chain groups, then same-box A/B of the chain groups on the cfg5 step (two repeats)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for rep in 1 2; do for g in 2 3; do
  PCADV_CHAIN_GROUPS=$g timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras > gpurun_out/r2z2_bench_g${g}_$rep.json 2> gpurun_out/r2z2_err.txt || tail -5 gpurun_out/r2z2_err.txt
  python - "$g" "$rep" <<'PY'
import json, sys
g, rep = sys.argv[1:3]
d = json.loads(open("gpurun_out/r2z2_bench_g%s_%s.json" % (g, rep)).read().strip().splitlines()[-1])
k = d["kernel_ms_per_step"]
print("CHAIN_GROUPS %s rep %s ms/step %.3f e2e %.0f clocks %s chains %s" % (g, rep, d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], {n: v for n, v in k.items() if n.startswith("chain")}))
PY
done; done
