# round 2: GPU suite with the swapped weight-gradient orientation, then same-box A/B on the cfg5 step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for rep in 1 2; do for sw in 0 1; do
  PCADV_WGRAD_SWAP=$sw timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras > gpurun_out/r3k_bench_sw${sw}_$rep.json 2> gpurun_out/r3k_err.txt || tail -5 gpurun_out/r3k_err.txt
  python - "$sw" "$rep" <<'PY'
import json, sys
sw, rep = sys.argv[1:3]
d = json.loads(open("gpurun_out/r3k_bench_sw%s_%s.json" % (sw, rep)).read().strip().splitlines()[-1])
k = d["kernel_ms_per_step"]
print("WGRAD_SWAP %s rep %s ms/step %.3f e2e %.0f clocks %s wgrads %s" % (sw, rep, d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], {n: v for n, v in k.items() if n.startswith("wgrad")}))
PY
done; done
