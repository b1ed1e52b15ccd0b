# round 2: GPU suite, then same-box A/B of d(cbias) as a by-product of the dz5 dgrad launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
for rep in 1 2; do for v in 0 1; do
  PCADV_DCB_FROM_DGRAD=$v timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras > gpurun_out/r3n_bench_dcb${v}_$rep.json 2> gpurun_out/r3n_err.txt || tail -5 gpurun_out/r3n_err.txt
  python - "$v" "$rep" <<'PY'
import json, sys
v, rep = sys.argv[1:3]
d = json.loads(open("gpurun_out/r3n_bench_dcb%s_%s.json" % (v, rep)).read().strip().splitlines()[-1])
k = d["kernel_ms_per_step"]
print("DCB_FROM_DGRAD %s rep %s ms/step %.3f e2e %.0f clocks %s dz5 %s fc1wg %s launches %s" % (v, rep, d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], k.get("linear:tc:k256:n512:maskbits"), [x for n, x in k.items() if n.startswith("wgrad:tc:n256:k960")], d["gpu_launches"]))
PY
done; done
