# same-box A/B: fused backward levels x discriminator phase on its own stream
mkdir -p gpurun_out
for lv in 0 1; do for ov in 0 1; do
  PCADV_LEVEL=$lv PCADV_OVERLAP_D=$ov timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-extras > gpurun_out/r2n_bench_l${lv}_o${ov}.json 2> gpurun_out/r2n_err.txt || tail -5 gpurun_out/r2n_err.txt
  python - "$lv" "$ov" <<'PY'
import json, sys
lv, ov = sys.argv[1:3]
d = json.loads(open("gpurun_out/r2n_bench_l%s_o%s.json" % (lv, ov)).read().strip().splitlines()[-1])
print("LEVEL %s OVERLAP_D %s ms/step %.3f e2e %.0f clocks %s" % (lv, ov, d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
PY
done; done
