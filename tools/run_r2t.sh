# round 2: same-box A/B/C of the backward-level routing on the cfg5 step, two repeats
mkdir -p gpurun_out
export PCADV_ONEHOT_LEVEL=0
for rep in 1 2; do
for cfg in off seg all; do
  case $cfg in
    off) export PCADV_LEVEL=0 PCADV_LEVEL_MAX_K=128;;
    seg) export PCADV_LEVEL=1 PCADV_LEVEL_MAX_K=0;;
    all) export PCADV_LEVEL=1 PCADV_LEVEL_MAX_K=128;;
  esac
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras > gpurun_out/r2t_bench_${cfg}_$rep.json 2> gpurun_out/r2t_err.txt || tail -5 gpurun_out/r2t_err.txt
  python - "$cfg" "$rep" <<'PY'
import json, sys
cfg, rep = sys.argv[1:3]
d = json.loads(open("gpurun_out/r2t_bench_%s_%s.json" % (cfg, rep)).read().strip().splitlines()[-1])
print("LEVELS %-4s rep %s ms/step %.3f e2e %.0f clocks %s" % (cfg, rep, d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
PY
done
done
