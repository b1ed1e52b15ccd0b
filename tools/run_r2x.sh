# round 2, last build: guard-band test of the level kernel, bench lines of the other BASELINE configs
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -q -x -k "overrun" -p no:cacheprovider 2>&1 | tail -2
for w in cfg1 cfg2 cfg3 cfg4; do
  timeout 300 python bench.py --workload $w --steps 20 > gpurun_out/r2x_bench_$w.json 2> gpurun_out/r2x_err_$w.txt || tail -3 gpurun_out/r2x_err_$w.txt
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
d = json.loads(open("gpurun_out/r2x_bench_%s.json" % w).read().strip().splitlines()[-1])
print("%s ms/step %.3f value %.0f e2e %.0f roofline %s %.3f cpu %.1f" % (w, d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel"][:40], d["roofline"]["frac"], d["cpu_baseline"]["value"]))
PY
done
