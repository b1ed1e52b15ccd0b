# round 2: same-box A/B of an older tree against the working tree.  Before the gpurun call, on the build box:
#   git worktree add /tmp/oldtree <commit> && (cd /tmp/oldtree && python -c 'from adversarial_learning_on_pointclouds_b200 import _build; _build.build()')
#   mkdir _ab_old && cp -r /tmp/oldtree/{adversarial_learning_on_pointclouds_b200,bench.py,oracle,BASELINE.json} MEASURED_PEAKS.json _ab_old/
# (_ab_old/ is scratch: excluded from git, it only travels with the gpurun snapshot)
mkdir -p gpurun_out
for rep in 1 2 3; do
  (cd _ab_old && timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras > ../gpurun_out/r3j_old_$rep.json 2> ../gpurun_out/r3j_old_err.txt) || tail -3 gpurun_out/r3j_old_err.txt
  timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-extras > gpurun_out/r3j_new_$rep.json 2> gpurun_out/r3j_new_err.txt || tail -3 gpurun_out/r3j_new_err.txt
  python - "$rep" <<'PY'
import json, sys
rep = sys.argv[1]
for v in ("old", "new"):
    d = json.loads(open("gpurun_out/r3j_%s_%s.json" % (v, rep)).read().strip().splitlines()[-1])
    print("%s rep %s ms/step %.3f value %.0f e2e %.0f clocks %s" % (v, rep, d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"]))
PY
done
