# round 2, final records: GPU suite, smoke, bench line, launch list of one eager step, ncu of the first layer
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 > gpurun_out/r3o_bench_cfg5.json 2> gpurun_out/r3o_err.txt || tail -5 gpurun_out/r3o_err.txt
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r3o_bench_reference.json 2>> gpurun_out/r3o_err.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 600 --csv --log-file gpurun_out/r3o_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/r3o_ncu.log 2>&1

python tools/microbench.py --only head_ce,head_lsm,lsm_bwd,conv1_simt,conv1_wgrad,maxbwd_dw,maxbwd_rows,chain_trunk,chain_tail,fc1,wg_fc1_g,conv6max,lv_fc4,lv_fc3,lv_fc2,lv5,lv4,lv3,lv1,dz4_mb,dz_fc1_mb > gpurun_out/r3o_microbench.txt 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3o_bench_cfg5.json").read().strip().splitlines()[-1])
print("ms/step %.3f value %.0f e2e %.0f roofline %.3f launches %s"%(d["ms_per_step"],d["value"],d["e2e"]["value"],d["roofline"]["frac"],d["gpu_launches"]))
print(open("gpurun_out/r3o_bench_reference.json").read()[:300])
PY
