# round 2, last build: one 8-GPU box, bench at N = 8 and N = 1 back to back
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 8 --steps 30 --no-extras --no-cpu-baseline > gpurun_out/r3c_bench_n8.json 2> gpurun_out/r3c_err_n8.txt || tail -5 gpurun_out/r3c_err_n8.txt
python bench.py --gpus 1 --steps 30 --no-extras --no-cpu-baseline > gpurun_out/r3c_bench_n1.json 2> gpurun_out/r3c_err_n1.txt || tail -5 gpurun_out/r3c_err_n1.txt
python - <<'PY'
import json
for n in (8, 1):
    d = json.loads(open("gpurun_out/r3c_bench_n%d.json" % n).read().strip().splitlines()[-1])
    print("N=%d ms/step %.3f value %.0f e2e %.0f by_rank %s %s" % (n, d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("ms_per_step_by_rank"), d["clocks"]))
PY
