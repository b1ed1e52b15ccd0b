set -x
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 30 --no-extras --no-cpu-baseline > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_err_n1.txt
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 30 --no-extras --no-cpu-baseline > gpurun_out/r2s_bench_n$n.json 2> gpurun_out/r2s_err_n$n.txt || tail -5 gpurun_out/r2s_err_n$n.txt
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 tests/dist_graph_check.py > gpurun_out/r2s_dist_check_fp32.txt 2>&1; tail -2 gpurun_out/r2s_dist_check_fp32.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29732 tests/dist_graph_check.py --clouds-per-rank 2 > gpurun_out/r2s_dist_check_fp32_n8.txt 2>&1; tail -2 gpurun_out/r2s_dist_check_fp32_n8.txt
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/r2s_bench_n%d.json"%n).read().strip().splitlines()[-1])
        if n==1: base=d["value"]
        print("N=%d ms/step %.3f value %.0f e2e %.0f eff %.4f clocks %s"%(n,d["ms_per_step"],d["value"],d["e2e"]["value"],d["value"]/(n*base),d["clocks"]))
    except Exception as e:
        print(n,"failed",e)
PY
