set -x
mkdir -p gpurun_out
for mode in ex noex; do
  if [ $mode = noex ]; then export PCADV_BENCH_NO_EXCHANGE=1; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 8 --steps 30 --no-extras --no-cpu-baseline > gpurun_out/r2s2_bench_n8_$mode.json 2> gpurun_out/r2s2_err_$mode.txt || tail -5 gpurun_out/r2s2_err_$mode.txt
done
unset PCADV_BENCH_NO_EXCHANGE
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29742 tests/dist_graph_check.py --clouds-per-rank 2 > gpurun_out/r2s2_dist_check_n8.txt 2>&1; grep -m3 "NVLS\|Connected all\|DIST_GRAPH" gpurun_out/r2s2_dist_check_n8.txt | cut -c1-200
python - <<'PY'
import json
for m in ("ex","noex"):
    d=json.loads(open("gpurun_out/r2s2_bench_n8_%s.json"%m).read().strip().splitlines()[-1])
    print(m,"ms/step %.3f"%d["ms_per_step"], d["ms_per_step_by_rank"], d["gradient_exchange"][:40])
PY
