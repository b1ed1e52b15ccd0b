set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_graphed.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
for v in 0 1 0 1; do
PCADV_OVERLAP_D=$v python bench.py --steps 30 --no-extras --no-cpu-baseline > gpurun_out/r2f_bench_ov$v.json 2>gpurun_out/r2f_err.txt || tail -5 gpurun_out/r2f_err.txt
python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_ov$v.json').read().strip().splitlines()[-1]); print('overlap=$v ms/step %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), d['clocks'])"
done
