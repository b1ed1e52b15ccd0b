set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -3
python tools/microbench.py --only maxbwd_dw,maxbwd_rows,head_ce,head_lsm,conv1_simt,conv1_wgrad,wg_d2,rowmax_wgrad 2>&1 | tail -9
python bench.py --steps 30 --no-extras --no-cpu-baseline --top-kernels 40 > gpurun_out/r2g_bench.json 2>gpurun_out/r2g_err.txt || tail -5 gpurun_out/r2g_err.txt
python -c "
import json; d=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1]); print('ms/step %.3f e2e %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step']), d['clocks']); print({k:v for k,v in d['kernel_ms_per_step'].items() if 'maxpool' in k or 'head' in k or 'simt' in k})"
