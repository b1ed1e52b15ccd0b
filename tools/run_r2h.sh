set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
python bench.py --steps 20 > gpurun_out/r2h_bench_cfg5.json 2> gpurun_out/r2h_err.txt || tail -5 gpurun_out/r2h_err.txt
python -c "
import json; d=json.loads(open('gpurun_out/r2h_bench_cfg5.json').read().strip().splitlines()[-1]); print('ms/step %.3f value %.0f e2e %.0f'%(d['ms_per_step'], d['value'], d['e2e']['value']), d['clocks'], d['roofline']['frac'], d['drop_in'], d['fp32_mode'], d['cpu_baseline'])"
