set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -2 gpurun_out/bench_g.err
python tools/microbench.py --only chain_trunk,chain_tail,conv2,conv3,fc3,fc4 > gpurun_out/microbench_chain.log 2>&1; cat gpurun_out/microbench_chain.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1200 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_bench_g.log 2>&1
python tools/microbench.py --only chain_trunk --iters 1 > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_chain -c 1 -f -o gpurun_out/chain_trunk_r1g python tools/microbench.py --only chain_trunk --iters 1 > gpurun_out/ncu_chain.log 2>&1
ls -la gpurun_out | tail -5
