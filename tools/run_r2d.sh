set -x
mkdir -p gpurun_out
python bench.py --steps 10 --no-extras --no-cpu-baseline --top-kernels 80 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log | cut -c1-300
