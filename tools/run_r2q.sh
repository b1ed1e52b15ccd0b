# round 2: ncu captures for the backward-level kernel (fused vs the separate dgrad / wgrad pair) and the first-layer wgrad
mkdir -p gpurun_out
python tools/level_ab.py --rounds 1 --iters 2 > gpurun_out/r2q_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'tc_level|tc_rows_lean|tc_wgrad_kernel' -c 40 -f -o gpurun_out/r2q_levels python tools/level_ab.py --rounds 1 --iters 1 --only-head > gpurun_out/r2q_ncu.log 2>&1
tail -2 gpurun_out/r2q_ncu.log
python tools/microbench.py --points 2097152 --only conv1_wgrad --iters 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:first_layer_wgrad -c 2 -f -o gpurun_out/r2q_conv1_wgrad python tools/microbench.py --points 2097152 --only conv1_wgrad --iters 1 > gpurun_out/r2q_ncu2.log 2>&1
tail -2 gpurun_out/r2q_ncu2.log
