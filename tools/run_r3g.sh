# round 2: ncu --set full of the chain kernel with three tiles in flight (trunk, discriminator chain) and of a head level
mkdir -p gpurun_out
python tools/microbench.py --only chain_trunk,chain_disc,lv_fc4 --iters 1 > gpurun_out/r3g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'tc_chain_kernel|tc_level_kernel' -s 3 -c 3 -f -o gpurun_out/r3g_chain_level python tools/microbench.py --only chain_trunk,chain_disc,lv_fc4 --iters 1 > gpurun_out/r3g_ncu.log 2>&1
tail -2 gpurun_out/r3g_ncu.log
