set -x
mkdir -p gpurun_out
python tools/microbench.py --only fc1,wg_fc1_g,chain_trunk,chain_tail,conv6max,dz4_mb,wg_conv5 --iters 10 2>&1 | tail -8
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -f -o gpurun_out/r2i_$1 python tools/microbench.py --only $3 --iters 1 > gpurun_out/r2i_ncu_$1.log 2>&1 || tail -3 gpurun_out/r2i_ncu_$1.log; }
cap fc1 tc_rows_lean fc1
cap head_ce softmax_head_rows head_ce
cap maxbwd_apply maxbwd_rows_apply16 maxbwd_rows
cap conv1 first_layer_kernel conv1_simt
