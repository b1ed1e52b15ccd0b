set -x
mkdir -p gpurun_out
python tools/microbench.py --only head_ce,head_lsm,lsm_bwd,rowmax_wgrad,rowmax_dgrad,conv1_simt,conv1_wgrad,maxbwd_dw,maxbwd_rows,rowmax_bwd,amax,wg_d2,d2_bits,chain_trunk,chain_tail > gpurun_out/r2e_microbench.txt 2>&1
cat gpurun_out/r2e_microbench.txt
cap() { # name kernel-regex case
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -f -o gpurun_out/r2e_$1 python tools/microbench.py --only $3 --iters 1 > gpurun_out/r2e_ncu_$1.log 2>&1 || tail -3 gpurun_out/r2e_ncu_$1.log
}
cap head_ce softmax_head_rows head_ce
cap maxbwd_apply maxbwd_rows_apply16 maxbwd_rows
cap maxbwd_dw maxbwd_dw16 maxbwd_dw
cap conv1 first_layer_kernel conv1_simt
cap conv1_wgrad first_layer_wgrad conv1_wgrad
cap rowmax_wgrad rowmax_wgrad_sorted rowmax_wgrad
cap rowmax_dgrad rowmax_dgrad rowmax_dgrad
cap wg_d2 tc_wgrad_kernel wg_d2
ls -la gpurun_out/*.ncu-rep | tail -12
