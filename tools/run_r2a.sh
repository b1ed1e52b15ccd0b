# round 2, call A: new parity tests, regression of the existing GPU suite, gradient-error table, bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_graphed.py tests/test_gpu_branch_parity.py -q -m gpu -s -p no:cacheprovider > gpurun_out/r2a_new_tests.log 2>&1
tail -60 gpurun_out/r2a_new_tests.log
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_graphed.py --deselect tests/test_gpu_branch_parity.py -p no:cacheprovider 2>&1 | tail -15
timeout 300 python tests/check_grad_error_vs_f64.py > gpurun_out/r2a_grad_error_vs_float64.txt 2> gpurun_out/r2a_grad_error.err; tail -3 gpurun_out/r2a_grad_error.err; tail -32 gpurun_out/r2a_grad_error_vs_float64.txt
timeout 400 python tests/check_grad_error_vs_f64.py --clouds 8 --points 2048 > gpurun_out/r2a_grad_error_vs_float64_8x2048.txt 2>> gpurun_out/r2a_grad_error.err; tail -3 gpurun_out/r2a_grad_error_vs_float64_8x2048.txt
timeout 600 python bench.py --steps 20 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err; cut -c1-1500 gpurun_out/r2a_bench.json
