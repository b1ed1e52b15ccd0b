set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_graphed.py tests/test_gpu_branch_parity.py -q -m gpu -s -p no:cacheprovider > gpurun_out/r2c_new_tests.log 2>&1
grep -n "^FAILED\|passed\|failed" gpurun_out/r2c_new_tests.log | tail -30
grep -n "adversarial step\|^\.*F*PointNet\|disc \|cfg\|graph vs\|one-pass vs" gpurun_out/r2c_new_tests.log | cut -c1-600
