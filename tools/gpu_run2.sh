set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/r2_fused_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_fused_tests.log
tail -30 gpurun_out/r2_fused_tests.log
