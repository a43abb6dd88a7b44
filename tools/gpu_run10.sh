mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_deep.py -x -q --durations=10 > gpurun_out/r2_deep.log 2>&1; echo "deep rc=$?"
tail -40 gpurun_out/r2_deep.log
