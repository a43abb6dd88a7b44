mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests12.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests12.log
tail -6 gpurun_out/r2_tests12.log
timeout 900 python bench.py --steps 10 --warmup 3 --legs configs,same_config --no-from-source --no-host-emit --no-pipelined > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench12.err
python tools/show_bench.py gpurun_out/r2_bench12.json | grep -E "^value|config|same"
