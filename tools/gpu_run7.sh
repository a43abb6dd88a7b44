mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests7.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests7.log
tail -4 gpurun_out/r2_tests7.log
timeout 900 python bench.py --steps 10 --warmup 3 --legs variants,kahn --no-from-source --no-host-emit --no-pipelined > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench7.err
python tools/show_bench.py gpurun_out/r2_bench7.json
