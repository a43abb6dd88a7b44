mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests14.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests14.log
tail -5 gpurun_out/r2_tests14.log
timeout 900 python bench.py --steps 10 --warmup 3 --legs same_config --no-from-source --no-host-emit > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench14.err
python tools/show_bench.py gpurun_out/r2_bench14.json | head -34
