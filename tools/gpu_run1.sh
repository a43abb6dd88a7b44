set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests1.log
tail -3 gpurun_out/r2_tests1.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench1.err
python tools/show_bench.py gpurun_out/r2_bench1.json 2>/dev/null | head -50 || true
