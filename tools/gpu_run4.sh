set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests4.log
tail -15 gpurun_out/r2_tests4.log
timeout 900 python bench.py --steps 10 --warmup 3 --legs same_config --no-from-source --no-host-emit > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench4.err
python tools/show_bench.py gpurun_out/r2_bench4.json | head -40
