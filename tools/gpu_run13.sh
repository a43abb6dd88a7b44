mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kahn.py tests/test_gpu_eval.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --legs sweeps --no-from-source --no-host-emit --no-pipelined --no-cpu-baseline > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2_bench13.json | grep -E "^value|sweeps"
