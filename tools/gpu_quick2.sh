mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_deep.py tests/test_gpu_fused.py tests/test_gpu_kahn.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --legs variants,deep --no-host-emit --no-cpu-baseline --no-pipelined --no-from-source > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/check_bench.json 2>/dev/null | grep -E "^value|variant|deep" | cut -c1-220
