mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_emit.py tests/test_packed.py tests/test_gpu_fused.py -x -q -m gpu --deselect tests/test_gpu_emit.py::test_baseline_sized_streams_against_the_oracle_emit 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --legs none --no-host-emit --no-cpu-baseline --no-pipelined --no-from-source > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/check_bench.json 2>/dev/null | head -4
