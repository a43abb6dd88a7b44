mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_emit.py::test_baseline_sized_streams_against_the_oracle_emit 2>&1 | tail -3
python tools/timeline.py 2>&1 | grep timeline | tail -30 | cut -c16- | awk '{ if ($NF+0 > 0.004 || 1) print }' | awk '$(NF-1)+0 >= 0.006'
timeout 900 python bench.py --steps 10 --warmup 3 --legs none --no-host-emit --no-cpu-baseline --no-pipelined --no-from-source > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/check_bench.json 2>/dev/null | head -3
