mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_emit.py -x -q -k "baseline_sized" --durations=3 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --legs none --no-host-emit --no-pipelined --no-cpu-baseline > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r2_bench15.json | grep -E "^value|from_source"
