set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --legs configs,same_config --no-from-source --no-host-emit --no-pipelined > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench2.err
python tools/show_bench.py gpurun_out/r2_bench2.json | grep -E "config|same_config|value"
