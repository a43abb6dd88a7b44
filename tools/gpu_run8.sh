mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --legs none --no-from-source --no-host-emit --no-pipelined --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench8.err
python tools/show_bench.py gpurun_out/r2_bench8.json
