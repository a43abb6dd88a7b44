mkdir -p gpurun_out
for st in packed packed6; do
timeout 900 python bench.py --steps 10 --warmup 3 --legs none --no-from-source --no-host-emit --no-pipelined --no-cpu-baseline --stream $st > gpurun_out/r2_bench5_$st.json 2> gpurun_out/r2_bench5.err; echo "bench $st rc=$?"
python tools/show_bench.py gpurun_out/r2_bench5_$st.json | head -12
done
