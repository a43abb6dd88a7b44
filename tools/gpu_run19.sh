mkdir -p gpurun_out
python tools/fused_trace.py 1024 2>&1 | head -4
timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_emit.py tests/test_gpu_cli.py tests/test_gpu_zz_sha256.py -x -q -m gpu --deselect tests/test_gpu_emit.py::test_baseline_sized_streams_against_the_oracle_emit 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --legs configs --no-host-emit --no-cpu-baseline --no-pipelined --no-from-source > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench19.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items():
    print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('gates','value','ms_per_step','fused_kernel_ms','cpu_backend_gates_per_s','parity_vs_oracle')}, v.get('e2e',{}).get('s_per_step'))
PY
