# size sweep of the headline workload (BASELINE config 5) through the bench: 100 K ... 50 M gates
mkdir -p gpurun_out
rm -f gpurun_out/r02_sweep.jsonl
for W in 183 1832 18315 54945 91575; do
  timeout 900 python bench.py --chains $W --steps 5 --warmup 3 --legs none --no-host-emit --no-cpu-baseline --no-pipelined --no-from-source 2>/dev/null | tail -1 >> gpurun_out/r02_sweep.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_sweep.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'][:40], round(d['value'] / 1e9, 3), 'G gates/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'] / 1e9, 3), round(d['e2e']['s_per_step'] * 1e3, 2), 'ms')
PY
