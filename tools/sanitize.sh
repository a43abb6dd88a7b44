#!/bin/bash
# compute-sanitizer passes over the GPU parity suite (the million-gate cases are left out: memcheck is 10-50x slower).
# usage (on a GPU box):  tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]      default: all four
set -u
cd "$(dirname "$0")/.."
sel='not million and not wraparound and not ten_million and not properties and not full_size and not baseline_sized and not scale and not shuffled_chains'
files="tests/test_gpu_emit.py tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_kahn.py tests/test_gpu_eval.py tests/test_gpu_cli.py tests/test_gpu_sharding.py tests/test_emission_goldens.py"
for tool in "${@:-memcheck racecheck synccheck initcheck}"; do
  for t in $tool; do
    echo "== $t"
    timeout ${SANITIZE_TIMEOUT:-1500} compute-sanitizer --tool "$t" --error-exitcode 9 --print-limit 5 python -m pytest $files -x -q -m gpu -k "$sel" 2>&1 | tail -4
    echo "== $t rc=$?"
  done
done
