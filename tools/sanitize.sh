#!/bin/bash
# compute-sanitizer passes over the GPU parity suite (the million-gate cases are left out: memcheck is 10-50x slower).
# usage (on a GPU box):  tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]      default: all four
set -u
cd "$(dirname "$0")/.."
sel='not million and not wraparound and not ten_million and not properties'
for tool in "${@:-memcheck racecheck synccheck initcheck}"; do
  for t in $tool; do
    echo "== $t"
    compute-sanitizer --tool "$t" --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_emit.py tests/test_gpu_parity.py \
        tests/test_gpu_kahn.py tests/test_gpu_eval.py tests/test_gpu_cli.py -x -q -k "$sel" 2>&1 | tail -3
  done
done
