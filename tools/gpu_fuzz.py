#!/usr/bin/env python
"""Randomised differential run, longer than the test-suite's (developer tool, GPU):  gpu_fuzz.py [n_seeds] [first_seed]
every stream goes through the single-kernel compile, the multi-kernel compile (deferred emit status, early totals) and - up to 9 000
events - the oracle; streams the reference rejects must give the oracle's status and event index on both paths."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from c2a_loader import c2a
import oracle_lib as orc
import test_gpu_fused as tf
from test_gpu_emit import valid_stream
import pytest

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
first = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
ctx = c2a.DeviceContext(0)
t0 = time.time(); stats = {"valid": 0, "cyclic": 0, "big": 0, "rejected": 0}
for seed in range(first, first + n_seeds):
    rng = np.random.RandomState(seed)
    big = seed % 5 == 4
    n = int(rng.choice([30000, 120000])) if big else int(rng.choice([12, 60, 400, 3000, 9000]))
    ev = valid_stream(rng, n, id_order="sequential", shape=["random", "path", "star"][seed % 3], p_redundant=[0.0, 0.15, 0.5][(seed // 3) % 3])
    kinds = ev[:, 0] & 0xFF
    sigs, gate_outs = ev[kinds <= 1, 1], ev[kinds == 2, 3]
    ins = rng.choice(sigs, size=min(5, len(sigs)), replace=False).astype(np.uint32)
    outs = rng.choice(gate_outs, size=min(4, len(gate_outs)), replace=False).astype(np.uint32) if len(gate_outs) else np.zeros(0, np.uint32)
    if seed % 4 == 1 and len(ins) > 1:
        ins = np.concatenate([ins, ins[:1]]); outs = np.concatenate([outs, ins[1:2]])
    res = tf.run_both(ctx, c2a, ev, ins, outs)
    if big:
        stats["big"] += 1
    else:
        tf.check_oracle(orc, ev, ins, outs, res)
    stats["cyclic" if res is None else "valid"] += 1
    # a rejected stream of the same seed through both paths
    for fused in (True, False):
        c2a.lib.c2a_set_fused_limits(tf.FUSED_MAX if fused else 0, 0)
        try:
            tf.test_streams_the_reference_rejects_fall_back_to_the_exact_replay.__wrapped__(ctx, c2a, orc, seed) if hasattr(tf.test_streams_the_reference_rejects_fall_back_to_the_exact_replay, "__wrapped__") else tf.test_streams_the_reference_rejects_fall_back_to_the_exact_replay(ctx, c2a, orc, seed - 4000)
        finally:
            c2a.lib.c2a_set_fused_limits(tf.FUSED_MAX, 0)
    stats["rejected"] += 2
    if (seed - first) % 25 == 24:
        print(f"seed {seed}: {stats} {time.time() - t0:.0f}s", flush=True)
print("FUZZ OK", stats, f"{time.time() - t0:.0f}s")
