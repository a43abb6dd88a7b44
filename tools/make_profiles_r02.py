#!/usr/bin/env python
"""Round-2 evidence set gpurun_out/r02_* -> profiles/r02_* (+ traffic.json).  Developer tool; tools/gpu_evidence.sh produces the set."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = lambda n: os.path.join(ROOT, "gpurun_out", "r02_" + n)
P = lambda n: os.path.join(ROOT, "profiles", n)
L = lambda p: json.loads(open(p).read().strip().splitlines()[-1])
shutil.copy(G("bench.json"), P("r02_bench.json"))
shutil.copy(G("bench_ref.json"), P("r02_bench_reference_arm.json"))
shutil.copy(G("launches.csv"), P("r02_launches_bench.csv"))
for n in (2, 4, 8):
    if os.path.exists(G(f"bench_{n}gpu.json")):
        shutil.copy(G(f"bench_{n}gpu.json"), P(f"r02_bench_{n}gpu.json"))
raw = subprocess.run(["ncu", "-i", G("full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open("/tmp/raw_r02.csv", "w").write(raw)
table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_table.py"), "/tmp/raw_r02.csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
def nbytes(r, k):
    v = float(r[ix[k]].replace(",", "")); u = rows[1][ix[k]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
traffic = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, bytes (largest launch of each kernel); ncu --set full, round 2 capture "
                    "(profiles/r02_ncu_full_summary.md); workload mimc_chains W=18315 late (10 018 305 gates), packed stream with implicit operands (4 B/event)"}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("c2a::", "").replace("void ", "")
    name = name.split("<")[0]
    name = {"k_pk_scatter_t": "k_ev_scatter", "k_pk_count": "k_ev_count", "k_msf_pick_first": "k_msf_pick", "k_scan_u32_t": "k_scan_u32", "k_msf_hook_t": "k_msf_hook",
            "k_level_pass": "k_level_sort", "k_deps_t": "k_deps", "k_relax_loop": "k_relax"}.get(name, name)
    t = nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum")
    if name not in traffic or t > traffic[name]: traffic[name] = int(t)
json.dump(traffic, open(P("traffic.json"), "w"), indent=1)
# launch list: shares
lrows = list(csv.reader(l for l in open(P("r02_launches_bench.csv")) if not l.startswith("==")))
lh = lrows[0]; lx = {h: i for i, h in enumerate(lh)}
acc = collections.defaultdict(lambda: [0, 0.0])
for x in lrows[1:]:
    if len(x) < len(lh) or x[lx["Metric Name"]] != "gpu__time_duration.sum": continue
    name = x[lx["Kernel Name"]].split("(")[0].replace("c2a::", "").replace("void ", "")
    v = float(x[lx["Metric Value"]].replace(",", "")); u = x[lx["Metric Unit"]]
    acc[name][0] += 1; acc[name][1] += v / 1000 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000)
ours = {k: v for k, v in acc.items() if k.startswith("k_")}
other = {k: v for k, v in acc.items() if not k.startswith("k_")}
tot = sum(v[1] for v in ours.values())
launch = "\n".join(f"| `{k}` | {v[0]} | {v[1]/v[0]:.1f} | {v[1]/tot:.3f} |" for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]))
d = L(P("r02_bench.json")); r = d["roofline"]
kern = "\n".join(f"| `{k}` | {v:.4f} | {r['per_kernel_gbs'].get(k, '')} | {v / d['ms_per_step']:.3f} |" for k, v in sorted(r["per_kernel_ms"].items(), key=lambda x: -x[1]))
open(P("r02_ncu_full_summary.md"), "w").write(f"""# Round 2: `ncu --set full --clock-control none --import-source on` of every kernel (B200, sm_100a)

Command: `ncu --set full --clock-control none --import-source on -k regex:k_ -s 43 -c 80 -o r02_full python tools/ncu_step.py`
(second pass of: emit + build of the 10 018 305-gate headline stream through the multi-kernel pipeline; the fused single-kernel
compile of the SHA-256-shaped circuit, 115 920 gates; the Kahn levels of the 10 M gate vector).  Times under ncu are cold-cache and
serialised: compare SHARES with the bench line, not absolutes.  DRAM MB = `dram__bytes_read.sum` / `dram__bytes_write.sum`.

{table}
Reading:
* `k_pk_scatter_t<1>` (the dominant kernel; `<2>` is the all-flagged instantiation, which exits on this stream): 499 MB of DRAM
  traffic for 542 MB algorithmic - no re-read waste; 134.5 M warp instructions, `smsp__issue_active` 64 %, DRAM 31 %, 128-thread CTAs, 80 registers / 6 CTAs per SM:
  instruction-issue-bound (profiles/r02_ncu_scatter_formats.md).  Alone (here) 194 us; inside the step 215-218 us, because the side stream
  fills 370 MB of union-find arrays at the same time (deliberately: a bandwidth-bound fill next to an issue-bound kernel).
* `k_pk_count` 17 us (was 36 us: per-byte tests and POPC replaced by byte-parallel arithmetic); `k_scan_u32_multi`: the three tile-count scans in one 30-CTA launch.
* streaming kernels (`k_ev_gates`, `k_gather`, `k_producer`, `k_deps_t`, `k_ev_finalize`, `k_wire_assign`): 49-76 % of DRAM peak (4.0-6.2 TB/s physical).
* `k_ev_nid_edges` (2.7 TB/s, L2 hit 38 %) and `k_wire_first` (49 % warps active at 59 registers) stay latency-bound on random 32-byte sectors.
* `k_relax_loop` / `k_tree_blocks` (cooperative, 592 CTAs): 11 us / 5 us on the headline stream (18 315 forward edges, no big block).
* `k_fused_compile`: 148 CTAs x 1024 threads, 64 registers, 53 us for 115 920 gates - ~17 grid barriers; DRAM traffic 1.2 MB: the whole circuit lives in L2.
* `k_kahn_walk_roots`: 867 us at 6 % DRAM: a 547-hop dependency chain per MiMC component (latency floor), see DESIGN.md.

## Launch list of the bench command (shares)

`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv python bench.py --steps 2 --warmup 3 --legs configs --extra-steps 3 --no-cpu-baseline --no-host-emit --no-pipelined --no-from-source`
-> `profiles/r02_launches_bench.csv` ({sum(v[0] for v in acc.values())} launches; {sum(v[0] for v in other.values())} of them not this repo's kernels: {', '.join(sorted(other)) or 'none'}).

| kernel | launches | mean us | share of kernel time |
|---|---:|---:|---:|
{launch}

## Per-kernel table of the bench line (`profiles/r02_bench.json`: {d['ms_per_step']:.3f} ms/step; CUDA events, extra steps after the timed region)

| phase | ms per step | algorithmic GB/s | share of the step |
|---|---:|---:|---:|
{kern}
""")
print("profiles written; dominant", r["kernel"], r["frac"], "traffic", traffic.get(r["kernel"]))
