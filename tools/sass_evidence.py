#!/usr/bin/env python
"""Per-kernel SASS evidence of the in-tree libc2a.so (cuobjdump -sass): TMA bulk copies (UBLKCP), mbarrier ops (SYNCS),
128-bit global accesses, reductions / atomics.  Writes profiles/r02_sass_tma.txt.

    python tools/sass_evidence.py [out_path]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "circom-2-arithc_b200", "libc2a.so")
MNEMONICS = ["UBLKCP", "UTMALDG", "SYNCS", "LDG.128", "STG.128", "RED", "ATOMG", "LDGSTS", "CCTL", "MEMBAR", "BAR.SYNC", "UCGABAR"]
PATTERNS = {"LDG.128": r"LDG\.E(\.NA)?\.128", "STG.128": r"STG\.E(\.NA)?\.128", "RED": r"\bRED\.E", "ATOMG": r"\bATOMG\."}


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_tma.txt")
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    cur, cnt, ninstr = None, collections.defaultdict(collections.Counter), collections.Counter()
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur is None or "/*" not in ln:
            continue
        if re.search(r"/\*[0-9a-f]{4}\*/", ln):
            ninstr[cur] += 1
        for mn in MNEMONICS:
            if re.search(PATTERNS.get(mn, re.escape(mn)), ln):
                cnt[cur][mn] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(ninstr), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(ninstr, demangle))
    lines = [f"# cuobjdump -sass circom-2-arithc_b200/libc2a.so : arch {arch}, {len(ninstr)} kernels",
             "# kernel | SASS instructions | " + " | ".join(MNEMONICS)]
    for k in sorted(ninstr, key=lambda x: names[x]):
        short = re.sub(r"\(.*", "", names[k]).replace("c2a::", "")
        lines.append(f"{short} | {ninstr[k]} | " + " | ".join(str(cnt[k].get(m, 0)) for m in MNEMONICS))
    tma = [re.sub(r"\(.*", "", names[k]).replace("c2a::", "") for k in ninstr if cnt[k].get("UBLKCP")]
    lines.append(f"# kernels with TMA bulk copies (UBLKCP): {', '.join(sorted(tma))}")
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"{out}: {len(ninstr)} kernels, TMA in {sorted(tma)}")


if __name__ == "__main__":
    main()
