#!/usr/bin/env python
"""Per-kernel SASS evidence for libmerlin_b200.so: counts of the Blackwell-specific instructions (tcgen05 MMA =
UTCHMMA, TMEM loads/stores = LDTM/STTM, TMA = UTMALDG/UTMASTG/UTMAREDG, bulk copies = UBLKCP, legacy HMMA) per
kernel, plus a short excerpt around the first UTCHMMA of the main kernels.  Writes profiles/r2_sass_summary.txt.
usage: python scripts/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "merlin_b200", "libmerlin_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
ops = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU.EX2",
       "REDG", "RED.E", "ATOMG", "STL", "LDL"]
kern, counts, excerpts = None, collections.OrderedDict(), {}
lines = sass.splitlines()
for i, ln in enumerate(lines):
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").split("(")[0]
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    mm = re.search(r"/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if not mm:
        continue
    op = mm.group(2)
    counts[kern]["_total"] += 1
    for o in ops:
        if op == o or op.startswith(o + "."):
            counts[kern][o] += 1
    if op.startswith("UTCHMMA") and kern not in excerpts and any(k in kern for k in ("gemm_bf16_kernel<2, false, false>", "attn_fwd2_kernel<128, true>", "attn_bwd_kernel<128, true>")):
        excerpts[kern] = [x.strip()[:150] for x in lines[max(0, i - 6):i + 7]]
out = ["# SASS summary of merlin_b200/libmerlin_b200.so (cuobjdump -sass, sm_100a) — scripts/sass_summary.py",
       "# columns: instructions | " + " | ".join(ops), ""]
tot = collections.Counter()
for k, c in counts.items():
    if c["_total"] == 0:
        continue
    out.append(f"{k[:110]:110s} {c['_total']:7d} | " + " | ".join(f"{c[o]:4d}" for o in ops))
    tot.update(c)
out += ["", "TOTAL".ljust(110) + f" {tot['_total']:7d} | " + " | ".join(f"{tot[o]:4d}" for o in ops), ""]
for k, ex in excerpts.items():
    out += [f"## excerpt around the first UTCHMMA of {k}", *ex, ""]
path = os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
open(path, "w").write("\n".join(out))
print(path, len(counts), "kernels;", {o: tot[o] for o in ops if tot[o]})
