#!/usr/bin/env python
"""Top stall-sample SASS lines of one kernel in an .ncu-rep: tools/ncu_hot.py rep kernel_regex [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# possibly several kernels: take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
si, ci = hdr.index("Source"), hdr.index("# Samples")
ei = hdr.index("Instructions Executed")
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    try: body.append((int(r[ci]), int(r[ei]), r[si].strip()))
    except ValueError: pass
tot = sum(b[0] for b in body) or 1
print(f"kernel {kern}: {len(body)} SASS lines, {tot} samples")
for idx, (c, e, s) in enumerate(body):
    body[idx] = (c, e, s, idx)
for c, e, s, idx in sorted(body, reverse=True)[:N]:
    print(f"{c/tot*100:5.1f}%  line {idx:4d} exec {e:10d}  {s[:100]}")
