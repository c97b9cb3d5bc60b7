#!/usr/bin/env python
"""Executed warp-instructions by opcode for one kernel in an .ncu-rep: tools/ncu_opmix.py rep kernel_regex [keys]
With `keys` (number of keys the launch processed) the figures are printed per key."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
keys = float(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *(["--kernel-id", ":::" + kern[3:]] if kern.startswith("id:") else ["--kernel-name", "regex:" + kern]), "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
si, ei, ti = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ops = collections.Counter(); thr = collections.Counter()
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    try: e = int(r[ei]); t = int(r[ti])
    except ValueError: continue
    toks = r[si].strip().split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2]) if op.startswith(("ATOMS", "LDS", "STS", "LDG", "STG", "SHFL", "BAR")) else op.split(".")[0]
    ops[op] += e; thr[op] += t
tot = sum(ops.values())
print(f"kernel {kern}: {tot:.3e} warp-instructions" + (f" = {tot/keys:.2f} per key, {sum(thr.values())/keys:.1f} thread-instr per key" if keys else ""))
for op, e in ops.most_common(28):
    print(f"  {op:14s} {e/tot*100:5.1f}%  lanes {thr[op]/max(e,1):5.1f}" + (f"  {e/keys:6.3f}/key" if keys else ""))
