#!/usr/bin/env python
"""Executed warp-instructions and stall samples per CUDA source line of one kernel:
   tools/ncu_lines.py rep kernel_regex [keys] [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
keys = float(sys.argv[3]) if len(sys.argv) > 3 else None
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *(["--kernel-id", ":::" + kern[3:]] if kern.startswith("id:") else ["--kernel-name", "regex:" + kern]), "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, hdr, lines = "?", None, []
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("", "Function Name", "Kernel Name"): continue
    try:
        ie = int(r[hdr.index("Instructions Executed")]); sm = int(r[hdr.index("# Samples")]); te = int(r[hdr.index("Thread Instructions Executed")])
    except ValueError: continue
    lines.append((ie, sm, te, fname, r[0], r[1].strip()))
tot_i = sum(l[0] for l in lines) or 1; tot_s = sum(l[1] for l in lines) or 1
print(f"kernel {kern}: {tot_i:.3e} warp-instr" + (f" = {tot_i/keys:.2f}/key" if keys else "") + f", {tot_s} samples")
for ie, sm, te, f, ln, src in sorted(lines, reverse=True)[:N]:
    print(f"{ie/tot_i*100:5.1f}% instr {sm/tot_s*100:5.1f}% stall  lanes {te/max(ie,1):4.1f}" + (f" {ie/keys:6.3f}/key" if keys else "") + f"  {f}:{ln}  {src[:95]}")
