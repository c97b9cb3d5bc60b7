"""Generates tests/golden/synth_reads.json: the first reads of the synthetic read generator (oracle/kmer_oracle.c
orc_synth_reads; SURVEY.md 8d R20M / R200M shapes) so that a change of the generator specification is caught on CPU.
Run from the repository root:  python tests/golden/make_synth_reads_fixture.py"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as orc  # noqa: E402

out = {}
for name, seed, profile, first in (("R20M", 43, 3, 0), ("R20M_late", 43, 3, 19_999_000), ("R200M", 45, 5, 0), ("R200M_late", 45, 5, 199_999_000)):
    seq, qual, _ = orc.synth_reads(seed, profile, first, 1000)
    out[name] = {"seed": seed, "profile": profile, "first_read": first, "reads": 1000,
                 "read0": seq[:150].tobytes().decode(), "qual0": qual[:150].tobytes().decode(),
                 "sha256_seq": hashlib.sha256(seq.tobytes()).hexdigest(), "sha256_qual": hashlib.sha256(qual.tobytes()).hexdigest()}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_reads.json"), "w") as f:
    json.dump(out, f, indent=1)
print("written", {k: v["sha256_seq"][:12] for k, v in out.items()})
