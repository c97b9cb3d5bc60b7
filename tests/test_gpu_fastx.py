"""kmg_count_fastx: FASTA / FASTQ file images parsed ON THE DEVICE, against the oracle's bio-compatible record parser
(oracle/kmer_oracle.c orc_parse_fastx) + counter.  Fixtures of the reference, multi-line FASTA with \\r\\n / trailing blanks /
blank lines / lower case / no final newline, records spanning staging chunks, 4-line FASTQ with qualities, and the inputs the
device parser must refuse (multi-line FASTQ) so that the host splitter takes over."""
import os

import numpy as np
import pytest

import krust_b200 as kb
from krust_b200 import _lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
PART = _lib.KMG_FLAG_FORCE_PARTITIONED


def oracle_counts(data: bytes, is_fastq: bool, k: int, q=None):
    seq, qual, offsets = orc.parse_fastx(data, is_fastq)
    keys, counts, windows = orc.count_batch(k, seq, qual, offsets, q, mode="rolling")
    return keys, counts, windows, len(offsets) - 1, int(offsets[-1])


def device_counts(data: bytes, is_fastq: bool, k: int, q=None, batch_bases=0, flags=0):
    with kb.GpuKmerCounter(k, min_quality=q, batch_bases=batch_bases, flags=flags, parts_log2=4 if flags == PART else 0) as c:
        n_rec = c.count_fastx(data, is_fastq)
        s = c.finalize()
        keys, counts = c.export(1, True)
    return keys, counts, s, n_rec


def same(dev, ora):
    keys, counts, s, n_rec = dev
    okeys, ocounts, owin, orec, obases = ora
    assert len(keys) == len(okeys) and (keys == okeys).all() and (counts == ocounts).all()
    assert s["n_windows"] == owin and n_rec == orec == s["n_records"] and s["n_bases"] == obases


def test_reference_fixtures(golden_dir):
    fx = os.path.join(golden_dir, "fixtures")
    for name, is_fastq in (("simple.fa", False), ("simple.fq", True), ("with_n.fa", False), ("with_n.fq", True), ("soft_masked.fa", False),
                           ("low_quality.fq", True)):
        data = open(os.path.join(fx, name), "rb").read()
        for k in (1, 3, 4, 7):
            for q in ((None, 20) if is_fastq else (None,)):
                same(device_counts(data, is_fastq, k, q), oracle_counts(data, is_fastq, k, q))
    assert device_counts(b"", False, 3)[3] == 0
    with pytest.raises(kb.SequenceParseError):
        device_counts(b"ACGT\n", False, 3)          # no header (bio: "header line must start with '>'")


def _fasta_image(rng, n_rec, width, eol, trailing_blank=False, final_newline=True, blank_lines=False):
    out = []
    for r in range(n_rec):
        n = int(rng.integers(0, 5000))
        seq = bytes(rng.choice(list(b"ACGTacgtNn"), p=[.2, .2, .2, .2, .04, .04, .04, .04, .02, .02], size=n).tolist())
        out.append(b">rec%d some description" % r + eol)
        for i in range(0, n, width):
            out.append(seq[i:i + width] + (b"  \t" if trailing_blank and i % (3 * width) == 0 else b"") + eol)
            if blank_lines and i % (7 * width) == 0:
                out.append(eol)
    data = b"".join(out)
    return data if final_newline else data.rstrip(b"\r\n")


@pytest.mark.parametrize("k", [1, 5, 21, 32])
def test_multiline_fasta_variants_and_chunk_spanning_records(k):
    rng = np.random.default_rng(k)
    for width, eol, tb, fn, bl in ((80, b"\n", False, True, False), (60, b"\r\n", True, True, True), (100, b"\n", True, False, False), (7, b"\n", False, True, True)):
        data = _fasta_image(rng, 40, width, eol, tb, fn, bl)
        ora = oracle_counts(data, False, k)
        same(device_counts(data, False, k), ora)
        # staging chunks of 4 KiB: most records span several chunks; cuts land inside records and right after headers
        same(device_counts(data, False, k, batch_bases=4096, flags=_lib.KMG_FLAG_FORCE_HASH), ora)
        same(device_counts(data, False, k, batch_bases=10_000, flags=PART), ora)
    one = b">single\n" + bytes(rng.choice(list(b"ACGT"), size=300_000).tolist()) + b"\n"   # one long single-line record, cut many times
    with pytest.raises(kb.SequenceParseError):
        device_counts(one, False, k, batch_bases=4096)   # a line longer than a chunk is refused, not miscounted
    same(device_counts(one, False, k), oracle_counts(one, False, k))


def test_fastq_reads_with_qualities_in_chunks():
    seq, qual, off = orc.synth_reads(43, 3, 0, 20_000)
    n = len(off) - 1
    lines = []
    s2, q2 = seq.reshape(n, 150), qual.reshape(n, 150).copy()
    q2[::7, 0] = ord("@")      # Phred 31 as the first quality of every 7th read: lines that LOOK like headers
    for i in range(n):
        lines.append(b"@read%d/1\n" % i + s2[i].tobytes() + b"\n+\n" + q2[i].tobytes() + b"\n")
    data = b"".join(lines)
    assert b"\n@" in data.replace(b"\n@read", b"")      # some quality lines begin with '@': the chunker must not cut there
    for k, q in ((31, 20), (21, None), (15, 2)):
        ora = oracle_counts(data, True, k, q)
        same(device_counts(data, True, k, q), ora)
        same(device_counts(data, True, k, q, batch_bases=100_000), ora)
        same(device_counts(data, True, k, q, batch_bases=70_001, flags=PART), ora)
    crlf = data.replace(b"\n", b"\r\n")
    same(device_counts(crlf, True, 31, 20, batch_bases=150_000), oracle_counts(crlf, True, 31, 20))


def test_multiline_fastq_falls_back_to_the_host_splitter(tmp_path):
    rec = b"@r1\nACGTAC\nGTACGT\n+\nIIIIII\nIIII!!\n@r2\nGATTACA\n+\nIIIIIII\n"
    with pytest.raises(kb.SequenceParseError):
        device_counts(rec, True, 4)
    p = tmp_path / "ml.fq"
    p.write_bytes(rec)
    seq, qual, offsets = orc.parse_fastx(rec, True)
    okeys, ocounts, _ = orc.count_batch(4, seq, qual, offsets, 20)
    got = kb.count_kmers_with_quality(p, 4, "auto", 20)        # api: device parser first, host splitter on a parse error
    assert got == {kb.unpack_to_string(int(a), 4): int(b) for a, b in zip(okeys, ocounts)}
    bad = tmp_path / "bad.fq"
    bad.write_bytes(b"@r1\nACGT\n+\nIII\n")                      # quality shorter than the sequence: an error on both paths
    with pytest.raises(kb.KmeRustError):
        kb.count_kmers(bad, 3)
