"""BASELINE.json configurations at (or near) their named shape, CUDA path through the C ABI against the CPU oracle.

C1 / C2 at full size live in test_gpu_parity.py::test_full_size_100mbp_oracle_and_properties.  Here:
C3 (k=31 FASTQ-shaped reads with N bases, -Q 20, --min-count 2), C4 (k=21, 3.1 Gbp: a key-space sample of the full
table against the oracle), C5 (k=21 skewed reads streamed in many calls: histogram + .kmix, also from shards)."""
import os

import numpy as np
import pytest

import krust_b200 as kb
from krust_b200 import _lib
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
PART = _lib.KMG_FLAG_FORCE_PARTITIONED


def hist_text(vals, freqs) -> bytes:
    return b"".join(b"%d\t%d\n" % (int(a), int(b)) for a, b in zip(vals, freqs))


def test_read_generators_device_equals_oracle():
    """kmg_synth_reads_device and orc_synth_reads are two implementations of one specification (oracle/kmer_oracle.c)."""
    import torch
    dev = torch.device("cuda:0")
    n = 20_000
    with kb.GpuKmerCounter(21) as c:
        for profile, seed, first in ((3, 43, 0), (3, 43, 123_456_789), (5, 45, 0), (5, 45, 199_990_000)):
            d_seq = torch.empty(n * 150, dtype=torch.uint8, device=dev)
            d_qual = torch.empty(n * 150, dtype=torch.uint8, device=dev)
            c.synth_reads_device(seed, profile, first, n, d_seq.data_ptr(), d_qual.data_ptr())
            seq, qual, _ = orc.synth_reads(seed, profile, first, n)
            assert (d_seq.cpu().numpy() == seq).all() and (d_qual.cpu().numpy() == qual).all()
    seq, qual, _ = orc.synth_reads(43, 3, 0, n)
    assert 0.004 < (seq == ord("N")).mean() < 0.008 and 0.12 < (qual < 53).mean() < 0.16   # N rate, bases below Q20
    seq5, _, _ = orc.synth_reads(45, 5, 0, n, False)
    polya = (seq5.reshape(n, 150) == ord("A")).all(axis=1).mean()
    assert 0.0005 < polya < 0.004   # ~1/640 of the reads are poly-A (before substitutions knock some out)


@pytest.mark.parametrize("flags", [0, PART], ids=["auto", "part"])
def test_c3_shape_reads_q20_min_count_2(flags):
    """C3: k=31, 150 bp reads sampled from G100 with substitutions, N bases and N runs, the 4-class quality mix with the
    3' bias, -Q 20, --min-count 2 -- 2 M reads (300 Mbp + 300 MB of qualities) through kmg_count_ascii."""
    n_reads, k = 2_000_000, 31
    seq, qual, offsets = orc.synth_reads(43, 3, 0, n_reads)
    okeys, ocounts, owin = orc.count_batch_mt(k, seq, qual, offsets, 20)
    with kb.GpuKmerCounter(k, min_quality=20, flags=flags) as c:   # no size hint: planned from the countable windows of the data
        c.count_batch(seq, qual, offsets)
        s = c.finalize()
        assert s["n_windows"] == owin and s["n_distinct"] == len(okeys) and s["n_records"] == n_reads
        assert s["path"] == (2 if flags else 0)   # ~6 M windows survive -Q 20: a small job for the single table unless forced
        if flags:
            assert s["table_capacity"] <= 4 * (owin // 3600 + 4)   # partitions planned for what survives, not for 300 M bases
        for m in (2, 1):
            gk, gc = c.export(m, True)
            fk, fc = orc.filter_min_count(okeys, ocounts, m)
            assert len(gk) == len(fk) and (gk == fk).all() and (gc == fc).all()
            assert hist_text(*c.histogram(m)) == hist_text(*orc.histogram(ocounts, m))
    # FASTA-style feed of the same reads (no qualities) ignores -Q (tests/quality_tests.rs:88-114)
    okeys2, ocounts2, owin2 = orc.count_batch_mt(k, seq, None, offsets)
    with kb.GpuKmerCounter(k, min_quality=20, flags=flags) as c:
        c.count_batch(seq, None, offsets)
        s = c.finalize()
        assert s["n_windows"] == owin2 and owin2 > 20 * owin
        gk, gc = c.export(2, True)
        fk, fc = orc.filter_min_count(okeys2, ocounts2, 2)
        assert len(gk) == len(fk) and (gk == fk).all() and (gc == fc).all()


def test_c5_shape_skewed_reads_streamed_histogram_and_kmix(tmp_path):
    """C5: k=21, skewed 150 bp reads (10 % from a satellite set incl. poly-A / poly-AC), streamed in 48 calls so that every
    call leaves a run and the runs are consolidated while streaming (result + runs merge, LSM style); --format histogram and
    --save .kmix (all k-mers, unfiltered: src/main.rs:155-212) against the oracle."""
    n_reads, k, calls = 1_200_000, 21, 48
    seq, _, offsets = orc.synth_reads(45, 5, 0, n_reads, False)
    okeys, ocounts, owin = orc.count_batch_mt(k, seq, None, offsets)
    per = n_reads // calls
    with kb.GpuKmerCounter(k, flags=PART, expected_distinct=owin) as c:
        for j in range(calls):
            lo, hi = j * per, (j + 1) * per
            c.count_batch(seq[lo * 150:hi * 150], None, offsets[lo:hi + 1] - offsets[lo])
        s = c.finalize()
        assert s["path"] == 2 and s["n_grows"] >= 2          # consolidated at least once while streaming
        assert s["n_windows"] == owin == n_reads * 130 and s["n_distinct"] == len(okeys) and s["max_count"] == int(ocounts.max())
        assert s["max_count"] > 100_000                      # the poly-A key
        assert hist_text(*c.histogram(1)) == hist_text(*orc.histogram(ocounts, 1))
        assert hist_text(*c.histogram(3)) == hist_text(*orc.histogram(ocounts, 3))
        gk, gc = c.export(1, True)
        assert len(gk) == len(okeys) and (gk == okeys).all() and (gc == ocounts).all()
        p = tmp_path / "c5.kmix"
        c.save_kmix(p)
        with pytest.raises(kb.GpuError):
            c.save_kmix(tmp_path / "c5.kmix.gz")             # no silent raw bytes under a .gz name
        # the same index written as 3 "shards" of one context: offsets / CRC combination (kmg_kmix_begin/_shard/_finish)
        p3 = tmp_path / "c5_one_shard.kmix"
        kb.kmix_begin(p3)
        n1, crc1 = c.save_kmix_shard(p3, 0)
        kb.kmix_finish(p3, k, [n1], [crc1])
    blob = p.read_bytes()
    assert blob == orc.kmix_encode(k, okeys, ocounts)        # sorted writer: byte-identical to the sorted reference encoding
    assert p3.read_bytes() == blob
    kk, ikeys, icounts = orc.kmix_decode(blob)
    assert kk == k and (ikeys == okeys).all() and (icounts == ocounts).all()


def test_kmix_from_two_shards_on_one_device(tmp_path):
    """Two contexts hold disjoint halves of the key space (as two GPUs would): the sharded writer must produce ONE valid
    .kmix whose record set is the union (parity definition (v): header, n, record set, valid CRC)."""
    import torch
    dev = torch.device("cuda:0")
    k = 21
    seq, _, offsets = orc.synth_reads(45, 5, 0, 40_000, False)
    okeys, ocounts, _ = orc.count_batch_mt(k, seq, None, offsets)
    owner = np.array([kb.owner_of(int(x), 2) for x in okeys[:2000]])
    assert 0 < owner.sum() < 2000
    # split the table by the engine's own owner function: extract into 2 owner buckets, insert each into its own context
    d_seq = torch.from_numpy(seq).to(dev)
    d_off = torch.from_numpy(offsets.astype(np.int64)).to(dev)
    out = torch.empty(len(seq), dtype=torch.int64, device=dev)
    shards = [kb.GpuKmerCounter(k, flags=PART, parts_log2=10) for _ in range(2)]
    try:
        counts = shards[0].extract_keys_device(d_seq.data_ptr(), len(seq), 2, out.data_ptr(), len(seq), d_offsets=d_off.data_ptr(),
                                               n_records=len(offsets) - 1)
        shards[0].reset()
        lo = 0
        p = tmp_path / "two.kmix"
        kb.kmix_begin(p)
        recs, crcs, off = [], [], 0
        for r in range(2):
            shards[r].insert_keys_device(out[lo:lo + int(counts[r])].data_ptr(), int(counts[r]))
            lo += int(counts[r])
            shards[r].finalize()
            n, crc = shards[r].save_kmix_shard(p, off)
            recs.append(n); crcs.append(crc); off += n
        kb.kmix_finish(p, k, recs, crcs)
    finally:
        for c in shards:
            c.close()
    kk, ikeys, icounts = orc.kmix_decode(p.read_bytes())     # validates size, magic, CRC, version, k, n*16
    assert kk == k and len(ikeys) == len(okeys) == sum(recs)
    order = np.argsort(ikeys, kind="stable")
    assert (ikeys[order] == okeys).all() and (icounts[order] == ocounts).all()
    assert kb.load_index(p).counts() == dict(zip(okeys.tolist(), ocounts.tolist()))


def test_c4_full_size_sampled_against_oracle():
    """C4 at its named size: k=21, 3.1 Gbp uniform genome (31 records x 100 Mbp, seed 44), counted by the headline pipeline
    (P1 = P2 = 928, speculative layouts).  The 50 GB table cannot be compared wholesale on this host, so two design-independent
    samples of the key space (key % 1009 == r, ~3.07 M keys each) are exported with kmg_export_shard and compared with the
    oracle counting exactly those keys over the whole stream; plus window / distinct totals and the histogram."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 120e9:
        pytest.skip("needs a 180 GB class device")
    import psutil
    if psutil.virtual_memory().available < 24e9:
        pytest.skip("the oracle needs ~8 GB of host memory for the 3.1 Gbp stream and its keys")
    dev = torch.device("cuda:0")
    n, n_rec, k = 3_100_000_000, 31, 21
    buf = torch.empty(n + 64, dtype=torch.uint8, device=dev)[:n]
    offsets_np = np.arange(0, n + 1, n // n_rec, dtype=np.uint64)
    offsets = torch.from_numpy(offsets_np.astype(np.int64)).to(dev)
    host = orc.synth_uniform(44, 0, n)
    want = {}
    owin = 0
    for rem in (0, 517):
        ok_, oc_, owin = orc.count_batch_mt(k, host, None, offsets_np, filter_mod=1009, filter_rem=rem)
        want[rem] = (ok_, oc_)
    with kb.GpuKmerCounter(k, expected_distinct=n) as c:
        c.synth_uniform_device(44, 0, n, buf.data_ptr())
        probe = buf[10_000_000:10_001_000].cpu().numpy()
        assert (probe == host[10_000_000:10_001_000]).all()
        del host
        c.count_device(buf.data_ptr(), n, d_offsets=offsets.data_ptr(), n_records=n_rec)
        s = c.finalize()
        assert s["path"] == 2 and s["n_windows"] == owin == n - n_rec * (k - 1)
        total_distinct = 0
        for rem, (ok_, oc_) in want.items():
            gk, gc = c.export_shard(1009, rem, 1, True)
            assert len(gk) == len(ok_) and (gk == ok_).all() and (gc == oc_).all()
        hv, hf = c.histogram(1)
        assert int((hv * hf).sum()) == s["n_windows"] and int(hf.sum()) == s["n_distinct"]
        # the sample's share of the count-of-counts is consistent with the whole (1/1009 of the keys, binomial 6 sigma)
        for rem, (ok_, oc_) in want.items():
            exp1 = float(hf[hv == 1][0]) / 1009
            got1 = float((oc_ == 1).sum())
            assert abs(got1 - exp1) < 6 * exp1 ** 0.5 + 10


def _oracle_text(keys, counts, k, fmt) -> bytes:
    names = kb.unpack_many(keys, k).tolist()
    if fmt == "tsv":
        return b"".join(b"%s\t%d\n" % (s, int(c)) for s, c in zip(names, counts.tolist()))
    return b"".join(b">%d\n%s\n" % (int(c), s) for s, c in zip(names, counts.tolist()))


def test_device_text_emitters_small_cases(tmp_path):
    """kmg_emit_text / kmg_write_text: fasta and tsv lines formatted on the device equal the reference's line formats
    (src/run.rs:452-470) over the sorted oracle table -- all k, multi-digit counts, min-count, every engine path."""
    import io
    rng = np.random.default_rng(9)
    recs = [b"A" * 5000, b"AC" * 700, b"ACGTTGCA" * 150] + [bytes(rng.choice(list(b"ACGTN"), size=int(rng.integers(0, 900))).tolist()) for _ in range(60)]
    for k, flags in ((1, 0), (3, 0), (12, 0), (21, 0), (21, PART), (32, _lib.KMG_FLAG_FORCE_HASH), (31, PART)):
        okeys, ocounts, _ = orc.count_records(k, recs, mode="rolling")
        with kb.GpuKmerCounter(k, flags=flags, parts_log2=3 if flags == PART else 0) as c:
            c.count_records(recs)
            c.finalize(False)
            for fmt in ("tsv", "fasta"):
                for m in (0, 1, 2, 50):
                    out = io.BytesIO()
                    n_rec, n_bytes = c.emit_text(out, fmt, m)
                    fk, fc = orc.filter_min_count(okeys, ocounts, max(m, 1))
                    want = _oracle_text(fk, fc, k, fmt)
                    assert out.getvalue() == want and n_rec == len(fk) and n_bytes == len(want)
            p = tmp_path / "o.tsv"
            assert c.write_text(p, "tsv", 1)[0] == len(okeys) and p.read_bytes() == _oracle_text(okeys, ocounts, k, "tsv")
    # the reference's own exact-value line tests: soft_masked.fa k=3 -> "AAA\t2" (tests/integration_tests.rs:263-281)
    out = io.BytesIO()
    kb.KmerCounter.new().k(3).format("tsv").count_to_writer(os.path.join(os.path.dirname(__file__), "golden", "fixtures", "soft_masked.fa"), out)
    assert out.getvalue() == b"AAA\t2\n"
    out = io.BytesIO()
    kb.KmerCounter.new().k(3).format("fasta").count_to_writer(os.path.join(os.path.dirname(__file__), "golden", "fixtures", "soft_masked.fa"), out)
    assert out.getvalue() == b">2\nAAA\n"


def test_c1_tsv_text_100mbp_equals_sorted_oracle_text(tmp_path):
    """C1 end to end at its named size: kmerust 21 on the 100 Mbp genome, --format tsv.  The 2.4 GB of text produced on the device
    equal the sorted oracle table formatted line by line (parity definition (iv): the reference's output order is random)."""
    import torch
    dev = torch.device("cuda:0")
    n, n_rec, k = 100_000_000, 100, 21
    host = orc.synth_uniform(42, 0, n)
    offs = np.arange(0, n + 1, n // n_rec, dtype=np.uint64)
    okeys, ocounts, _ = orc.count_batch_mt(k, host, None, offs)
    assert int(ocounts.max()) < 10   # one-digit counts: every tsv line is k + 3 bytes, the oracle text can be built as a matrix
    p = tmp_path / "c1.tsv"
    with kb.GpuKmerCounter(k, expected_distinct=n) as c:
        c.count_batch(host, None, offs)
        c.finalize(False)
        n_out, n_bytes = c.write_text(p, "tsv", 1)
    assert n_out == len(okeys) and n_bytes == len(okeys) * (k + 3) == os.path.getsize(p)
    got = np.memmap(p, dtype=np.uint8, mode="r").reshape(-1, k + 3)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    shifts = (np.arange(k - 1, -1, -1, dtype=np.uint64) * np.uint64(2))
    step = 4_000_000
    for i in range(0, len(okeys), step):
        kk = okeys[i:i + step]
        want = np.empty((len(kk), k + 3), dtype=np.uint8)
        want[:, :k] = lut[((kk[:, None] >> shifts[None, :]) & np.uint64(3)).astype(np.uint8)]
        want[:, k] = ord("\t"); want[:, k + 1] = ocounts[i:i + step].astype(np.uint8) + ord("0"); want[:, k + 2] = ord("\n")
        assert (got[i:i + step] == want).all()


def test_index_round_trip_and_batched_queries_on_the_device(tmp_path):
    """save -> kmg_index_open -> kmg_query_*: the `query` subcommand's semantics (src/main.rs:233-281: upper-case, canonicalise,
    look up; absent -> 0) batched on the device, for every table kind; index round trips like src/index.rs:510-524."""
    rng = np.random.default_rng(3)
    recs = [bytes(rng.choice(list(b"ACGTN"), p=[.245, .245, .245, .245, .02], size=int(rng.integers(50, 4000))).tolist()) for _ in range(200)]
    recs.append(b"A" * 3000)
    for k, flags in ((1, 0), (5, 0), (16, _lib.KMG_FLAG_FORCE_HASH), (21, PART), (32, PART), (21, 0)):
        okeys, ocounts, _ = orc.count_records(k, recs, mode="rolling")
        table = dict(zip(okeys.tolist(), ocounts.tolist()))
        p = tmp_path / f"q{k}.kmix"
        # queries: present keys, absent keys, k-mer strings in both orientations and cases, an invalid one
        absent = np.array([x for x in rng.integers(0, 4 ** k, size=300, dtype=np.uint64).tolist() if x not in table][:100] or [0], dtype=np.uint64)
        qk = np.concatenate([okeys[:: max(1, len(okeys) // 500)], absent])
        want = np.array([table.get(int(x), 0) for x in qk], dtype=np.uint64)
        strs = [kb.unpack_to_string(int(x), k).encode() for x in okeys[:50]]
        comp = bytes.maketrans(b"ACGT", b"TGCA")
        strs += [s.translate(comp)[::-1].lower() for s in strs[:25]] + [b"N" * k]
        want_s = [table[int(x)] for x in okeys[:50]] + [table[int(x)] for x in okeys[:25]] + [0]
        with kb.GpuKmerCounter(k, flags=flags, parts_log2=6 if flags == PART else 0) as c:
            c.count_records(recs)
            c.finalize(False)
            assert (c.query_keys(qk) == want).all()
            got_s, bad = c.query_kmers(strs)
            assert got_s.tolist() == want_s and bad == 1
            c.save_kmix(p)
        with kb.GpuKmerCounter.open_index(p) as idx:
            assert idx.k == k
            s = idx.finalize()
            assert s["n_distinct"] == len(okeys) and s["n_windows"] == int(ocounts.sum())
            assert (idx.query_keys(qk) == want).all()
            got_s, bad = idx.query_kmers(strs)
            assert got_s.tolist() == want_s and bad == 1
            gk, gc = idx.export(1, True)
            assert (gk == okeys).all() and (gc == ocounts).all()
            p2 = tmp_path / "again.kmix"
            idx.save_kmix(p2)
        assert p2.read_bytes() == p.read_bytes() == orc.kmix_encode(k, okeys, ocounts)
    empty = tmp_path / "empty.kmix"
    empty.write_bytes(orc.kmix_encode(7, np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64)))
    with kb.GpuKmerCounter.open_index(empty) as idx:
        assert idx.k == 7 and idx.query_keys(np.array([5], dtype=np.uint64)).tolist() == [0]
