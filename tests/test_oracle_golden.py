"""Pins the CPU oracle against the reference's own known-answer tests and
cross-checks oracle #1 (literal restatement) against oracle #2 (rolling form)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import kats


@pytest.mark.parametrize("kat", kats.COUNT_KATS, ids=[k[0] for k in kats.COUNT_KATS])
@pytest.mark.parametrize("mode", ["literal", "rolling"])
def test_count_kats(kat, mode):
    _, _, k, records, quals, q, expected = kat
    got = orc.count_dict(k, records, quals, q, mode=mode)
    assert got == expected


@pytest.mark.parametrize("seq,canon,is_rc", kats.CANONICAL_KATS)
def test_canonical_kats(seq, canon, is_rc):
    bits, rc = orc.canonical(seq)
    assert orc.unpack(bits, len(seq)).decode() == canon
    assert rc == is_rc


@pytest.mark.parametrize("seq,pos", kats.INVALID_BASE_KATS)
def test_invalid_base_positions(seq, pos):
    norm, err = orc.from_sub(seq)
    assert norm is None and err == (ord("N"), pos)


def test_pack_and_unpack():
    for s, bits in kats.PACK_KATS:
        assert orc.pack(s) == bits
        assert orc.unpack(bits, len(s)) == s
    # kmer.rs:733-760 round trips for k = 1..32
    rng = np.random.default_rng(1)
    for k in range(1, 33):
        s = bytes(rng.choice(list(b"ACGT"), size=k).tolist())
        assert orc.unpack(orc.pack(s), k) == s
    # kmer.rs:764-786 soft-masked input is upper-cased
    assert orc.from_sub(b"gattaca")[0] == b"GATTACA"


def test_kmer_length_bounds():
    # kmer.rs:100-111; tests/library_tests.rs:155-170
    assert not orc.kmer_length_ok(0) and not orc.kmer_length_ok(33)
    assert orc.kmer_length_ok(1) and orc.kmer_length_ok(32)
    with pytest.raises(ValueError):
        orc.count_records(0, [b"ACGT"])
    with pytest.raises(ValueError):
        orc.count_records(33, [b"ACGT"])


def test_gattaca_equals_reverse_complement():
    # tests/library_tests.rs:220-230
    assert orc.count_dict(3, [b"GATTACA"]) == orc.count_dict(3, [b"TGTAATC"])


def test_kmer_plus_rc_is_one_entry_count_two():
    # tests/property_tests.rs:288-330, all k <= 32
    rng = np.random.default_rng(7)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    for k in range(1, 33):
        s = bytes(rng.choice(list(b"ACGT"), size=k).tolist())
        rc = bytes(comp[b] for b in reversed(s))
        keys, counts, windows = orc.count_records(k, [s, rc])
        assert len(keys) == 1 and counts[0] == 2 and windows == 2


def test_canonical_properties():
    # tests/property_tests.rs:64-126: idempotent, RC-invariant, lexicographically smallest
    rng = np.random.default_rng(11)
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    for _ in range(500):
        k = int(rng.integers(1, 33))
        s = bytes(rng.choice(list(b"ACGT"), size=k).tolist())
        rc = bytes(comp[b] for b in reversed(s))
        bits, _ = orc.canonical(s)
        canon = orc.unpack(bits, k)
        assert canon == min(s, rc)
        assert orc.canonical(rc)[0] == bits
        assert orc.canonical(canon)[0] == bits
        # case-insensitive (property_tests.rs:148-177)
        assert orc.canonical(s.lower())[0] == bits


def test_histogram_kats():
    # histogram.rs:176-287: {1,1,2,2} -> {1:2, 2:2}; keys ascending; sum(freq) = #distinct
    vals, freqs = orc.histogram(np.array([1, 1, 2, 2], dtype=np.uint64))
    assert vals.tolist() == [1, 2] and freqs.tolist() == [2, 2]
    vals, freqs = orc.histogram(np.array([100, 1, 50, 1, 7], dtype=np.uint64))
    assert vals.tolist() == [1, 7, 50, 100] and freqs.tolist() == [2, 1, 1, 1]
    assert orc.histogram(np.array([], dtype=np.uint64))[0].tolist() == []
    # integration_tests.rs:767-799: AAAAAAAA k=3 -> "6\t1"
    keys, counts, _ = orc.count_records(3, [b"AAAAAAAA"])
    vals, freqs = orc.histogram(counts)
    assert list(zip(vals.tolist(), freqs.tolist())) == [(6, 1)]
    # simple.fa k=3 derived: 1:4, 3:1, 4:1
    keys, counts, windows = orc.count_records(3, [b"ACGTACGT", b"GATTACA"])
    vals, freqs = orc.histogram(counts)
    assert list(zip(vals.tolist(), freqs.tolist())) == [(1, 4), (3, 1), (4, 1)] and windows == 11
    # min-count applies before the histogram (run.rs:447-476; integration_tests.rs:708-733)
    vals, freqs = orc.histogram(counts, min_count=2)
    assert list(zip(vals.tolist(), freqs.tolist())) == [(3, 1), (4, 1)]
    st = orc.histogram_stats(np.array([1, 2], dtype=np.uint64), np.array([2, 2], dtype=np.uint64))
    assert st["distinct_kmers"] == 4 and st["total_kmers"] == 6 and st["mean_count"] == 1.5
    st = orc.histogram_stats(np.array([42], dtype=np.uint64), np.array([1], dtype=np.uint64))
    assert st["mean_count"] == 42.0 and st["mode_count"] == 42


def test_crc32_kats():
    for data, crc in kats.CRC_KATS:
        assert orc.crc32(data) == crc
    import zlib
    blob = os.urandom(4097)
    assert orc.crc32(blob) == zlib.crc32(blob)


def test_kmix_round_trip_and_rejects():
    # index.rs:510-524 round trips for k in {1,5,16,21,32}; property_tests.rs:244-261 arbitrary u64 keys
    rng = np.random.default_rng(3)
    for k in (1, 5, 16, 21, 32):
        keys = rng.integers(0, 2**63, size=50, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
        counts = rng.integers(1, 2**62, size=50, dtype=np.uint64)
        blob = orc.kmix_encode(k, keys, counts)
        assert len(blob) == 18 + 16 * 50 and blob[:4] == b"KMIX" and blob[4] == 1 and blob[5] == k
        k2, keys2, counts2 = orc.kmix_decode(blob)
        assert k2 == k and (keys2 == keys).all() and (counts2 == counts).all()
    empty = orc.kmix_encode(21, np.array([], dtype=np.uint64), np.array([], dtype=np.uint64))
    assert len(empty) == 18 and orc.kmix_decode(empty)[1].size == 0
    # index.rs:527-573 rejects: < 18 bytes, bad magic, flipped byte
    with pytest.raises(ValueError, match="too small"):
        orc.kmix_decode(b"KMIX")
    with pytest.raises(ValueError, match="magic"):
        orc.kmix_decode(b"XXXX" + blob[4:])
    bad = bytearray(blob); bad[20] ^= 0xFF
    with pytest.raises(ValueError, match="checksum"):
        orc.kmix_decode(bytes(bad))


def test_parser_fixtures(golden_dir):
    fx = os.path.join(golden_dir, "fixtures")
    seq, qual, off = orc.parse_fastx(open(os.path.join(fx, "simple.fa"), "rb").read(), False)
    assert seq.tobytes() == b"ACGTACGTGATTACA" and off.tolist() == [0, 8, 15] and qual is None
    seq, qual, off = orc.parse_fastx(open(os.path.join(fx, "low_quality.fq"), "rb").read(), True)
    assert seq.tobytes() == b"ACGTACGTGATTACA" and qual.tobytes() == b"IIII!!!!IIIIIII" and off.tolist() == [0, 8, 15]
    # FASTA == FASTQ counts on the paired fixtures (integration_tests.rs:486-523)
    for stem in ("simple", "with_n"):
        a = orc.parse_fastx(open(os.path.join(fx, stem + ".fa"), "rb").read(), False)
        q = orc.parse_fastx(open(os.path.join(fx, stem + ".fq"), "rb").read(), True)
        ka, ca, _ = orc.count_batch(3, a[0], None, a[2])
        kq, cq, _ = orc.count_batch(3, q[0], q[1], q[2])
        assert (ka == kq).all() and (ca == cq).all()
    # multi-line FASTA accepted (library_tests.rs:233-241), CRLF trimmed, empty file / header only -> no k-mers
    seq, _, off = orc.parse_fastx(b">s desc\nACGT\r\nACGT\n>t\nGG\n", False)
    assert seq.tobytes() == b"ACGTACGTGG" and off.tolist() == [0, 8, 10]
    assert orc.parse_fastx(b"", False)[2].tolist() == [0]
    assert orc.parse_fastx(b">only\n", False)[2].tolist() == [0, 0]
    with pytest.raises(ValueError):
        orc.parse_fastx(b"ACGT\n", False)


def _random_records(rng, n_rec, max_len, alphabet):
    recs, quals = [], []
    for _ in range(n_rec):
        n = int(rng.integers(0, max_len))
        recs.append(bytes(rng.choice(alphabet, size=n).tolist()))
        quals.append(bytes((rng.integers(0, 42, size=n) + 33).astype(np.uint8).tolist()))
    return recs, quals


def test_literal_equals_rolling_randomised():
    """oracle #1 == oracle #2 on random inputs incl. N / IUPAC / blanks / lower case and quality
    thresholds {None, 0, 19, 20, 30, 93, 250}, k = 1..32."""
    rng = np.random.default_rng(2024)
    alphabet = list(b"ACGT" * 12 + b"acgtNnRY -")
    for trial in range(300):
        k = int(rng.integers(1, 33))
        recs, quals = _random_records(rng, int(rng.integers(1, 6)), 120, alphabet)
        q = [None, 0, 19, 20, 30, 93, 250][trial % 7]
        use_qual = trial % 3 != 0
        a = orc.count_records(k, recs, quals if use_qual else None, q, mode="literal")
        b = orc.count_records(k, recs, quals if use_qual else None, q, mode="rolling")
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2]
        assert int(a[1].sum()) == a[2]  # sum of counts == counted windows
        assert a[2] <= sum(max(0, len(r) - k + 1) for r in recs)  # property_tests.rs:263-286


def test_reference_path_threads_equals_oracle():
    rng = np.random.default_rng(5)
    recs, quals = _random_records(rng, 200, 400, list(b"ACGT" * 20 + b"N"))
    seq, qual, off = orc.make_batch(recs, quals)
    for k, q in ((21, None), (5, 20), (32, 10)):
        keys, counts, windows = orc.count_batch(k, seq, qual, off, q, mode="literal")
        w, d, k2, c2 = orc.reference_path_count(k, seq, qual, off, q, threads=4, export=True)
        assert w == windows and d == len(keys) and (k2 == keys).all() and (c2 == counts).all()


def test_synth_uniform_is_counter_based():
    a = orc.synth_uniform(42, 0, 1000)
    b = orc.synth_uniform(42, 500, 500)
    assert (a[500:] == b).all() and set(a.tolist()) <= set(b"ACGT")


def test_multithreaded_oracle_equals_literal_and_rolling():
    """orc_count_batch_mt (used for the BASELINE-sized comparisons) against oracle #1 and #2, incl. the key-space filter."""
    for k in (1, 2, 5, 12, 21, 31, 32):
        seq, qual, off = orc.synth_reads(43, 3, 7, 1500)
        lit = orc.count_batch(k, seq, qual, off, 20, mode="literal")
        rol = orc.count_batch(k, seq, qual, off, 20, mode="rolling")
        for threads in (1, 3):
            mt = orc.count_batch_mt(k, seq, qual, off, 20, threads=threads)
            assert (mt[0] == lit[0]).all() and (mt[1] == lit[1]).all() and mt[2] == lit[2] == rol[2]
        flt = orc.count_batch_mt(k, seq, qual, off, 20, threads=2, filter_mod=5, filter_rem=2)
        m = lit[0] % np.uint64(5) == 2
        assert (flt[0] == lit[0][m]).all() and (flt[1] == lit[1][m]).all() and flt[2] == lit[2]
    # empty input, records shorter than k
    e = orc.count_batch_mt(21, np.zeros(0, dtype=np.uint8), None, np.zeros(1, dtype=np.uint64))
    assert len(e[0]) == 0 and e[2] == 0
    e = orc.count_batch_mt(21, np.frombuffer(b"ACGTACGT", dtype=np.uint8), None, np.array([0, 4, 8], dtype=np.uint64))
    assert len(e[0]) == 0 and e[2] == 0


def test_synthetic_read_generator_matches_committed_fixture(golden_dir):
    import hashlib
    import json
    fx = json.load(open(os.path.join(golden_dir, "synth_reads.json")))
    for name, f in fx.items():
        seq, qual, off = orc.synth_reads(f["seed"], f["profile"], f["first_read"], f["reads"])
        assert seq[:150].tobytes().decode() == f["read0"] and qual[:150].tobytes().decode() == f["qual0"], name
        assert hashlib.sha256(seq.tobytes()).hexdigest() == f["sha256_seq"], name
        assert hashlib.sha256(qual.tobytes()).hexdigest() == f["sha256_qual"], name
        assert len(off) == f["reads"] + 1 and int(off[-1]) == len(seq)
    # shape of the R20M / R200M specifications
    seq, qual, _ = orc.synth_reads(43, 3, 0, 20000)
    assert 0.004 < (seq == ord("N")).mean() < 0.008
    q = qual.reshape(-1, 150)
    assert 0.08 < (q[:, :120] < 53).mean() < 0.12 and 0.27 < (q[:, 120:] < 53).mean() < 0.33
    seq5, qual5, _ = orc.synth_reads(45, 5, 0, 20000)
    assert (qual5 == ord("I")).all() and (seq5 != ord("N")).all()
