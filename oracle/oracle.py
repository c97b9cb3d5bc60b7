"""ctypes wrapper around oracle/libkmer_oracle.so (the CPU restatement of kmerust's counting path).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package krust_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkmer_oracle.so")
_lib = None

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no GPU, no reference sources involved)."""
    src = os.path.join(_HERE, "kmer_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libkmer_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_kmer_length_ok.argtypes = [C.c_uint64]; L.orc_kmer_length_ok.restype = C.c_int
        L.orc_from_sub.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, u8p]; L.orc_from_sub.restype = C.c_int64
        L.orc_pack_bytes.argtypes = [C.c_char_p, C.c_uint64]; L.orc_pack_bytes.restype = C.c_uint64
        L.orc_canonical.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_int)]; L.orc_canonical.restype = C.c_uint64
        L.orc_unpack.argtypes = [C.c_uint64, C.c_uint64, C.c_char_p]; L.orc_unpack.restype = None
        L.orc_counter_new.argtypes = [C.c_uint32, C.c_int, C.c_uint8, C.c_int]; L.orc_counter_new.restype = C.c_void_p
        L.orc_counter_free.argtypes = [C.c_void_p]; L.orc_counter_free.restype = None
        L.orc_counter_add.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p]; L.orc_counter_add.restype = None
        L.orc_counter_add_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_counter_add_batch.restype = None
        for f in ("windows", "records", "bases", "distinct"):
            getattr(L, "orc_counter_" + f).argtypes = [C.c_void_p]
            getattr(L, "orc_counter_" + f).restype = C.c_uint64
        L.orc_counter_export_sorted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_counter_export_sorted.restype = C.c_uint64
        L.orc_filter_min_count.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_filter_min_count.restype = C.c_uint64
        L.orc_histogram.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_histogram.restype = C.c_uint64
        L.orc_histogram_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, u64p, u64p, u64p, u64p, C.POINTER(C.c_double)]
        L.orc_histogram_stats.restype = None
        L.orc_crc32.argtypes = [C.c_char_p, C.c_uint64]; L.orc_crc32.restype = C.c_uint32
        L.orc_kmix_encode.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_kmix_encode.restype = C.c_uint64
        L.orc_kmix_decode.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_kmix_decode.restype = C.c_int64
        L.orc_parse_fastx.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_parse_fastx.restype = C.c_int64
        L.orc_reference_path_count.argtypes = [C.c_uint32, C.c_int, C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_uint64, C.c_uint32, u64p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_reference_path_count.restype = C.c_uint64
        L.orc_synth_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_synth_uniform.restype = None
        L.orc_synth_reads.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_synth_reads.restype = None
        L.orc_free.argtypes = [C.c_void_p]; L.orc_free.restype = None
        L.orc_count_batch_mt.argtypes = [C.c_uint32, C.c_int, C.c_uint8, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                         C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), u64p]
        L.orc_count_batch_mt.restype = C.c_uint64
        _lib = L
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------- primitives
def kmer_length_ok(k: int) -> bool:
    return bool(lib().orc_kmer_length_ok(k))


def from_sub(sub: bytes) -> Tuple[Optional[bytes], Optional[Tuple[int, int]]]:
    """Returns (normalised, None) or (None, (base, position))."""
    out = C.create_string_buffer(len(sub) + 1)
    bad = C.c_uint8(0)
    pos = lib().orc_from_sub(sub, len(sub), out, C.byref(bad))
    if pos < 0:
        return out.raw[: len(sub)], None
    return None, (bad.value, int(pos))


def pack(kmer: bytes) -> int:
    return int(lib().orc_pack_bytes(kmer, len(kmer)))


def canonical(kmer: bytes) -> Tuple[int, bool]:
    norm, err = from_sub(kmer)
    assert err is None, err
    rc = C.c_int(0)
    bits = lib().orc_canonical(norm, len(norm), C.byref(rc))
    return int(bits), bool(rc.value)


def unpack(bits: int, k: int) -> bytes:
    out = C.create_string_buffer(k + 1)
    lib().orc_unpack(bits, k, out)
    return out.raw[:k]


# ---------------------------------------------------------------------------- batches
def make_batch(records: Sequence[bytes], quals: Optional[Sequence[bytes]] = None):
    """Lay records back to back -> (seq u8[], qual u8[]|None, offsets u64[n+1])."""
    offsets = np.zeros(len(records) + 1, dtype=np.uint64)
    if len(records):
        offsets[1:] = np.cumsum([len(r) for r in records], dtype=np.uint64)
    seq = np.frombuffer(b"".join(records), dtype=np.uint8).copy()
    qual = None
    if quals is not None:
        assert all(len(q) == len(r) for q, r in zip(quals, records))
        qual = np.frombuffer(b"".join(quals), dtype=np.uint8).copy()
    return seq, qual, offsets


def count_batch(k: int, seq: np.ndarray, qual: Optional[np.ndarray], offsets: np.ndarray,
                min_quality: Optional[int] = None, mode: str = "rolling"):
    """Sorted (keys, counts, windows) for a batch; mode 'literal' (oracle #1) or 'rolling' (#2)."""
    L = lib()
    c = L.orc_counter_new(k, int(min_quality is not None), int(min_quality or 0), 0 if mode == "literal" else 1)
    if not c:
        raise ValueError(f"invalid k-mer length {k} (must be 1..=32)")
    try:
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if qual is not None:
            qual = np.ascontiguousarray(qual, dtype=np.uint8)
        L.orc_counter_add_batch(c, _ptr(seq), _ptr(qual), _ptr(offsets), len(offsets) - 1)
        n = L.orc_counter_distinct(c)
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint64)
        m = L.orc_counter_export_sorted(c, _ptr(keys), _ptr(counts))
        assert m == n
        return keys, counts, int(L.orc_counter_windows(c))
    finally:
        L.orc_counter_free(c)


def count_records(k: int, records: Iterable[bytes], quals: Optional[Iterable[bytes]] = None,
                  min_quality: Optional[int] = None, mode: str = "rolling"):
    records = list(records)
    quals = None if quals is None else list(quals)
    seq, qual, offsets = make_batch(records, quals)
    return count_batch(k, seq, qual, offsets, min_quality, mode)


def count_dict(k: int, records, quals=None, min_quality=None, mode="literal"):
    """{kmer string: count} -- reads like the reference's HashMap<String,u64> (run.rs:573-582)."""
    keys, counts, _ = count_records(k, records, quals, min_quality, mode)
    return {unpack(int(a), k).decode(): int(b) for a, b in zip(keys, counts)}


def filter_min_count(keys: np.ndarray, counts: np.ndarray, min_count: int):
    keys = keys.copy(); counts = counts.copy()
    m = lib().orc_filter_min_count(_ptr(keys), _ptr(counts), len(keys), min_count)
    return keys[:m], counts[:m]


def histogram(counts: np.ndarray, min_count: int = 1):
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    vals = np.empty(len(counts) + 1, dtype=np.uint64)
    freqs = np.empty(len(counts) + 1, dtype=np.uint64)
    b = lib().orc_histogram(_ptr(counts), len(counts), min_count, _ptr(vals), _ptr(freqs))
    return vals[:b].copy(), freqs[:b].copy()


def histogram_stats(vals: np.ndarray, freqs: np.ndarray):
    t, d, mc, mf = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    mean = C.c_double()
    vals = np.ascontiguousarray(vals, dtype=np.uint64); freqs = np.ascontiguousarray(freqs, dtype=np.uint64)
    lib().orc_histogram_stats(_ptr(vals), _ptr(freqs), len(vals), C.byref(t), C.byref(d), C.byref(mc), C.byref(mf), C.byref(mean))
    return dict(total_kmers=t.value, distinct_kmers=d.value, mode_count=mc.value, mode_frequency=mf.value, mean_count=mean.value)


def crc32(data: bytes) -> int:
    return int(lib().orc_crc32(data, len(data)))


def kmix_encode(k: int, keys: np.ndarray, counts: np.ndarray) -> bytes:
    keys = np.ascontiguousarray(keys, dtype=np.uint64); counts = np.ascontiguousarray(counts, dtype=np.uint64)
    buf = np.empty(18 + 16 * len(keys), dtype=np.uint8)
    n = lib().orc_kmix_encode(k, _ptr(keys), _ptr(counts), len(keys), _ptr(buf))
    return buf[:n].tobytes()


KMIX_ERRORS = {-1: "file too small", -2: "invalid magic bytes", -3: "checksum mismatch", -4: "unsupported version",
               -5: "invalid k-mer length", -6: "data size mismatch"}


def kmix_decode(data: bytes):
    """Returns (k, keys, counts); raises ValueError(reason) like KmeRustError::InvalidIndex."""
    k = C.c_uint32()
    n = lib().orc_kmix_decode(data, len(data), C.byref(k), None, None, 0)
    if n < 0:
        raise ValueError(KMIX_ERRORS[int(n)])
    keys = np.empty(n, dtype=np.uint64); counts = np.empty(n, dtype=np.uint64)
    lib().orc_kmix_decode(data, len(data), C.byref(k), _ptr(keys), _ptr(counts), n)
    return k.value, keys, counts


def parse_fastx(data: bytes, is_fastq: bool):
    """bio-3.0.0-compatible record parser -> (seq, qual|None, offsets)."""
    seq = np.empty(len(data) + 1, dtype=np.uint8)
    qual = np.empty(len(data) + 1, dtype=np.uint8) if is_fastq else None
    offsets = np.zeros(data.count(b"\n") + 3, dtype=np.uint64)
    n = lib().orc_parse_fastx(data, len(data), int(is_fastq), _ptr(seq), _ptr(qual), _ptr(offsets), len(offsets) - 1)
    if n < 0:
        raise ValueError("sequence parse error")
    offsets = offsets[: n + 1].copy()
    total = int(offsets[-1])
    return seq[:total].copy(), (qual[:total].copy() if is_fastq else None), offsets


def reference_path_count(k: int, seq: np.ndarray, qual: Optional[np.ndarray], offsets: np.ndarray,
                         min_quality: Optional[int] = None, threads: int = 0, export: bool = False):
    """The multi-threaded 'restated reference CPU path' (timed CPU baseline).  Returns
    (windows, distinct[, keys, counts sorted])."""
    threads = threads or (os.cpu_count() or 1)
    d = C.c_uint64()
    seq = np.ascontiguousarray(seq, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    L = lib()
    if not export:
        w = L.orc_reference_path_count(k, int(min_quality is not None), int(min_quality or 0), _ptr(seq), _ptr(qual),
                                       _ptr(offsets), len(offsets) - 1, threads, C.byref(d), None, None, 0)
        return int(w), int(d.value)
    cap = max(1, int(offsets[-1]))
    keys = np.empty(cap, dtype=np.uint64); counts = np.empty(cap, dtype=np.uint64)
    w = L.orc_reference_path_count(k, int(min_quality is not None), int(min_quality or 0), _ptr(seq), _ptr(qual),
                                   _ptr(offsets), len(offsets) - 1, threads, C.byref(d), _ptr(keys), _ptr(counts), cap)
    n = int(d.value)
    order = np.argsort(keys[:n], kind="stable")
    return int(w), n, keys[:n][order], counts[:n][order]


def synth_uniform(seed: int, first_base: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    lib().orc_synth_uniform(seed, first_base, n, _ptr(out))
    return out


READ_LEN = 150  # SYN_READ_LEN of the read generator


def synth_reads(seed: int, profile: int, first_read: int, n_reads: int, want_qual: bool = True):
    """Synthetic reads of SURVEY.md 8d (profile 3 = R20M shape for C3, 5 = R200M shape for C5), laid back to back.
    Returns (seq u8[n*150], qual u8[n*150] | None, offsets u64[n+1])."""
    seq = np.empty(n_reads * READ_LEN, dtype=np.uint8)
    qual = np.empty(n_reads * READ_LEN, dtype=np.uint8) if want_qual else None
    lib().orc_synth_reads(seed, profile, first_read, n_reads, _ptr(seq), _ptr(qual))
    return seq, qual, np.arange(0, (n_reads + 1) * READ_LEN, READ_LEN, dtype=np.uint64)


def count_batch_mt(k: int, seq: np.ndarray, qual: Optional[np.ndarray], offsets: np.ndarray,
                   min_quality: Optional[int] = None, threads: int = 0, filter_mod: int = 0, filter_rem: int = 0):
    """Multi-threaded rolling oracle for BASELINE-sized inputs; optional key-space sample key % filter_mod == filter_rem.
    Returns (keys, counts, windows) with keys ascending; `windows` counts all windows (before the filter)."""
    threads = threads or (os.cpu_count() or 1)
    seq = np.ascontiguousarray(seq, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if qual is not None:
        qual = np.ascontiguousarray(qual, dtype=np.uint8)
    kp, cp, w = C.c_void_p(), C.c_void_p(), C.c_uint64()
    L = lib()
    n = L.orc_count_batch_mt(k, int(min_quality is not None), int(min_quality or 0), _ptr(seq), _ptr(qual), _ptr(offsets),
                             len(offsets) - 1, threads, filter_mod, filter_rem, C.byref(kp), C.byref(cp), C.byref(w))
    keys = np.ctypeslib.as_array(C.cast(kp, u64p), shape=(max(n, 1),))[:n].copy()
    counts = np.ctypeslib.as_array(C.cast(cp, u64p), shape=(max(n, 1),))[:n].copy()
    L.orc_free(kp); L.orc_free(cp)
    return keys, counts, int(w.value)
