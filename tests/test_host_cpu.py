"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the host FASTA/FASTQ splitter agrees with the oracle's, the .kmix host helpers follow the
reference format, and creating a counter without a GPU fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import krust_b200 as kb
from krust_b200 import _lib
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "kmerust_gpu.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(kmg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/kmerust_gpu.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.kmg_abi_version() == 1
    assert b"1..=32" in L.kmg_status_string(_lib.KMG_ERR_INVALID_K)


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.KmgConfig) == 48
    assert C.sizeof(_lib.KmgSummary) == 88
    assert C.sizeof(_lib.KmgBatch) == 56


def test_kmer_length_errors_like_reference():
    # src/kmer.rs:100-111; tests/library_tests.rs:155-170
    for bad in (0, 33, 100):
        with pytest.raises(kb.KmerLengthError) as e:
            kb.KmerLength(bad)
        assert e.value.k == bad and e.value.min == 1 and e.value.max == 32
    assert kb.KmerLength(1).get() == 1 and kb.KmerLength(32).as_u8() == 32
    with pytest.raises(kb.KmerLengthError):
        kb.KmerCounter.new().k(0)
    with pytest.raises(kb.BuilderError):
        kb.KmerCounter.new().count("nonexistent.fa")  # k not set is reported first (src/builder.rs:246)


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(kb.GpuError) as e:
        kb.GpuKmerCounter(21)
    assert e.value.status == _lib.KMG_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(kb.KmeRustError):
        kb.count_kmers_from_sequences([b"ACGT"], kb.KmerLength(3))


def test_create_rejects_bad_config_before_touching_cuda():
    import ctypes as C
    L = _lib.load()
    ctx = C.c_void_p()
    cfg = _lib.KmgConfig(abi_version=1, k=0, device=-1)
    assert L.kmg_create(C.byref(cfg), C.byref(ctx)) == _lib.KMG_ERR_INVALID_K
    assert b"out of range" in L.kmg_last_error(None)
    cfg = _lib.KmgConfig(abi_version=7, k=21, device=-1)
    assert L.kmg_create(C.byref(cfg), C.byref(ctx)) == _lib.KMG_ERR_ABI
    cfg = _lib.KmgConfig(abi_version=1, k=21, device=-1, flags=_lib.KMG_FLAG_FORCE_DIRECT)
    assert L.kmg_create(C.byref(cfg), C.byref(ctx)) == _lib.KMG_ERR_INVALID_ARG


def test_unpack_matches_oracle():
    rng = np.random.default_rng(0)
    for k in (1, 5, 21, 31, 32):
        keys = rng.integers(0, 2**63, size=64, dtype=np.uint64)
        if k < 32:
            keys &= np.uint64((1 << (2 * k)) - 1)
        names = kb.unpack_many(keys, k)
        for key, name in zip(keys.tolist(), names.tolist()):
            assert name == orc.unpack(key, k) and kb.unpack_to_string(key, k) == name.decode()


PARSER_CASES = [
    (b">seq1\nACGTACGT\n>seq2\nGATTACA\n", False),
    (b">s desc\nACGT\r\nACGT\n>t\nGG\n\n", False),
    (b">a\n>b\nAC\n", False),
    (b"", False),
    (b">only\n", False),
    (b">x\nAC GT \nNN\n", False),
    (b"@r1\nACGT\n+\nIIII\n@r2\nGG\nTT\n+r2\n!!\n##\n", True),
    (b"@r1\nACGT\n+\nIIII", True),
    (b"", True),
]


@pytest.mark.parametrize("data,is_fastq", PARSER_CASES)
def test_host_parser_equals_oracle_parser(data, is_fastq):
    a = kb.parse_fastx(data, is_fastq)
    b = orc.parse_fastx(data, is_fastq)
    assert a[0].tobytes() == b[0].tobytes() and a[2].tolist() == b[2].tolist()
    if is_fastq:
        assert a[1].tobytes() == b[1].tobytes()


def test_host_parser_errors_and_fixtures(golden_dir):
    with pytest.raises(kb.SequenceParseError):
        kb.parse_fastx(b"ACGT\n", False)
    with pytest.raises(kb.SequenceParseError):
        kb.parse_fastx(b"@r\nACGT\n+\nII\n", True)
    with pytest.raises(kb.SequenceParseError):
        kb.parse_fastx(b"@r\nACGT\n", True)
    fx = os.path.join(golden_dir, "fixtures")
    seq, qual, off = kb.read_records(os.path.join(fx, "low_quality.fq"))
    assert seq.tobytes() == b"ACGTACGTGATTACA" and qual.tobytes() == b"IIII!!!!IIIIIII" and off.tolist() == [0, 8, 15]
    assert kb.SequenceFormat.from_extension("reads.fastq.gz") == "fastq"
    assert kb.SequenceFormat.from_extension("genome.fa.gz") == "fasta"
    assert kb.SequenceFormat.from_extension("noext") == "fasta"
    assert kb.SequenceFormat.resolve("auto", None) == "fasta"
    with pytest.raises(kb.KmeRustError):
        kb.read_records("/nonexistent/file.fa")


def test_index_host_round_trip_and_rejects(tmp_path):
    # src/index.rs:510-573; tests/property_tests.rs:244-261
    rng = np.random.default_rng(9)
    for k in (1, 5, 16, 21, 32):
        counts = {int(a): int(b) for a, b in zip(rng.integers(0, 2**63, 40, dtype=np.uint64), rng.integers(1, 2**40, 40))}
        p = tmp_path / f"t{k}.kmix"
        kb.save_index(kb.KmerIndex(kb.KmerLength(k), counts), p)
        blob = p.read_bytes()
        k2, keys, cnts = orc.kmix_decode(blob)  # the oracle's reader accepts our writer's bytes
        assert k2 == k and dict(zip(keys.tolist(), cnts.tolist())) == counts
        idx = kb.load_index(p)
        assert idx.k() == k and idx.counts() == counts and len(idx) == len(counts)
        pz = tmp_path / f"t{k}.kmix.gz"
        kb.save_index(kb.KmerIndex(kb.KmerLength(k), counts), pz)
        assert kb.load_index(pz).counts() == counts
    (tmp_path / "small").write_bytes(b"KMIX")
    with pytest.raises(kb.InvalidIndexError, match="too small"):
        kb.load_index(tmp_path / "small")
    (tmp_path / "magic").write_bytes(b"XXXX" + blob[4:])
    with pytest.raises(kb.InvalidIndexError, match="magic"):
        kb.load_index(tmp_path / "magic")
    bad = bytearray(blob); bad[30] ^= 1
    (tmp_path / "crc").write_bytes(bytes(bad))
    with pytest.raises(kb.InvalidIndexError, match="checksum"):
        kb.load_index(tmp_path / "crc")
    # the oracle's writer is readable by our loader too
    (tmp_path / "o.kmix").write_bytes(orc.kmix_encode(21, np.array([5, 9], dtype=np.uint64), np.array([1, 2], dtype=np.uint64)))
    assert kb.load_index(tmp_path / "o.kmix").counts() == {5: 1, 9: 2}


def test_histogram_host_helpers():
    assert kb.compute_histogram_packed({1: 1, 2: 1, 3: 2, 4: 2}) == {1: 2, 2: 2}
    st = kb.histogram_stats({1: 2, 2: 2})
    assert st["distinct_kmers"] == 4 and st["total_kmers"] == 6 and st["mean_count"] == 1.5
    assert kb.histogram_stats({})["mean_count"] == 0.0


def test_owner_of_is_a_partition():
    keys = np.random.default_rng(4).integers(0, 2**42, 2000, dtype=np.uint64)
    for n in (1, 2, 4, 8):
        owners = [kb.owner_of(int(x), n) for x in keys]
        assert all(0 <= o < n for o in owners)
        if n > 1:
            assert len(set(owners)) == n


def test_mixed_key_constants_are_consistent():
    """The partitioned pipeline stores mix64(key) in its runs and un-mixes on export: the constants in kmg_device.cuh must be
    modular inverses of each other, EMPTY_MIX must be mix64(EMPTY_KEY), and the host-visible owner function must agree with a
    Python restatement of the mix."""
    import re
    import pathlib
    import random
    import krust_b200 as kb
    src = (pathlib.Path(__file__).resolve().parents[1] / "krust_b200" / "csrc" / "kmg_device.cuh").read_text()
    M = (1 << 64) - 1
    muls = [int(x, 16) for x in re.findall(r"x \*= (0x[0-9a-fA-F]+)ull", src)]
    assert len(muls) == 4, muls                      # two in mix64, two in unmix64
    m1, m2, i2, i1 = muls
    assert (m1 * i1) & M == 1 and (m2 * i2) & M == 1
    empty_mix = int(re.search(r"EMPTY_MIX = (0x[0-9a-fA-F]+)ull", src).group(1), 16)

    def mix(x):
        x ^= x >> 33; x = (x * m1) & M; x ^= x >> 33; x = (x * m2) & M; x ^= x >> 33
        return x

    def unmix(x):
        x ^= x >> 33; x = (x * i2) & M; x ^= x >> 33; x = (x * i1) & M; x ^= x >> 33
        return x

    assert mix(M) == empty_mix
    rnd = random.Random(5)
    for _ in range(2000):
        key = rnd.getrandbits(64)
        assert unmix(mix(key)) == key
        n = rnd.choice([2, 3, 8, 657, 928, 4096])
        assert kb.owner_of(key, n) == ((mix(key) >> 32) * n) >> 32


def test_sharded_kmix_crc_combination_on_host(tmp_path):
    """kmg_kmix_begin / kmg_kmix_finish are host-only: records written by 'shards' at their offsets, per-shard CRC-32s combined
    into the file CRC (src/index.rs:222-279 format; zlib's crc32 as the independent CRC)."""
    import zlib
    rng = np.random.default_rng(3)
    for sizes in ([5, 0, 11], [1], [0, 0], [4096, 1, 70000]):
        p = tmp_path / "s.kmix"
        kb.kmix_begin(p)
        keys = rng.integers(0, 4**21, size=sum(sizes), dtype=np.uint64)
        counts = rng.integers(1, 1000, size=sum(sizes), dtype=np.uint64)
        crcs, off = [], 0
        with open(p, "r+b") as f:
            for n in sizes:
                body = np.stack([keys[off:off + n], counts[off:off + n]], axis=1).astype("<u8").tobytes()
                f.seek(14 + 16 * off); f.write(body)
                crcs.append(zlib.crc32(body) & 0xFFFFFFFF)
                off += n
        kb.kmix_finish(p, 21, sizes, crcs)
        blob = p.read_bytes()
        assert blob == orc.kmix_encode(21, keys, counts)
        assert kb.load_index(p).k().get() == 21
    with pytest.raises(kb.KmeRustError):
        kb.kmix_begin(tmp_path / "x.kmix.gz")


def test_index_open_rejects_invalid_files_like_the_reference(tmp_path):
    """kmg_index_open validates before it touches CUDA: the checks and their order are read_index's
    (src/index.rs:282-401; reference tests src/index.rs:527-573: too small, bad magic, flipped byte)."""
    good = orc.kmix_encode(5, np.array([1, 7, 300], dtype=np.uint64), np.array([2, 9, 1], dtype=np.uint64))
    cases = {"small": (good[:10], "file too small"), "magic": (b"XMIX" + good[4:], "invalid magic"),
             "flip": (good[:20] + bytes([good[20] ^ 1]) + good[21:], "checksum mismatch"),
             "trunc": (good[:-20] + good[-4:], "checksum mismatch")}
    import struct, zlib
    def with_crc(body):
        return body + struct.pack("<I", zlib.crc32(body) & 0xFFFFFFFF)
    cases["version"] = (with_crc(good[:4] + b"\x02" + good[5:-4]), "unsupported version 2")
    cases["k"] = (with_crc(good[:5] + b"\x28" + good[6:-4]), "invalid k-mer length")
    cases["size"] = (with_crc(good[:6] + struct.pack("<Q", 9) + good[14:-4]), "data size mismatch")
    for name, (blob, msg) in cases.items():
        p = tmp_path / f"{name}.kmix"
        p.write_bytes(blob)
        with pytest.raises(kb.InvalidIndexError) as e:
            kb.GpuKmerCounter.open_index(p)
        assert msg in str(e.value), (name, str(e.value))
        with pytest.raises(Exception):
            kb.load_index(p)            # the host-side reader rejects the same files
    with pytest.raises(kb.KmeRustError):
        kb.GpuKmerCounter.open_index(tmp_path / "missing.kmix")


def test_rust_sys_bindings_are_generated_from_the_header():
    """bindings/rust/kmerust-gpu-sys/src/lib.rs cannot be compiled here (no Rust toolchain): it is GENERATED from
    include/kmerust_gpu.h by tools/gen_rust_sys.py, and this test fails when the committed file and the header drift."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    committed = open(gen.OUT).read()
    assert committed == gen.render(), "run `python tools/gen_rust_sys.py` after changing include/kmerust_gpu.h"
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "kmerust_gpu.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(kmg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(re.findall(r"pub fn (kmg_\w+)", committed))
    # every sys function the safe wrapper calls exists
    wrapper = open(os.path.join(ROOT, "bindings", "rust", "kmerust-gpu", "src", "lib.rs")).read()
    assert set(re.findall(r"sys::(kmg_[a-z0-9_]+)\(", wrapper)) <= declared
    # field-for-field agreement of the three #[repr(C)] structs with the ctypes mirror
    for name, st in (("kmg_config", _lib.KmgConfig), ("kmg_summary", _lib.KmgSummary), ("kmg_batch", _lib.KmgBatch)):
        body = re.search(r"pub struct %s \{(.*?)\}" % name, committed, flags=re.S).group(1)
        assert re.findall(r"pub (\w+):", body) == [f for f, _ in st._fields_]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the arm the driver times beside ours) runs without a GPU: the restated reference CPU path on a
    bounded sample, one JSON line with the base contract's keys plus impl / cpu_baseline / e2e."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-bases", "2e6"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "canonical k-mers counted/sec" and d["unit"] == "kmers/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
