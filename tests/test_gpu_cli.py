"""The C++ host above the C ABI (krust_b200/host: kmerust.hpp + the kmerust-b200 binary) driven like the reference's
integration tests drive its CLI (tests/integration_tests.rs): exact-value lines, histogram, --save + query, exit codes."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "krust_b200", "host", "kmerust-b200")
FX = os.path.join(ROOT, "tests", "golden", "fixtures")


def run(*args, stdin=None):
    return subprocess.run([BIN, *args], input=stdin, capture_output=True, timeout=120)


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-C", os.path.dirname(BIN), "-s"])


def test_exact_value_lines_and_histogram(tmp_path):
    r = run("3", os.path.join(FX, "soft_masked.fa"), "--format", "tsv", "-q")
    assert r.returncode == 0 and r.stdout == b"AAA\t2\n"                      # tests/integration_tests.rs:263-281
    r = run("3", os.path.join(FX, "soft_masked.fa"), "-q")
    assert r.returncode == 0 and r.stdout == b">2\nAAA\n"
    a8 = tmp_path / "a8.fa"
    a8.write_bytes(b">s\nAAAAAAAA\n")
    r = run("3", str(a8), "--format", "histogram", "-q")
    assert r.returncode == 0 and r.stdout == b"6\t1\n"                        # tests/integration_tests.rs:767-799
    fa = run("3", os.path.join(FX, "simple.fa"), "-f", "tsv", "-q").stdout
    fq = run("3", os.path.join(FX, "simple.fq"), "-f", "tsv", "-q").stdout
    assert fa == fq == b"AAT\t1\nACA\t1\nACG\t4\nATC\t1\nGTA\t3\nTAA\t1\n"   # FASTA == FASTQ (:486-523), sorted
    r = run("3", os.path.join(FX, "simple.fa"), "-f", "tsv", "-q", "--min-count", "2")
    assert r.stdout == b"ACG\t4\nGTA\t3\n"
    r = run("4", os.path.join(FX, "low_quality.fq"), "-f", "tsv", "-q", "-Q", "20")
    assert r.stdout == b"AATC\t1\nACGT\t1\nATTA\t1\nGTAA\t1\nTACA\t1\n"       # src/streaming.rs:1165-1189
    r = run("3", "-", "-f", "tsv", "-q", stdin=b">x\nACGT\n")
    assert r.returncode == 0 and r.stdout == b"ACG\t2\n"                      # stdin (tests/library_tests.rs:23-33)
    r = run("3", "-", "-f", "json", "-q", stdin=b">x\nACGT\n")
    assert b'"kmer": "ACG"' in r.stdout and b'"count": 2' in r.stdout


def test_bad_arguments_and_missing_files():
    assert run("0", os.path.join(FX, "simple.fa")).returncode == 2             # src/cli.rs:103-114
    assert run("33", os.path.join(FX, "simple.fa")).returncode == 2
    assert run().returncode == 2
    r = run("3", "/nonexistent/file.fa")
    assert r.returncode == 1 and b"File not found" in r.stderr                # src/main.rs:58-67


def test_save_index_and_query_on_the_device(tmp_path):
    rng = np.random.default_rng(4)
    recs = [bytes(rng.choice(list(b"ACGT"), size=500).tolist()) for _ in range(40)] + [b"GATTACAGATTACAGATTACA"]
    hit, canon = b"GATTACAGATT", b"AATCTGTAATC"   # occurs twice in the last record; stored under its (smaller) reverse complement
    fa = tmp_path / "in.fa"
    fa.write_bytes(b"".join(b">r%d\n%s\n" % (i, s) for i, s in enumerate(recs)))
    idx = tmp_path / "out.kmix"
    r = run("11", str(fa), "-f", "tsv", "-q", "--min-count", "2", "--save", str(idx))
    okeys, ocounts, _ = orc.count_records(11, recs, mode="rolling")
    assert r.returncode == 0
    assert idx.read_bytes() == orc.kmix_encode(11, okeys, ocounts)              # the index is never min-count filtered (src/main.rs:155-212)
    want = b"".join(b"%s\t%d\n" % (orc.unpack(int(k), 11), int(c)) for k, c in zip(okeys, ocounts) if c >= 2)
    assert r.stdout == want
    table = {orc.unpack(int(k), 11): int(c) for k, c in zip(okeys, ocounts)}
    r = run("query", str(idx), hit.decode())
    assert r.returncode == 0 and int(r.stdout) == table[canon] >= 2
    r = run("query", str(idx), "aatctgtaatc")                                 # reverse complement, lower case
    assert r.returncode == 0 and int(r.stdout) == table[canon]
    absent = next(s for s in (b"AAAAAAAAAAA", b"CCCCCCCCCCC", b"ACACACACACA", b"AGAGAGAGAGA") if s not in table)
    assert int(run("query", str(idx), absent.decode()).stdout) == 0
    assert run("query", str(idx), "GATTA").returncode == 1                    # length mismatch
    assert run("query", str(idx), "GATTNCAGATT").returncode == 1              # invalid base
    bad = tmp_path / "bad.kmix"
    bad.write_bytes(b"XXXX" + idx.read_bytes()[4:])
    r = run("query", str(bad), hit.decode())
    assert r.returncode == 1 and b"invalid magic" in r.stderr
    gz = tmp_path / "out.kmix.gz"
    r = run("11", str(fa), "-q", "--save", str(gz), "-f", "histogram")
    assert r.returncode == 0
    import gzip
    assert gzip.decompress(gz.read_bytes()) == idx.read_bytes()               # a .gz name gets a gzip stream (src/index.rs:159-168)
    assert int(run("query", str(gz), hit.decode()).stdout) == table[canon]
