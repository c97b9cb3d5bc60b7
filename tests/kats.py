"""Known-answer tests taken from the reference's own test-suite.

Each entry: (name, reference citation, k, records, quals|None, min_quality|None, expected {kmer: count}).
`expected=None` together with `expect_len` pins only the number of distinct k-mers (that is all the
reference test asserts).  The same table is run against the CPU oracle (tests/test_oracle_golden.py)
and, through the C ABI, against the CUDA path (tests/test_gpu_kats.py).
"""

A32 = "A" * 32

COUNT_KATS = [
    ("basic_ACGT_k3", "tests/library_tests.rs:23-33", 3, [b"ACGT"], None, None, {"ACG": 2}),
    ("canonical_TTT", "tests/library_tests.rs:55-64", 3, [b"TTT"], None, None, {"AAA": 1}),
    ("n_bases", "tests/library_tests.rs:67-80", 3, [b"ACGNACG"], None, None, {"ACG": 2}),
    ("soft_masked_acgt", "tests/library_tests.rs:83-90", 3, [b"acgt"], None, None, {"ACG": 2}),
    ("mixed_case", "tests/library_tests.rs:93-99", 3, [b"AcGt"], None, None, {"ACG": 2}),
    ("short_sequence", "tests/library_tests.rs:102-108", 3, [b"AC"], None, None, {}),
    ("exact_length", "tests/library_tests.rs:111-118", 3, [b"ACG"], None, None, {"ACG": 1}),
    ("multiple_sequences", "tests/library_tests.rs:121-127", 3, [b"ACG", b"ACG"], None, None, {"ACG": 2}),
    ("k1", "tests/library_tests.rs:130-140", 1, [b"ACGT"], None, None, {"A": 2, "C": 2}),
    ("k32_polyA", "tests/library_tests.rs:143-152", 32, [A32.encode()], None, None, {A32: 1}),
    ("k4_palindrome", "tests/library_tests.rs:199-206", 4, [b"ACGT"], None, None, {"ACGT": 1}),
    ("AAAAA_k3", "tests/library_tests.rs:209-217", 3, [b"AAAAA"], None, None, {"AAA": 3}),
    ("soft_masked_fixture", "tests/library_tests.rs:262-270; tests/integration_tests.rs:263-281", 3, [b"AAAa"], None, None, {"AAA": 2}),
    ("AAAA_TTTT_k4", "src/streaming.rs:1150-1162", 4, [b"AAAA", b"TTTT"], None, None, {"AAAA": 2}),
    ("quality_q20", "src/streaming.rs:1165-1189", 4, [b"ACGTACGT"], [b"IIII!!!!"], 20, {"ACGT": 1}),
    ("hist_A8_k3", "tests/integration_tests.rs:767-799", 3, [b"AAAAAAAA"], None, None, {"AAA": 6}),
    ("empty_input", "tests/library_tests.rs:178-196", 3, [], None, None, {}),
    ("header_only_record", "tests/library_tests.rs:178-196", 3, [b""], None, None, {}),
    # Derived fixture answers (SURVEY.md 8c "Derived fixture answers"; consistent with every assertion above).
    ("simple_fa_k3", "fixtures/simple.fa", 3, [b"ACGTACGT", b"GATTACA"], None, None,
     {"AAT": 1, "ACA": 1, "ACG": 4, "ATC": 1, "GTA": 3, "TAA": 1}),
    ("simple_fa_k4", "fixtures/simple.fa", 4, [b"ACGTACGT", b"GATTACA"], None, None,
     {"AATC": 1, "ACGT": 2, "ATTA": 1, "CGTA": 2, "GTAA": 1, "GTAC": 1, "TACA": 1}),
    ("simple_fa_k1", "fixtures/simple.fa", 1, [b"ACGTACGT", b"GATTACA"], None, None, {"A": 9, "C": 6}),
    ("with_n_fa_k3", "fixtures/with_n.fa", 3, [b"ACGTNACGT", b"NNNGATTACANNN"], None, None,
     {"AAT": 1, "ACA": 1, "ACG": 4, "ATC": 1, "GTA": 1, "TAA": 1}),
    ("with_n_fa_k4", "fixtures/with_n.fa", 4, [b"ACGTNACGT", b"NNNGATTACANNN"], None, None,
     {"AATC": 1, "ACGT": 2, "ATTA": 1, "GTAA": 1, "TACA": 1}),
    ("low_quality_fq_k4_noQ", "fixtures/low_quality.fq", 4, [b"ACGTACGT", b"GATTACA"], [b"IIII!!!!", b"IIIIIII"], None,
     {"AATC": 1, "ACGT": 2, "ATTA": 1, "CGTA": 2, "GTAA": 1, "GTAC": 1, "TACA": 1}),
    ("low_quality_fq_k4_Q0", "src/streaming.rs:1207-1222", 4, [b"ACGTACGT", b"GATTACA"], [b"IIII!!!!", b"IIIIIII"], 0,
     {"AATC": 1, "ACGT": 2, "ATTA": 1, "CGTA": 2, "GTAA": 1, "GTAC": 1, "TACA": 1}),
    ("low_quality_fq_k4_Q20", "fixtures/low_quality.fq", 4, [b"ACGTACGT", b"GATTACA"], [b"IIII!!!!", b"IIIIIII"], 20,
     {"AATC": 1, "ACGT": 1, "ATTA": 1, "GTAA": 1, "TACA": 1}),
    ("no_qual_ignores_Q", "src/streaming.rs:1225-1239; tests/quality_tests.rs:88-114", 4, [b"ACGTACGT"], None, 20,
     {"ACGT": 2, "CGTA": 2, "GTAC": 1}),
    ("ACGTx3_k5", "SURVEY 8c", 5, [b"ACGTACGTACGT"], None, None, {"ACGTA": 4, "CGTAC": 4}),
    ("proptest_regression_AA", "tests/property_tests.proptest-regressions:7", 1, [b"AA"], None, None, {"A": 2}),
]

# kmer.rs:688-728 canonical KATs: (input, canonical string, is_reverse_complement)
CANONICAL_KATS = [
    (b"TTT", "AAA", True),
    (b"AAA", "AAA", False),
    (b"ACGT", "ACGT", False),
    (b"GATTACA", "GATTACA", False),
    (b"TGTAATC", "GATTACA", True),
]

# kmer.rs:646-661 invalid base positions
INVALID_BASE_KATS = [(b"NACNN", 0), (b"ANCNG", 1), (b"AANTG", 2), (b"CCCNG", 3), (b"AACTN", 4)]

# kmer.rs:299-302, :836-841
PACK_KATS = [(b"ACGT", 0b00011011)]

# index.rs:588-592
CRC_KATS = [(b"", 0), (b"123456789", 0xCBF43926)]
