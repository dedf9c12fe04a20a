"""Pins the CPU oracle (oracle/ntref.hpp) against the reference's own known-answer tests.

Every test names the reference test it transcribes (file:line under /root/reference).
CPU only — runs in the build container and on the GPU box alike (no /root/reference needed:
the reference's test *data* travels in tests/golden/ref_fixtures.tar.gz).
"""
import hashlib
import json
import os

import pytest

import oracle_lib as O
from conftest import GOLDEN, load_fixtures, parse_specimen_index


# ---------------------------------------------------------------- sequence.rs
def test_normalize_vectors():  # src/sequence.rs:316-344
    assert O.normalize(b"ACGTU", False) == (b"ACGTT", True)
    assert O.normalize(b"acgtu", False) == (b"ACGTT", True)
    assert O.normalize(b"N.N-N~N N", False) == (b"N-N-N-NN", True)
    assert O.normalize(b"BDHVRYSWKM", True) == (b"BDHVRYSWKM", False)      # None
    assert O.normalize(b"bdhvryswkm", True) == (b"BDHVRYSWKM", True)
    assert O.normalize(b"BDHVRYSWKM", False) == (b"NNNNNNNNNN", True)
    assert O.normalize(b"bdhvryswkm", False) == (b"NNNNNNNNNN", True)


def test_normalize_doc_and_python_vectors():  # src/sequence.rs:215-225 ; test_python.py:101-139
    assert O.normalize(b"ADGH", False)[0] == b"ANGN"
    assert O.normalize(b"ADGH", True) == (b"ADGH", False)
    assert O.normalize(b"ACGU", True)[0] == b"ACGT"
    for s in (b"N-N-N-N",):
        assert O.normalize(s, False) == (s, False)
    assert O.normalize(b"N.N.N.N", False)[0] == b"N-N-N-N"
    assert O.normalize(b"N~N~N~N", False)[0] == b"N-N-N-N"
    for ws in (b" ", b"\t", b"\n", b"\r"):
        assert O.normalize(b"N" + ws + b"N" + ws + b"N" + ws + b"N", False)[0] == b"NNNN"
    for ch in b"!@#$%^&*|":
        s = bytes([78, ch, 78, ch, 78, ch, 78])
        assert O.normalize(s, False)[0] == b"NNNNNNN"
    assert O.normalize(b"N9N5N1N", False)[0] == b"NNNNNNN"


def test_complement_vectors():  # src/sequence.rs:347-352 ; test_python.py:142-149
    assert O.complement(ord("a")) == ord("t")
    assert O.complement(ord("c")) == ord("g")
    assert O.complement(ord("g")) == ord("c")
    assert O.complement(ord("n")) == ord("n")
    assert O.reverse_complement(b"AACC") == b"GGTT"      # doc-test :197-201
    assert O.reverse_complement(b"atcg") == b"cgat"
    assert O.reverse_complement(b"ATCG") == b"CGAT"


def test_str_canonical_and_minimizer():  # src/sequence.rs:355-367
    assert O.str_canonical(b"A") == b"A"
    assert O.str_canonical(b"T") == b"A"
    assert O.str_canonical(b"AAGT") == b"AAGT"
    assert O.str_canonical(b"ACTT") == b"AAGT"
    assert O.str_canonical(b"GC") == b"GC"
    assert O.str_minimizer(b"ATTTCG", 3) == b"AAA"


def test_quality_mask():  # src/sequence.rs:370-374
    assert O.quality_mask(b"AGCT", b"AAA0", ord("5")) == b"AGCN"


def test_decode_phred():  # src/quality.rs:34-64 ; test_python.py:152-168
    exp = (2, 27, 14, 27, 14, 33, 33, 37, 37, 37, 33, 37, 27)
    assert O.decode_phred(b"#</</BBFFFBF<") == exp
    assert O.decode_phred(b"B[N[Naaeeeae[", base64=True) == exp
    assert O.decode_phred(b"") == ()
    assert O.decode_phred(b"#</</BBFFFBF ") is None
    assert O.decode_phred(b"B[N[Naaeeeae?", base64=True) is None


# ---------------------------------------------------------------- kmer.rs
def test_canonical_kmers_vectors():  # src/kmer.rs:170-226
    got = O.canonical_kmers(b"AGCT", 1)
    assert [(k, f) for _, k, f in got] == [(b"A", False), (b"C", True), (b"C", False), (b"A", True)]
    got = O.canonical_kmers(b"AGCTA", 2)
    assert [k for _, k, _ in got] == [b"AG", b"GC", b"AG", b"TA"]
    got = O.canonical_kmers(b"AGNTA", 2)
    assert [(p, k) for p, k, _ in got] == [(0, b"AG"), (3, b"TA")]
    # tie (reverse palindrome) => rc slice, was_rc = True  (src/kmer.rs:124-128)
    assert O.canonical_kmers(b"ACGT", 4) == [(0, b"ACGT", True)]
    # doc example compiles with k=3 on ACGT (src/kmer.rs:61-72)
    assert [p for p, _, _ in O.canonical_kmers(b"ACGT", 3)] == [0, 1]


# ---------------------------------------------------------------- bitkmer.rs
def test_bit_kmers_vectors():  # src/bitkmer.rs:192-250
    assert list(O.bit_kmers(b"AGCT", 1, False)[1]) == [0b00, 0b10, 0b01, 0b11]
    assert list(O.bit_kmers(b"ACNGT", 2, False)[1]) == [0b0001, 0b1011]
    assert list(O.bit_kmers(b"ACNG", 2, False)[1]) == [0b0001]
    assert list(O.bit_kmers(b"AC", 2, False)[1]) == [0b0001]
    pos, km, fl = O.bit_kmers(b"ACGTA", 3, False)
    assert list(zip(pos, km, fl)) == [(0, 6, 0), (1, 27, 0), (2, 44, 0)]
    assert len(O.bit_kmers(b"TA", 3, False)[0]) == 0


def test_bit_rc_canonical_minimizer_vectors():  # src/bitkmer.rs:255-266
    assert O.bit_reverse_complement(0b000000, 3) == 0b111111
    assert O.bit_reverse_complement(0b111111, 3) == 0b000000
    assert O.bit_reverse_complement(0b00000000, 4) == 0b11111111
    assert O.bit_reverse_complement(0b00011011, 4) == 0b00011011
    assert O.bit_minimizer(0b001011, 3, 2) == 0b0010
    assert O.bit_minimizer(0b001011, 3, 1) == 0b00
    assert O.bit_minimizer(0b11000011, 4, 2) == 0b0000
    assert O.bit_minimizer(0b110001, 3, 2) == 0b0001
    # quirk A.9: RC taken at width k -> minimizer((TTT,3),2) == 3, not 0
    assert O.bit_minimizer(0b111111, 3, 2) == 3
    # tie => (kmer, false)  (src/bitkmer.rs:136-143)
    assert O.bit_canonical(0b00011011, 4) == (0b00011011, False)


def test_bitmer_bytes_vectors():  # src/bitkmer.rs:271-286
    assert O.bytes_to_bitmer(b"C") == 1
    assert O.bytes_to_bitmer(b"TTA") == 60
    assert O.bytes_to_bitmer(b"AAA") == 0
    assert O.bitmer_to_bytes(1, 1) == b"C"
    assert O.bitmer_to_bytes(60, 3) == b"TTA"
    assert O.bitmer_to_bytes(0, 3) == b"AAA"


# ---------------------------------------------------------------- parser/fastq.rs
def test_simple_fastq_lf_crlf():  # src/parser/fastq.rs:473-511
    for text, le in ((b"@test\nAGCT\n+test\n~~a!\n@test2\nTGCA\n+test\nWUI9", "unix"),
                     (b"@test\r\nAGCT\r\n+test\r\n~~a!\r\n@test2\r\nTGCA\r\n+test\r\nWUI9", "windows")):
        p = O.parse_fastx(text)
        assert p.format == "fastq" and p.err_kind is None and p.line_ending == le
        assert [(r["id"], r["raw_seq"], r["qual"]) for r in p.records] == [
            (b"test", b"AGCT", b"~~a!"), (b"test2", b"TGCA", b"WUI9")]


def test_fastq_eof_rules():  # src/parser/fastq.rs:514-550
    p = O.parse_fastx(b"@test\nACGT\n+\nIII")
    assert len(p.records) == 0 and p.err_kind == "UnequalLengths"
    p = O.parse_fastx(b"@test\nAGCT\n+test\n~~a!\n@test2\nTGCA")
    assert len(p.records) == 1 and p.err_kind == "UnexpectedEnd"
    p = O.parse_fastx(b"@test\nAGCT\n+test\n~~a!\n\n")
    assert len(p.records) == 1 and p.err_kind is None
    p = O.parse_fastx(b"@test\nAGCT\n+test\n~~a!\n\n@TEST\nA\n+TEST\n~")
    assert len(p.records) == 1 and p.err_kind == "InvalidStart"


def test_fastq_empty_records():  # src/parser/fastq.rs:553-576
    p = O.parse_fastx(b"@\n\n+\n\n@test2\nTGCA\n+test2\n~~~~\n")
    assert p.err_kind is None
    r0, r1 = p.records
    assert (r0["id"], r0["raw_seq"], r0["qual"], r0["all"]) == (b"", b"", b"", b"@\n\n+\n")
    assert (r1["id"], r1["raw_seq"], r1["qual"], r1["all"]) == (b"test2", b"TGCA", b"~~~~", b"@test2\nTGCA\n+test2\n~~~~")


def test_fastq_weird_ncbi_line_numbers():  # src/parser/fastq.rs:579-594
    s = b"ACGTACGATCGTACGTAGCTGCTAGCTAGCATGCATGACACACACGTACGATCGTACGTAGCTGCTAGCTAGCATGCATGACACAC"
    q = b"0" * 86
    h = b"@NCBI actually has files like this\n"
    text = h + s + b"\n+\n" + q + b"\n" + h + b"\n+\n\n" + h + s + b"\n+\n" + q
    p = O.parse_fastx(text)
    assert p.err_kind is None and [r["line"] for r in p.records] == [1, 5, 9]


def test_fastq_mismatched_lengths():  # src/parser/fastq.rs:597-603
    p = O.parse_fastx(b"@test\nAGCT\n+\nIII\n@TEST\nA\n+\nI")
    assert len(p.records) == 0 and p.err_kind == "UnequalLengths"


def test_fastq_file_fixtures(fixtures):  # src/parser/fastq.rs:607-628
    p = O.parse_fastx(fixtures["data/bad_header.fastq"])
    assert len(p.records) == 1 and p.err_kind == "UnexpectedEnd"
    p = O.parse_fastx(fixtures["data/random_tsv.fq"])
    assert len(p.records) == 1 and p.err_kind == "InvalidSeparator"


def test_record_positions():  # src/parser/record.rs:258-285
    p = O.parse_fastx(b"@test\nACGT\n+\nIIII\n@test2\nACGT\n+\nIIII")
    assert [r["line"] for r in p.records] == [1, 5]
    p = O.parse_fastx(b"@test1\nACGT\n+\nIIII\n@test222\nACGT\n+\nIIII\n@test3\nACGT\n+\nIIII")
    assert [r["byte"] for r in p.records] == [0, 19, 40]
    p = O.parse_fastx(b"@test1\nACGT\n+\nIIII")
    assert O.decode_phred(p.records[0]["qual"]) == (40, 40, 40, 40)


# ---------------------------------------------------------------- parser/fasta.rs
def test_fasta_basic():  # src/parser/fasta.rs:389-406
    p = O.parse_fastx(b">test\nACGT\n>test2\nTGCA\n")
    assert p.format == "fasta" and p.err_kind is None and p.line_ending == "unix"
    assert [(r["id"], r["raw_seq"]) for r in p.records] == [(b"test", b"ACGT"), (b"test2", b"TGCA")]
    assert p.records[0]["all"] == b">test\nACGT"


def test_fasta_wrapped_lf_crlf():  # src/parser/fasta.rs:409-446
    p = O.parse_fastx(b">test\nACGT\nACGT\n>test2\nTGCA\nTG")
    assert [(r["id"], r["raw_seq"], r["num_bases"]) for r in p.records] == [
        (b"test", b"ACGT\nACGT", 8), (b"test2", b"TGCA\nTG", 6)]
    p = O.parse_fastx(b">test\r\nACGT\r\nACGT\r\n>test2\r\nTGCA\r\nTG")
    assert p.line_ending == "windows"
    assert [(r["id"], r["raw_seq"], r["num_bases"], r["line"]) for r in p.records] == [
        (b"test", b"ACGT\r\nACGT", 8, 1), (b"test2", b"TGCA\r\nTG", 6, 4)]


def test_fasta_premature_ending():  # src/parser/fasta.rs:449-463
    for text in (b">test\nAGCT\n>test2", b">test\r\nAGCT\r\n>test2\r\n"):
        p = O.parse_fastx(text)
        assert len(p.records) == 1 and p.err_kind == "UnexpectedEnd"


def test_fasta_empty_records():  # src/parser/fasta.rs:466-482
    for text in (b">\n\n>shine\nAGGAGGU", b">\r\n\r\n>shine\r\nAGGAGGU"):
        p = O.parse_fastx(text)
        assert p.err_kind is None
        assert [(r["id"], r["raw_seq"]) for r in p.records] == [(b"", b""), (b"shine", b"AGGAGGU")]


# ---------------------------------------------------------------- parser/mod.rs + python
def test_sniff_rules(fixtures):  # src/parser/mod.rs:182-200,37-46 ; test_python.py:171-226
    assert O.parse_fastx(b"").err_kind == "EmptyFile"
    assert O.parse_fastx(b"@").err_kind == "EmptyFile"
    assert O.parse_fastx(b"Not a valid file").err_kind == "UnknownFormat"
    assert O.parse_fastx(fixtures["data/bad_test.fa"]).err_kind == "UnknownFormat"
    p = O.parse_fastx(fixtures["data/test.fa"])
    assert [(r["id"], O.strip_returns(r["raw_seq"])[0]) for r in p.records] == [
        (b"test", b"AGCTGATCGA"), (b"test2", b"TAGC")]
    p = O.parse_fastx(fixtures["specimen/FASTQ/example.fastq"])
    assert (p.records[0]["id"], p.records[0]["raw_seq"], p.records[0]["qual"]) == (
        b"EAS54_6_R1_2_1_413_324", b"CCCTTCTTGTCTTCAGCGTTTCTCC", b";;3;;;;;;;;;;;;7;;;;;;;88")
    assert (p.records[1]["id"], p.records[1]["raw_seq"], p.records[1]["qual"]) == (
        b"EAS54_6_R1_2_1_540_792", b"TTGGCAGGCCAAGGCCGATGGATCA", b";;;;;;;;;;;7;;;;;-;;;3;83")


def test_stdin_example():  # tests/test_stdin.rs:30-31,138-139 ; examples/stdin_pipe.rs
    t = O.tally_fastx(b">id1\nAGTCGTCA", k=4, query=b"AAAA")
    assert t["n_bases"] == 8 and t["n_query"] == 0


# ---------------------------------------------------------------- corpus verdicts
def test_specimen_fasta(fixtures):  # tests/format_specimens.rs:29-46
    idx = parse_specimen_index(fixtures["specimen/FASTA/index.toml"].decode())
    assert len(idx["valid"]) + len(idx["invalid"]) == 48
    for t in idx["valid"]:
        if "comments" in t["tags"]:
            continue
        p = O.parse_fastx(fixtures["specimen/FASTA/" + t["filename"]])
        assert p.err_kind is None, t["filename"]


def test_specimen_fastq(fixtures):  # tests/format_specimens.rs:48-94
    idx = parse_specimen_index(fixtures["specimen/FASTQ/index.toml"].decode())
    assert len(idx["valid"]) + len(idx["invalid"]) == 59
    skip_valid = {"wrapping_original_sanger.fastq", "longreads_original_sanger.fastq", "tricky.fastq"}
    for t in idx["valid"]:
        if t["filename"] in skip_valid:
            continue
        p = O.parse_fastx(fixtures["specimen/FASTQ/" + t["filename"]])
        assert p.err_kind is None, t["filename"]
    for t in idx["invalid"]:
        fn = t["filename"]
        if fn == "error_diff_ids.fastq" or fn.startswith("error_qual_") or fn in ("error_spaces.fastq", "error_tabs.fastq"):
            continue
        p = O.parse_fastx(fixtures["specimen/FASTQ/" + fn])
        assert p.err_kind is not None, fn


# ---------------------------------------------------------------- end-to-end pinned constants
def test_benchmark_constants_28S(fixtures):  # benches/benchmark.rs:43-44,66-67,151,166,180
    data = fixtures["data/28S.fasta"]
    t = O.tally_fastx(data, k=31, m=0, iupac=True)
    assert t["err_kind"] is None
    assert t["n_records"] == 570 and t["n_bases"] == 738_580
    assert t["n_kmers"] == 718_007 and t["n_not_rc"] == 350_983
    # bit path: strip_returns -> bit_kmers(31, canonical=true)  (benches/benchmark.rs:55-67)
    p = O.parse_fastx(data)
    n_total = n_canon = 0
    for r in p.records:
        s, _ = O.strip_returns(r["raw_seq"])
        pos, km, fl = O.bit_kmers(s, 31, True)
        n_total += len(pos); n_canon += int((fl == 0).sum())
    assert (n_total, n_canon) == (718_007, 350_983)


def test_fastq_bases_constant(fixtures):  # benches/benchmark.rs:97,111,125
    p = O.parse_fastx(fixtures["data/PRJNA271013_head.fq"])
    assert p.err_kind is None and len(p.records) == 2000
    assert sum(r["num_bases"] for r in p.records) == 250_000


def test_derived_goldens(fixtures):
    """Derived (oracle-computed) goldens from SURVEY.md §8(c); recorded in tests/golden/derived.json
    so the GPU parity tests and later rounds can detect an oracle drift."""
    with open(os.path.join(GOLDEN, "derived.json")) as f:
        g = json.load(f)
    t = O.tally_fastx(fixtures["data/28S.fasta"], k=4, m=0, iupac=False, query=b"AAAA")
    assert (t["n_records"], t["n_kmers"], t["n_not_rc"], t["n_query"]) == (570, 736_277, 350_631, 8_108)
    assert t == g["28S_k4_AAAA"]
    t = O.tally_fastx(fixtures["data/PRJNA271013_head.fq"], k=31, m=21, iupac=False)
    assert (t["n_records"], t["n_kmers"], t["n_not_rc"]) == (2000, 189_960, 95_997)
    assert t == g["PRJNA_k31_m21"]


def test_fixture_bundle_integrity(fixtures):
    with open(os.path.join(GOLDEN, "MANIFEST.json")) as f:
        man = json.load(f)
    assert set(man) == set(fixtures)
    for k, v in man.items():
        assert hashlib.sha256(fixtures[k]).hexdigest() == v["sha256"], k
    ref = "/root/reference/tests"
    if os.path.isdir(ref):   # build container only: the bundle equals the reference's files
        for k in man:
            assert open(os.path.join(ref, k), "rb").read() == fixtures[k], k
