"""GPU parity tests: libntgpu (through its C ABI) vs the CPU oracle on the same inputs.

Bit-exact everywhere (integer / byte / index work).  Run on a B200:  pytest tests -m gpu
"""
import random

import numpy as np
import pytest

import oracle_lib as O
from conftest import load_fixtures, parse_specimen_index

pytestmark = pytest.mark.gpu

TALLY_KEYS = ["n_records", "n_bases", "n_kmers", "n_not_rc", "kmer_sum_lo", "kmer_sum_hi", "n_query",
              "n_minimizers", "minimizer_sum", "err_kind"]


@pytest.fixture(scope="module")
def ctx():
    import needletail_b200 as nt
    c = nt.Context(0)
    yield c
    c.close()


def assert_tallies(ctx, data, k, m=0, query=None, what=""):
    exp = O.tally_fastx(bytes(data), k=k, m=m, iupac=False, query=query)
    got = ctx.tally(data, k=k, m=m, query=query)
    for key in TALLY_KEYS:
        assert got[key] == exp[key], f"{what} k={k} m={m}: {key}: gpu {got[key]} != oracle {exp[key]}"
    return got


def assert_parse(ctx, data, what=""):
    exp = O.parse_fastx(bytes(data))
    got = ctx.parse(data)
    assert got.format == exp.format, what
    assert got.err_kind == exp.err_kind, f"{what}: {got.err_kind} vs {exp.err_kind}"
    assert len(got.records) == len(exp.records), what
    if exp.err_kind not in (None, "EmptyFile", "UnknownFormat"):
        assert got.err_line == exp.err_line, what
        assert got.err_id == exp.err_id, what
    if exp.records:
        assert got.line_ending == exp.line_ending, what
        # columns: start,id_b,id_e,seq_b,seq_e,qual_b,qual_e,all_e,num_bases,line  (oracle adds pos_byte == start)
        et = exp.table[:, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]].copy()
        gt = got.table.copy()
        empty = et[:, 3] == et[:, 4]          # empty raw_seq: the slice offset is arbitrary
        et[empty, 3] = et[empty, 4] = gt[empty, 3] = gt[empty, 4] = 0
        assert np.array_equal(gt, et), what
        assert np.array_equal(exp.table[:, 10], exp.table[:, 0])      # position().byte() == start
    assert (got.final_line, got.final_byte) == (exp.final_line, exp.final_byte), what
    return got


# ------------------------------------------------------------------ the reference's own vectors, on the GPU
def test_normalize_vectors_gpu(ctx):  # src/sequence.rs:316-344 ; test_python.py:101-139
    seqs = [b"ACGTU", b"acgtu", b"N.N-N~N N", b"BDHVRYSWKM", b"bdhvryswkm", b"", b"N\tN\nN\rN", b"N9N5N1N", b"ADGH"]
    for iupac in (False, True):
        got, ch = ctx.normalize(seqs, iupac)
        for s, g, c in zip(seqs, got, ch):
            assert (g, c) == O.normalize(s, iupac), (s, iupac)
    assert ctx.normalize([b"BDHVRYSWKM"], True) == ([b"BDHVRYSWKM"], [False])


def test_python_face_vectors(ctx):  # test_python.py:101-149,171-226
    import needletail_b200 as nt
    assert nt.normalize_seq("ACGTU", iupac=False, ctx=ctx) == "ACGTT"
    rec = nt.Record("test", "AGCTGYrtcga")                                   # test_python.py:37-42
    rec.normalize(iupac=True, ctx=ctx)
    assert rec.seq == "AGCTGYRTCGA"
    rec.normalize(ctx=ctx)
    assert rec.seq == "AGCTGNNTCGA"
    assert nt.normalize_seq("bdhvryswkm", iupac=True, ctx=ctx) == "BDHVRYSWKM"
    assert nt.reverse_complement("atcg", ctx=ctx) == "cgat"
    assert nt.reverse_complement("ATCG", ctx=ctx) == "CGAT"
    assert nt.reverse_complement("n", ctx=ctx) == "n"
    fx = load_fixtures()
    recs = list(nt.parse_fastx_string(fx["data/test.fa"], ctx=ctx))
    assert [(r.id, r.seq, r.qual) for r in recs] == [("test", "AGCTGATCGA", None), ("test2", "TAGC", None)]
    recs = list(nt.parse_fastx_string(fx["specimen/FASTQ/example.fastq"], ctx=ctx))
    assert (recs[0].id, recs[0].seq, recs[0].qual) == ("EAS54_6_R1_2_1_413_324", "CCCTTCTTGTCTTCAGCGTTTCTCC", ";;3;;;;;;;;;;;;7;;;;;;;88")
    with pytest.raises(nt.NeedletailError):
        list(nt.parse_fastx_string("Not a valid file", ctx=ctx))
    recs = list(nt.parse_fastx_string(fx["data/test.fa.gz"], ctx=ctx))      # tests/test_compressed.rs:21-33
    assert [(r.id, r.seq) for r in recs] == [("test", "AGCTGATCGA"), ("test2", "TAGC")]


def test_strip_revcomp_qmask_gpu(ctx):
    seqs = [b"ACGT\nACGT", b"ACGT\r\nAC\r", b"", b"AACC", b"atcgRYKMBVDHSWrykmbvdhswNn-U*", b"\n\n"]
    got, ch = ctx.strip_returns(seqs)
    for s, g, c in zip(seqs, got, ch):
        assert (g, c) == O.strip_returns(s)
    for s, g in zip(seqs, ctx.reverse_complement(seqs)):
        assert g == O.reverse_complement(s)
    assert ctx.reverse_complement([b"AACC"]) == [b"GGTT"]             # src/sequence.rs:197-201
    assert ctx.quality_mask([b"AGCT"], [b"AAA0"], ord("5")) == [b"AGCN"]   # src/sequence.rs:370-374


def test_kmers_windows_and_strip_nul_and_qmask_lengths(ctx):
    import needletail_b200 as nt
    # Sequence::kmers (src/kmer.rs:13-41, src/sequence.rs:245): every window, whatever the bytes
    seqs = [b"ACGNT\nAC", b"", b"AC", b"ACG", b"xyz-!"]
    for k in (1, 3, 5, 9):
        got = ctx.kmers(seqs, k)
        assert got == [[s[i:i + k] for i in range(max(0, len(s) - k + 1))] for s in seqs], k
    with pytest.raises(nt.NtgError):
        ctx.kmers(seqs, 0)
    # strip_returns deletes \r and \n only: a NUL byte is data (round-1 ADVICE)
    out, ch = ctx.strip_returns([b"AC\0GT\r\n", b"\0", b"ACGT"])
    assert out == [b"AC\0GT", b"\0", b"ACGT"] and ch == [True, False, False]
    # quality_mask: sequence and quality lengths must agree per record (sequence.rs:280-297 works on validated records)
    assert ctx.quality_mask([b"ACGT", b"GG"], [b"I!I!", b"!I"], 34) == [b"ANGN", b"NG"]
    with pytest.raises(nt.NtgError):
        ctx.quality_mask([b"ACGT", b"GG"], [b"I!I", b"!I!"], 34)
    with pytest.raises(ValueError):
        ctx.quality_mask([b"ACGT"], [], 34)
    rng = random.Random(17)
    for _ in range(20):
        n = rng.randrange(1, 40)
        ss = [bytes(rng.choice(b"ACGTN") for _ in range(rng.randrange(0, 300))) for _ in range(n)]
        qq = [bytes(rng.randrange(33, 75) for _ in range(len(s))) for s in ss]
        sc = rng.randrange(33, 75)
        assert ctx.quality_mask(ss, qq, sc) == [bytes(78 if q < sc else b for b, q in zip(s, qv)) for s, qv in zip(ss, qq)]


def oracle_items_canonical(seqs, k, rcs=None):
    pos, fl, lo, hi, offs = [], [], [], [], [0]
    for i, s in enumerate(seqs):
        rc = rcs[i] if rcs is not None else None
        for p, kmer, f in O.canonical_kmers(s, k, rc):
            pos.append(p); fl.append(int(f))
            v = 0
            for ch in kmer:
                v = (v << 2) | {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}.get(ch, 0)
            lo.append(v & (2**64 - 1)); hi.append(v >> 64)
        offs.append(len(pos))
    return pos, fl, lo, hi, offs


def check_canonical(ctx, seqs, k, rcs=None):
    it = ctx.canonical_kmers(seqs, k, rcs)
    pos, fl, lo, hi, offs = oracle_items_canonical(seqs, k, rcs)
    assert list(it.item_offs) == offs
    assert list(it.pos) == pos and list(it.was_rc) == fl
    assert [int(x) for x in it.val_lo] == lo
    if k > 32:
        assert [int(x) for x in it.val_hi] == hi


def test_canonical_kmers_vectors_gpu(ctx):  # src/kmer.rs:170-226
    check_canonical(ctx, [b"AGCT"], 1)
    check_canonical(ctx, [b"AGCTA"], 2)
    check_canonical(ctx, [b"AGNTA"], 2)
    check_canonical(ctx, [b"ACGT"], 4)          # tie => was_rc = True
    it = ctx.canonical_kmers([b"AGNTA"], 2)
    assert list(it.pos) == [0, 3]
    rng = random.Random(7)
    seqs = [bytes(rng.choice(b"ACGTacgtNn-") for _ in range(rng.randrange(0, 200))) for _ in range(50)]
    for k in (1, 3, 4, 15, 31, 32, 33, 51, 64):
        check_canonical(ctx, seqs, k)
    # caller-supplied rc buffers (any bytes): CanonicalKmers::new(buffer, rc_buffer, k)
    rcs = [bytes(rng.choice(b"ACGTacgt") for _ in range(len(s))) for s in seqs]
    check_canonical(ctx, seqs, 5, rcs)


def test_bit_kmers_vectors_gpu(ctx):  # src/bitkmer.rs:192-266
    assert list(ctx.bit_kmers([b"AGCT"], 1).val_lo) == [0, 2, 1, 3]
    assert list(ctx.bit_kmers([b"ACNGT"], 2).val_lo) == [0b0001, 0b1011]
    it = ctx.bit_kmers([b"ACGTA"], 3)
    assert list(zip(it.pos, it.val_lo, it.was_rc)) == [(0, 6, 0), (1, 27, 0), (2, 44, 0)]
    assert ctx.bit_kmers([b"TA"], 3).pos.size == 0
    assert list(ctx.bitkmer_reverse_complement([0b000000, 0b111111], 3)) == [0b111111, 0]
    assert list(ctx.bitkmer_reverse_complement([0, 0b00011011], 4)) == [0xFF, 0b00011011]
    assert list(ctx.bitkmer_minimizer([0b001011], 3, 2)) == [0b0010]
    assert list(ctx.bitkmer_minimizer([0b001011], 3, 1)) == [0]
    assert list(ctx.bitkmer_minimizer([0b11000011], 4, 2)) == [0]
    assert list(ctx.bitkmer_minimizer([0b110001, 0b111111], 3, 2)) == [1, 3]
    v, f = ctx.bitkmer_canonical([0b00011011], 4)
    assert (int(v[0]), int(f[0])) == (0b00011011, 0)       # tie => (kmer, false)
    rng = random.Random(11)
    seqs = [bytes(rng.choice(b"ACGTacgtN") for _ in range(rng.randrange(0, 300))) for _ in range(40)]
    for k, canon in ((1, False), (5, True), (21, False), (31, True), (32, True), (32, False)):
        it = ctx.bit_kmers(seqs, k, canon)
        o = 0
        for i, s in enumerate(seqs):
            pos, km, fl = O.bit_kmers(s, k, canon)
            sl = it.of(i)
            assert np.array_equal(it.pos[sl], pos.astype(np.uint32)) and np.array_equal(it.val_lo[sl], km) and np.array_equal(it.was_rc[sl], fl)
    for k, m in ((21, 11), (31, 21), (31, 15), (32, 32), (5, 1)):
        it = ctx.bit_minimizers(seqs, k, m)
        for i, s in enumerate(seqs):
            pos, km, _ = O.bit_kmers(s, k, False)
            exp = np.array([O.bit_minimizer(int(v), k, m) for v in km], dtype=np.uint64)
            sl = it.of(i)
            assert np.array_equal(it.pos[sl], pos.astype(np.uint32)) and np.array_equal(it.val_lo[sl], exp)


def test_invalid_arguments(ctx):
    import needletail_b200 as nt
    for bad in (lambda: ctx.canonical_kmers([b"ACGT"], 0), lambda: ctx.canonical_kmers([b"ACGT"], 65),
                lambda: ctx.bit_kmers([b"ACGT"], 33), lambda: ctx.bit_minimizers([b"ACGT"], 4, 5),
                lambda: ctx.tally(b"@a\nA\n+\nI\n", k=0), lambda: ctx.tally(b"@a\nA\n+\nI\n", k=40, m=3)):
        with pytest.raises(nt.NtgError) as e:
            bad()
        assert e.value.kind == "InvalidArgument"


# ------------------------------------------------------------------ scanner vectors + corpus
SCANNER_VECTORS = [
    b"@test\nAGCT\n+test\n~~a!\n@test2\nTGCA\n+test\nWUI9",
    b"@test\r\nAGCT\r\n+test\r\n~~a!\r\n@test2\r\nTGCA\r\n+test\r\nWUI9",
    b"@test\nACGT\n+\nIII", b"@test\nAGCT\n+test\n~~a!\n@test2\nTGCA", b"@test\nAGCT\n+test\n~~a!\n\n",
    b"@test\nAGCT\n+test\n~~a!\n\n@TEST\nA\n+TEST\n~", b"@\n\n+\n\n@test2\nTGCA\n+test2\n~~~~\n",
    b"@test\nAGCT\n+\nIII\n@TEST\nA\n+\nI", b"@test1\nACGT\n+\nIIII\n@test222\nACGT\n+\nIIII\n@test3\nACGT\n+\nIIII",
    b">test\nACGT\n>test2\nTGCA\n", b">test\nACGT\nACGT\n>test2\nTGCA\nTG", b">test\r\nACGT\r\nACGT\r\n>test2\r\nTGCA\r\nTG",
    b">test\nAGCT\n>test2", b">test\r\nAGCT\r\n>test2\r\n", b">\n\n>shine\nAGGAGGU", b">\r\n\r\n>shine\r\nAGGAGGU",
    b"", b"@", b">", b"Not a valid file", b">id1\nAGTCGTCA", b"@a\nAC\n+\nII\n\n\n", b"@a\nAC\n+\nII\n\n\n\n", b"@a\nAC\n+\nII\n\r\n",
    b"@a\nAC\n+\nII\nX", b"@a\nAC\n+\nII\n@b", b"@a\nAC\n+\nII\n@b\n", b"@a\nAC\n+\nII\n@b\nA\n", b"@a\nAC\n+\nII\n@b\nA\n+\n",
    b"@id with space\nAC\nX\nII\n", b"@a\n\n+\n\n", b"@@\n@@\n+@\n@@\n", b">a\n", b">a\n\n", b">a\nAC\n\n\n", b">a\n>b\nAC", b">a\nAC>b\nGT\n",
    b">a b c\nACGTN\nacgtn\n>\n", b">x\nAC GT\tAC\n",
]


def test_scanner_vectors_gpu(ctx):  # fastq.rs:473-628, fasta.rs:389-482, record.rs:258-285, mod.rs:182-200
    for v in SCANNER_VECTORS:
        assert_parse(ctx, v, what=repr(v))
        for k, m in ((1, 0), (2, 1), (3, 2), (4, 0)):
            assert_tallies(ctx, v, k, m, what=repr(v))


def test_corpus_gpu(ctx):  # tests/format_specimens.rs + tests/data
    fx = load_fixtures()
    for name, data in sorted(fx.items()):
        if name.endswith((".toml", ".gz", ".bz2", ".xz", ".zst")):
            continue
        assert_parse(ctx, data, what=name)
        for k, m in ((4, 0), (21, 11), (31, 21), (31, 0), (32, 22), (51, 0)):
            assert_tallies(ctx, data, k, m, what=name)


def test_pinned_constants_gpu(ctx):  # benches/benchmark.rs:43-44,66-67,97,151 ; lib.rs example (C1)
    fx = load_fixtures()
    t = ctx.tally(fx["data/28S.fasta"], k=31, m=0, iupac=True)
    assert (t["n_records"], t["n_bases"], t["n_kmers"], t["n_not_rc"]) == (570, 738_580, 718_007, 350_983)
    t = ctx.tally(fx["data/28S.fasta"], k=4, query=b"AAAA")
    assert (t["n_kmers"], t["n_not_rc"], t["n_query"]) == (736_277, 350_631, 8_108)
    t = ctx.tally(fx["data/PRJNA271013_head.fq"], k=31, m=21)
    assert (t["n_records"], t["n_bases"], t["n_kmers"], t["n_not_rc"]) == (2000, 250_000, 189_960, 95_997)


# ------------------------------------------------------------------ differential fuzz (mirrors fuzz/fuzz_targets)
def test_fuzz_random_bytes(ctx):
    rng = random.Random(1234)
    alphabet = b"ACGTNacgtn\n\n\r@>+ I!~-.\tU"
    for it in range(300):
        n = rng.randrange(0, 400)
        body = bytes(rng.choice(alphabet) if rng.random() < 0.9 else rng.randrange(256) for _ in range(n))
        data = (b"@" if it % 2 else b">") + body
        assert_parse(ctx, data, what=repr(data))
        assert_tallies(ctx, data, k=rng.choice((1, 2, 3, 5, 8)), m=0, what=repr(data))
        k = rng.choice((2, 3, 5, 8)); m = rng.randrange(1, k + 1)
        assert_tallies(ctx, data, k=k, m=m, what=repr(data))


def mutate_fastq(rng, n_rec, L, crlf=False, n_rate=0.02):
    nl = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n_rec):
        ln = rng.randrange(0, L + 1) if rng.random() < 0.3 else L
        seq = bytes(rng.choice(b"ACGT") if rng.random() > n_rate else rng.choice(b"NnacgtRY-") for _ in range(ln))
        qual = bytes(rng.randrange(33, 75) for _ in range(ln))
        out.append(b"@r%d some text" % i + nl + seq + nl + b"+" + (b"r%d" % i if i % 3 == 0 else b"") + nl + qual + nl)
    return b"".join(out)


def test_fastq_multi_tile_and_errors(ctx):
    rng = random.Random(99)
    data = mutate_fastq(rng, 3000, 150)                       # ~1 MB: many tiles, look-back exercised
    assert len(data) > 10 * 57344
    assert_parse(ctx, data, "fastq lf")
    assert_tallies(ctx, data, 31, 21, what="fastq lf")
    assert_tallies(ctx, data, 21, 11, what="fastq lf")
    assert_tallies(ctx, data, 51, 0, what="fastq lf")
    assert_tallies(ctx, data, 15, 9, what="fastq lf generic window")
    crlf = mutate_fastq(rng, 1200, 100, crlf=True)
    assert_parse(ctx, crlf, "fastq crlf")
    assert_tallies(ctx, crlf, 31, 21, what="fastq crlf")
    assert_tallies(ctx, data[:-1], 31, 21, what="no trailing newline")
    assert_tallies(ctx, data + b"\n\n", 31, 21, what="blank tail")
    # errors deep inside the stream: records before the first error are still tallied
    pos = data.index(b"@r1500 ")
    for bad in (data[:pos] + b"X" + data[pos + 1:],                                   # InvalidStart
                data[:pos] + data[pos:].replace(b"\n+", b"\n-", 1),                   # InvalidSeparator
                data[:pos] + data[pos:].replace(b"\n+\n", b"\n+\nI", 1),              # UnequalLengths
                data[:pos + 40],                                                      # UnexpectedEnd
                data + b"\n\n\n"):                                                    # 3 blank lines => InvalidStart
        assert_parse(ctx, bad, "fastq error")
        got = assert_tallies(ctx, bad, 31, 21, what="fastq error")
        assert got["err_kind"] is not None


def test_long_lines(ctx):
    rng = random.Random(5)
    reads = [bytes(rng.choice(b"ACGT") if rng.random() > 0.001 else 78 for _ in range(L)) for L in (600, 5000, 70000, 130000, 31, 30, 1, 0, 513, 512)]
    fq = b"".join(b"@long%d\n" % i + r + b"\n+\n" + b"I" * len(r) + b"\n" for i, r in enumerate(reads))
    assert_parse(ctx, fq, "long fastq")
    for k, m in ((31, 21), (21, 11), (51, 0), (4, 0)):
        assert_tallies(ctx, fq, k, m, what="long fastq")
    fa = b"".join(b">long%d desc\n" % i + r + b"\n" for i, r in enumerate(reads) if r)
    assert_parse(ctx, fa, "long fasta")
    for k, m in ((31, 21), (21, 11), (51, 0)):
        assert_tallies(ctx, fa, k, m, what="long fasta")
    # wrapped FASTA, 70 columns, LF and CRLF, k-mers span the line breaks
    for nl in (b"\n", b"\r\n"):
        wrapped = b"".join(b">w%d\n" % i + nl.join(r[j:j + 70] for j in range(0, len(r), 70)) + nl for i, r in enumerate(reads) if r)
        wrapped = wrapped.replace(b">w0\n", b">w0" + nl)
        assert_parse(ctx, wrapped, "wrapped fasta")
        for k, m in ((31, 21), (21, 11), (4, 2)):
            assert_tallies(ctx, wrapped, k, m, what="wrapped fasta")


def test_newline_dense_inputs_fall_back(ctx):
    # > NLMAX newlines per tile and whitespace runs longer than the halo: the exact path must take over
    data = b">a\n" + b"A\n" * 60000 + b">b\n" + b"ACGT" + b"\n" * 500 + b"ACGTACGT\n"
    assert_tallies(ctx, data, 4, 2, what="dense newlines fasta")
    assert_tallies(ctx, data, 8, 0, what="dense newlines fasta")
    fq = b"".join(b"@%d\nACGTAC\n+\nIIIIII\n" % i for i in range(20000))
    assert_tallies(ctx, fq, 4, 2, what="short fastq records")
    ws = b">s\nACGT" + b" " * 300 + b"ACGT\n" + b"ACGT" + b"\n" * 200 + b"TTTT\n"
    assert_tallies(ctx, ws, 6, 3, what="whitespace runs")


def test_blank_line_run_after_normal_start(ctx):
    """A normal first 64 KiB (so the tile stays at its full 84 KiB) followed by tens of thousands of blank lines: the
    per-round newline counts of one tile exceed 16 bits (round-1 ADVICE: the packed scan wrapped there).  A blank-line
    run is a valid FASTA body and a valid FASTQ tail."""
    fq_head = O.gen_fastq(0x5EED0002, 0, 220, 150, 0).tobytes()          # 69 520 B of ordinary records
    fa_head = O.gen_fasta(0x5EED0003, 0, 8, 10000, 0).tobytes()          # 80 096 B of unwrapped long reads
    for pad in (0, 1, 7, 255):
        for nblank in (65530, 66000, 70000):
            data = fq_head + b"\n" * (nblank + pad)
            assert_tallies(ctx, data, 31, 21, what=f"fastq + {nblank + pad} blank lines")
            data = fa_head[:-1] + b"\n" * (nblank + pad) + b"ACGTACGTAC\n>x\nAC\n"
            assert_tallies(ctx, data, 8, 4, what=f"fasta + {nblank + pad} blank lines")


# ------------------------------------------------------------------ synthetic generator + full-shape properties
def test_synth_matches_oracle_and_tallies(ctx):
    L, nrec, seed = 150, 20000, 0x5EED0002
    nbytes = nrec * (2 * L + 16)
    d = ctx.device_alloc(nbytes)
    try:
        for n_thresh in (0, 655):
            ctx.synth_fastq_device(d, seed, 12345, nrec, L, n_thresh)
            got = ctx.d2h(d, nbytes)
            exp = O.gen_fastq(seed, 12345, nrec, L, n_thresh)
            assert np.array_equal(got, exp)
            t = ctx.tally_device(d, nbytes, k=31, m=21)
            e = O.tally_fastx(exp.tobytes(), k=31, m=21)
            for key in TALLY_KEYS:
                assert t[key] == e[key], (n_thresh, key)
            assert t["fallback"] == 0                      # the fused single-pass kernel, not the exact re-run
            if n_thresh == 0:
                assert t["n_kmers"] == nrec * (L - 30) and t["n_bases"] == nrec * L
    finally:
        ctx.device_free(d)
    L, nrec, seed = 10000, 300, 0x5EED0003
    nbytes = nrec * (L + 12)
    d = ctx.device_alloc(nbytes)
    try:
        ctx.synth_fasta_device(d, seed, 7, nrec, L, 0)
        got = ctx.d2h(d, nbytes)
        exp = O.gen_fasta(seed, 7, nrec, L, 0)
        assert np.array_equal(got, exp)
        t = ctx.tally_device(d, nbytes, k=21, m=11)
        e = O.tally_fastx(exp.tobytes(), k=21, m=11)
        for key in TALLY_KEYS:
            assert t[key] == e[key], key
        assert t["fallback"] == 0
    finally:
        ctx.device_free(d)


def test_record_owned_fast_path_and_its_fallbacks(ctx):
    """Short-read FASTQ goes through the record-owned kernel (fastq_warp.cuh); what it cannot take falls back to the tile
    kernel with identical results; NTG_TALLY_NO_FASTPATH forces the tile kernel."""
    rng = random.Random(77)
    fq = O.gen_fastq(0x5EED0002, 0, 20000, 150, 655).tobytes()               # 6.3 MB: ~640 chunks
    exp = O.tally_fastx(fq, k=31, m=21)
    got = ctx.tally(fq, k=31, m=21)
    assert got["fast_path"] and got["fallback"] == 0
    for key in TALLY_KEYS:
        assert got[key] == exp[key], key
    ctx.tally_flags = 4
    try:
        slow = ctx.tally(fq, k=31, m=21)
    finally:
        ctx.tally_flags = 0
    assert not slow["fast_path"]
    for key in TALLY_KEYS:
        assert slow[key] == exp[key], key
    # every tail shape, CRLF, other k / m (generic walkers), reads of mixed length
    for data in (fq[:-1], fq + b"\n\n", fq[:-200], fq[:len(fq) // 2 + 77], mutate_fastq(rng, 4000, 150, crlf=True), mutate_fastq(rng, 5000, 120)):
        for k, m in ((31, 21), (21, 11), (15, 9), (51, 0), (31, 0)):
            assert_tallies(ctx, data, k, m, what="fast path shapes")
    # quality lines that start with '@' and '+' everywhere: no unique local evidence in many chunks -> fix-up launch / fallback
    tricky = b"".join(b"@r%d\n" % i + bytes(rng.choice(b"ACGT") for _ in range(100)) + b"\n+\n" + rng.choice((b"@", b"+", b"I")) * 100 + b"\n" for i in range(6000))
    t = assert_tallies(ctx, tricky, 31, 21, what="tricky quality lines")
    assert t["n_records"] == 6000
    # a long read in the middle (longer than the slack) and a phase-shifting error: tile kernel, same answers
    long_rec = b"@long\n" + b"ACGT" * 2000 + b"\n+\n" + b"I" * 8000 + b"\n"
    mixed = fq[: 316 * 500] + long_rec + fq[316 * 500:]
    t = assert_tallies(ctx, mixed, 31, 21, what="long read inside short reads")
    assert not t["fast_path"]
    shifted = fq[: 316 * 700] + fq[316 * 700:].replace(b"\n+\n", b"\n", 1)
    assert_tallies(ctx, shifted, 31, 21, what="a missing separator line")
    assert_parse(ctx, shifted, "a missing separator line")


def test_fast_path_is_taken_on_clean_files(ctx):
    fx = load_fixtures()
    assert ctx.tally(fx["data/28S.fasta"], k=31, m=21)["fallback"] == 0
    assert ctx.tally(fx["data/PRJNA271013_head.fq"], k=31, m=21)["fallback"] == 0
    assert ctx.tally(fx["data/PRJNA271013_head.fq"][:-1], k=51)["fallback"] == 0
    t = ctx.tally(fx["data/bad_header.fastq"], k=4)                                # parse error -> truncated replay, not the exact path
    assert t["fallback"] == 0 and t["err_kind"] is not None
    assert ctx.tally(b">a\n" + b"A\n" * 60000, k=4)["fallback"] & 2               # newline-dense tile


def test_shard_additivity_at_scale(ctx):
    """Size-independent property at a size the oracle cannot cover: tallies over N records equal the
    sum of tallies over two halves (records are independent units), device path vs host-fed path."""
    L, nrec, seed = 150, 2_000_000, 0x5EED0002
    rb = 2 * L + 16
    d = ctx.device_alloc(nrec * rb)
    try:
        ctx.synth_fastq_device(d, seed, 0, nrec, L, 0)
        whole = ctx.tally_device(d, nrec * rb, k=31, m=21)
        half = nrec // 2
        a = ctx.tally_device(d, half * rb, k=31, m=21)
        ctx.synth_fastq_device(d, seed, half, nrec - half, L, 0)
        b = ctx.tally_device(d, (nrec - half) * rb, k=31, m=21)
        for key in TALLY_KEYS[:-1]:
            assert (a[key] + b[key]) % 2**64 == whole[key], key
        assert whole["n_records"] == nrec and whole["n_kmers"] == nrec * (L - 30) and whole["fallback"] == 0
        host = ctx.d2h(d, (nrec - half) * rb)
        h = ctx.tally(host, k=31, m=21)                      # chunk-pipelined host feed
        for key in TALLY_KEYS:
            assert h[key] == b[key], key
    finally:
        ctx.device_free(d)


def test_cpp_host_mirror(tmp_path):
    """The C++ host mirror (needletail.hpp) on the C ABI: the reference's tests re-stated in C++ (README loop on 28S.fasta)."""
    import os, subprocess
    from conftest import ROOT
    exe = str(tmp_path / "test_host_mirror")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe,
                           "-L", os.path.join(ROOT, "needletail_b200"), "-lntgpu", "-Wl,-rpath," + os.path.join(ROOT, "needletail_b200")])
    fa = tmp_path / "28S.fasta"
    fa.write_bytes(load_fixtures()["data/28S.fasta"])
    out = subprocess.run([exe, str(fa)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "738580 bases" in out.stdout and "8108 AAAAs" in out.stdout


def test_non_speculative_path_agrees(ctx):
    """The non-speculative kernel path (NTG_TALLY_NO_SPECULATION: what a mis-speculated call is re-run with) must produce
    the same tallies as the default path and the oracle."""
    fq = O.gen_fastq(0x5EED0004, 0, 30000, 150, 655).tobytes()
    crlf = fq[:316 * 2000].replace(b"\n", b"\r\n")
    datas = (fq, fq[:-1], crlf, load_fixtures()["data/PRJNA271013_head.fq"], fq[:316 * 900] + b"X" + fq[316 * 900 + 1:])
    try:
        for flags in (1, 0):
            ctx.tally_flags = flags
            for data in datas:
                for k, m in ((31, 21), (21, 11), (31, 0), (15, 9), (51, 0)):
                    t = ctx.tally(data, k=k, m=m); e = O.tally_fastx(data, k=k, m=m)
                    assert all(t[key] == e[key] for key in e), (flags, k, m, {key: (t[key], e[key]) for key in e if t[key] != e[key]})
    finally:
        ctx.tally_flags = 0


def test_record_writers(ctx, fixtures):  # src/parser/record.rs:158-247, tests at record.rs:258-293
    import needletail_b200 as nt
    assert nt.write_fasta(b"id", b"ACGT") == b">id\nACGT\n" and nt.write_fasta(b"id", b"ACGT", "windows") == b">id\r\nACGT\r\n"
    assert nt.write_fastq(b"id", b"ACGT", None) == b"@id\nACGT\n+\nIIII\n" and nt.write_fastq(b"i", b"AC", b"!!", "windows") == b"@i\r\nAC\r\n+\r\n!!\r\n"
    rng = random.Random(41)
    for name in ("data/PRJNA271013_head.fq", "data/28S.fasta", "data/test.fa"):
        data = fixtures[name]
        p = ctx.parse(data)
        n = len(p.records)
        for le in (None, "unix", "windows"):
            keep = np.array([rng.random() < 0.6 for _ in range(n)], dtype=np.uint8)
            for kp in (None, keep, np.zeros(n, np.uint8)):
                eol = b"\r\n" if (le == "windows" or (le is None and p.line_ending == "windows")) else b"\n"
                want = []
                for i, r in enumerate(p.records):
                    if kp is not None and not kp[i]:
                        continue
                    idb = data[int(p.table[i, 1]):int(p.table[i, 2])]
                    if p.format == "fasta":
                        want.append(b">" + idb + eol + r.raw_seq + eol)
                    else:
                        want.append(b"@" + idb + eol + r.raw_seq + eol + b"+" + eol + r.qual.encode() + eol)
                assert ctx.write_records(data, p, kp, le) == b"".join(want), (name, le)
    # a FASTQ written back unfiltered with its own line ending parses to the same table of sequences
    fq = fixtures["data/PRJNA271013_head.fq"]
    p = ctx.parse(fq)
    back = ctx.parse(ctx.write_records(fq, p))
    assert [r.seq for r in back.records] == [r.seq for r in p.records] and [r.qual for r in back.records] == [r.qual for r in p.records]


def test_quality_mask_fused_into_the_tally(ctx, fixtures):  # src/sequence.rs:280-297 applied before the per-record loop
    rng = random.Random(53)

    def masked_text(data, score):
        out = bytearray(data)
        for row in O.parse_fastx(bytes(data)).table:
            sb, se, qb, qe = (int(row[i]) for i in (3, 4, 5, 6))
            for j in range(min(se - sb, qe - qb)):
                if data[qb + j] < score:
                    out[sb + j] = ord("N")
        return bytes(out)

    cases = [(O.gen_fastq(0x5EED0002, 0, 3000, 150, 200).tobytes(), True), (fixtures["data/PRJNA271013_head.fq"], None),
             (mutate_fastq(rng, 1500, 120, crlf=True), None),
             (b"@long\n" + bytes(rng.choice(b"ACGT") for _ in range(9000)) + b"\n+\n" + bytes(rng.randrange(33, 75) for _ in range(9000)) + b"\n", False)]
    for data, fast in cases:
        for score in (34, 50, 74):
            want_text = masked_text(data, score)
            for k, m in ((31, 21), (15, 9)):
                exp = O.tally_fastx(want_text, k=k, m=m)
                got = ctx.tally(data, k=k, m=m, qmask=score)
                for key in TALLY_KEYS:
                    assert got[key] == exp[key], (score, k, key, got[key], exp[key])
                if fast is not None:
                    assert got["fast_path"] == fast
        plain = ctx.tally(data, k=31, m=21)
        assert plain["n_kmers"] >= ctx.tally(data, k=31, m=21, qmask=60)["n_kmers"]
    # errors keep their semantics under the mask; FASTA ignores it (no quality)
    fq = cases[0][0]
    bad = fq[: 316 * 900] + b"X" + fq[316 * 900 + 1:]
    got = ctx.tally(bad, k=31, m=21, qmask=50)
    exp = O.tally_fastx(masked_text(bad[: 316 * 900], 50), k=31, m=21)
    assert got["err_kind"] == "InvalidStart" and got["n_records"] == 900 and got["kmer_sum_lo"] == exp["kmer_sum_lo"]
    fa = fixtures["data/28S.fasta"]
    assert ctx.tally(fa, k=31, qmask=50)["kmer_sum_lo"] == ctx.tally(fa, k=31)["kmer_sum_lo"]


def test_record_table_many_tiles(ctx):  # the delimiter index of the record scanner across 8 KiB tiles (src/parser/fasta.rs:102-107,220-243)
    rng = random.Random(77)

    def seqline(n):
        return bytes(rng.choice(b"ACGTNacgt") for _ in range(n))

    # wrapped FASTA: CRLF and LF mixed, '\r' inside headers and inside sequence lines, blank lines, '>' inside lines,
    # records from empty to several tiles long; then the same through small windows (unaligned window starts)
    parts = []
    for r in range(400):
        hdr = b">r%d desc\rwith cr" % r if r % 7 == 0 else b">r%d" % r
        le = b"\r\n" if r % 3 == 0 else b"\n"
        parts.append(hdr + le)
        total = rng.choice((0, 1, 59, 60, 61, 500, 9000, 30000))
        while total > 0:
            w = min(total, rng.choice((60, 70, 1, 8192)))
            line = bytearray(seqline(w))
            if rng.random() < 0.05:
                line[rng.randrange(len(line))] = ord("\r")
            if rng.random() < 0.05:
                line[rng.randrange(len(line))] = ord(">")
            parts.append(bytes(line) + le)
            total -= w
        if r % 11 == 0:
            parts.append(le)
    fa = b"".join(parts)
    assert len(fa) > 40 * 8192
    assert_parse(ctx, fa, "wrapped fasta")
    assert_parse(ctx, fa[:-1], "wrapped fasta, no final newline")
    assert_parse(ctx, fa + b">last", "wrapped fasta, header only at the end")
    exp = O.parse_fastx(fa)
    for window in (100_003, 1 << 20):
        rows = np.concatenate([p.table for p in ctx.parse_chunks(fa, window, with_records=False)])
        et = exp.table[:, :10].copy()
        empty = et[:, 3] == et[:, 4]
        et[empty, 3] = et[empty, 4] = rows[empty, 3] = rows[empty, 4] = 0
        assert np.array_equal(rows, et), window
    # FASTQ across tiles: CRLF, long reads, a quality line starting with '@'
    recs = []
    for r in range(3000):
        L = rng.choice((1, 100, 151, 5000))
        le = b"\r\n" if r % 2 else b"\n"
        q = bytearray(rng.choice(b"@+IJ#") for _ in range(L))
        recs.append(b"@q%d" % r + le + seqline(L) + le + b"+" + le + bytes(q) + le)
    fq = b"".join(recs)
    assert_parse(ctx, fq, "fastq tiles")
    assert_parse(ctx, fq[:-2], "fastq tiles, no final newline")
    assert_parse(ctx, fq[: len(fq) // 2], "fastq tiles, truncated")
    exp = O.parse_fastx(fq)
    rows = np.concatenate([p.table for p in ctx.parse_chunks(fq, 70_001, with_records=False)])
    assert np.array_equal(rows, exp.table[:, :10])
