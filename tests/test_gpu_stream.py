"""GPU parity of the streaming feeder (ntg_stream_*, ntg_tally_fastx over segments, gzip in front, the chunked record scanner)
and of the error replay (records before the first error are tallied, the error is reported with the reference's kind / line /
id).  Everything goes through the C ABI; the checker is the CPU oracle (tests/oracle_lib.py)."""
import gzip
import os
import random

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

TALLY_KEYS = ["n_records", "n_bases", "n_kmers", "n_not_rc", "kmer_sum_lo", "kmer_sum_hi", "n_query",
              "n_minimizers", "minimizer_sum", "err_kind"]


@pytest.fixture(scope="module")
def ctx():
    import needletail_b200 as nt
    c = nt.Context(0)
    yield c
    c.close()


def same(got, exp, what, keys=TALLY_KEYS):
    for key in keys:
        assert got[key] == exp[key], f"{what}: {key}: {got[key]} != {exp[key]}"


def stream_tally(ctx, data, k, m, pieces):
    s = ctx.stream(k=k, m=m)
    off = 0
    for p in pieces:
        s.feed(data[off:off + p])
        off += p
    if off < len(data):
        s.feed(data[off:])
    return s.finish()


def fastq(rng, n_rec, L=100, nl=b"\n"):
    out = []
    for i in range(n_rec):
        seq = bytes(rng.choice(b"ACGT") for _ in range(L))
        out.append(b"@r%d x" % i + nl + seq + nl + b"+" + nl + bytes(rng.randrange(33, 75) for _ in range(L)) + nl)
    return b"".join(out)


# ------------------------------------------------------------------ error replay (whole-buffer entry points)
def test_error_kind_line_id_match_oracle(ctx):
    rng = random.Random(7)
    data = fastq(rng, 4000, 120)                                   # ~1 MB, a dozen tiles
    recs = [i for i in range(len(data)) if data[i:i + 2] == b"@r" and (i == 0 or data[i - 1] == 10)]
    cases = []
    for r in (0, 1, 7, 1999, 2000, 3998, 3999):
        pos = recs[r]
        end = recs[r + 1] if r + 1 < len(recs) else len(data)
        rec = data[pos:end]
        cases.append(data[:pos] + b"X" + data[pos + 1:])                                     # InvalidStart
        cases.append(data[:pos] + rec.replace(b"\n+\n", b"\n-\n", 1) + data[end:])             # InvalidSeparator
        cases.append(data[:pos] + rec.replace(b"\n+\n", b"\n+\nI", 1) + data[end:])            # UnequalLengths (qual longer)
        cases.append(data[:pos] + rec[:-2] + b"\n" + data[end:])                               # UnequalLengths (qual shorter)
        cases.append(data[:pos + len(rec) // 2])                                               # UnexpectedEnd / truncated
        cases.append(data[:pos] + rec[:rec.index(b"\n") + 1])                                  # header line only
        cases.append(data[:pos] + rec[:-1])                                                    # no trailing newline: valid
    for i, bad in enumerate(cases):
        exp = O.tally_fastx(bad, k=31, m=21)
        pexp = O.parse_fastx(bad)
        for label, got in (("host", ctx.tally(bad, k=31, m=21)),
                           ("stream", stream_tally(ctx, bad, 31, 21, [len(bad) // 3, 1, 70000]))):
            same(got, exp, f"case {i} {label}")
            assert got["err_kind"] == pexp.err_kind, (i, label)
            if pexp.err_kind is not None:
                assert got["err_line"] == pexp.err_line, (i, label, got["err_line"], pexp.err_line)
        d = ctx.device_alloc(len(bad) + 16)
        ctx.h2d(d, np.frombuffer(bad, dtype=np.uint8))
        same(ctx.tally_device(d, len(bad), k=31, m=21), exp, f"case {i} resident")
        ctx.device_free(d)


def test_fasta_end_of_stream_error(ctx):
    for data in (b">a\nACGTACGTAC\n>b\nACGTTTGA\n>c", b">a\nACGTACGTAC\n>b\nACGTTTGA\n>c\n", b">a\nACGTACGTAC\n>b x y\n", b">a"):
        exp = O.tally_fastx(data, k=4, m=2)
        pexp = O.parse_fastx(data)
        for label, got in (("host", ctx.tally(data, k=4, m=2)), ("stream", stream_tally(ctx, data, 4, 2, [3, 5]))):
            same(got, exp, f"{data!r} {label}")
            assert got["err_kind"] == pexp.err_kind == "UnexpectedEnd"
            assert got["err_line"] == pexp.err_line, (data, label)
            assert got["fallback"] == 0, "the single pass handles the end-of-stream rule itself"


# ------------------------------------------------------------------ streaming sessions
def test_stream_small_inputs_and_sniff(ctx):
    import needletail_b200 as nt
    rng = random.Random(3)
    for data in (b"", b"@", b">", b"ACGT\n", b">s\nACGT\n", b"@s\nACGT\n+\nIIII\n", fastq(rng, 50)):
        exp = O.tally_fastx(data, k=3, m=2)
        for pieces in ([], [1], [1, 1, 1], [5, 0, 2]):
            same(stream_tally(ctx, data, 3, 2, pieces), exp, f"{data[:20]!r} {pieces}")
    with pytest.raises(nt.NtgError):
        s = ctx.stream(k=0)
    s = ctx.stream(k=5)
    s.feed(b">a\nACGTACGT\n")
    s.finish()
    with pytest.raises(nt.NtgError):
        s.feed(b"x")                                                # closed


def test_stream_multi_segment_matches_resident(ctx):
    """~210 MB of synthetic FASTQ: three 64 MiB segments + a partial one; pieces of awkward sizes; the look-back state and the
    history bytes carry across launches.  The reference point is the resident single launch over the same bytes."""
    L, nrec, seed = 150, 700_000, 0x5EED0002
    nbytes = nrec * (2 * L + 16)
    d = ctx.device_alloc(nbytes)
    ctx.synth_fastq_device(d, seed, 0, nrec, L, 655)
    exp = ctx.tally_device(d, nbytes, k=31, m=21)
    assert exp["n_records"] == nrec and exp["err_kind"] is None and exp["fallback"] == 0
    host = ctx.d2h(d, nbytes)
    ctx.device_free(d)
    got = ctx.tally(host, k=31, m=21)                               # ntg_tally_fastx: segments straight from caller memory
    same(got, exp, "host segments")
    assert got["fallback"] == 0
    rng = random.Random(1)
    pieces = []
    while sum(pieces) < nbytes:
        pieces.append(rng.choice((1, 4097, 1 << 20, 33 << 20, 70 << 20)))
    got = stream_tally(ctx, host, 31, 21, pieces)
    same(got, exp, "stream session")
    assert got["fallback"] == 0
    # FASTA long reads through the session (general look-back state across launches)
    Lf, nf = 10_000, 15_000
    nb = nf * (Lf + 12)
    d = ctx.device_alloc(nb)
    ctx.synth_fasta_device(d, 0x5EED0003, 0, nf, Lf, 0)
    exp = ctx.tally_device(d, nb, k=21, m=11)
    hostf = ctx.d2h(d, nb)
    ctx.device_free(d)
    same(stream_tally(ctx, hostf, 21, 11, [50 << 20, 50 << 20]), exp, "fasta stream")
    # an error in the record that straddles the first segment boundary, and one deep in the third segment
    tb_probe = ctx.tally(host[: 1 << 20], k=31, m=21)
    assert tb_probe["err_kind"] in (None, "UnexpectedEnd")
    rec = 2 * L + 16
    for target in (64 << 20, (64 << 20) + 40000, 150 << 20):
        r = target // rec
        bad = host.copy()
        bad[r * rec + 12 + L + 1] = ord("-")                        # the separator of record r
        expb = ctx.tally(bad[: r * rec], k=31, m=21)               # everything before the failing record
        assert expb["n_records"] == r and expb["err_kind"] is None
        for label, got in (("host", ctx.tally(bad, k=31, m=21)), ("stream", stream_tally(ctx, bad, 31, 21, [1 << 20] * 10 + [90 << 20]))):
            same(got, expb, f"error at record {r} {label}", keys=TALLY_KEYS[:-1])
            assert got["err_kind"] == "InvalidSeparator" and got["err_line"] == 4 * r + 3, (label, got["err_kind"], got["err_line"])
        bad[r * rec + 12 + L + 1] = ord("+")
        # truncated inside record r: in the sequence line (UnexpectedEnd), in the quality line (a last record without newline
        # whose lengths differ: fastq.rs:337-343)
        for cut, kind in ((100, "UnexpectedEnd"), (200, "UnequalLengths")):
            bad2 = host[: r * rec + cut]
            for label, got in (("host", ctx.tally(bad2, k=31, m=21)), ("stream", stream_tally(ctx, bad2, 31, 21, [60 << 20]))):
                same(got, expb, f"truncated in record {r} {label}", keys=TALLY_KEYS[:-1])
                assert got["err_kind"] == kind, (label, cut, got["err_kind"])


# ------------------------------------------------------------------ gzip in front
def test_gzip_members_bgzf_and_file(ctx, tmp_path):
    from needletail_b200 import bgzf
    import needletail_b200 as nt
    rng = random.Random(11)
    data = fastq(rng, 6000, 150)
    exp = O.tally_fastx(data, k=31, m=21)
    one = gzip.compress(data, 1)
    two = gzip.compress(data[:400_000], 1) + gzip.compress(data[400_000:], 6)      # multi-member == MultiGzDecoder (mod.rs:98)
    bg = bgzf.compress(data)
    for label, blob, threads in (("gzip", one, 1), ("two members", two, 1), ("bgzf sequential", bg, 1), ("bgzf 4 threads", bg, 4),
                                 ("gzip asked for threads", one, 4), ("padded", one + b"\0" * 512, 1)):
        for cut in (None, 1, 77777):
            s = ctx.stream(k=31, m=21)
            if cut is None:
                s.feed_gz(blob, threads)
            else:
                for o in range(0, len(blob), cut if cut > 1 else max(1, len(blob) // 50)):
                    s.feed_gz(blob[o:o + (cut if cut > 1 else max(1, len(blob) // 50))], threads)
            same(s.finish(), exp, f"{label} cut={cut}")
    # truncated compressed streams are I/O errors (ParseErrorKind::Io, errors.rs:144-153)
    for blob, threads in ((one[:-100], 1), (bg[:len(bg) // 2 + 5], 4), (one[:50], 1)):
        s = ctx.stream(k=31, m=21)
        s.feed_gz(blob, threads)
        assert s.finish()["err_kind"] == "Io"
    # files: plain, gzip, BGZF, empty, empty gzip (tests/test_compressed.rs, mod.rs:204-253)
    for name, blob, kind in (("a.fq", data, None), ("a.fq.gz", one, None), ("a.bgz", bg, None), ("empty", b"", "EmptyFile"),
                             ("empty.gz", gzip.compress(b""), "EmptyFile"), ("one.gz", gzip.compress(b">"), "EmptyFile")):
        p = tmp_path / name
        p.write_bytes(blob)
        got = ctx.tally_file(str(p), k=31, m=21, threads=3)
        if kind is None:
            same(got, exp, name)
        else:
            assert got["err_kind"] == kind, (name, got["err_kind"])
    assert ctx.tally_file(str(tmp_path / "missing"), k=31)["err_kind"] == "Io"
    p = tmp_path / "a.bz2"
    p.write_bytes(b"BZh91AY&SY")
    with pytest.raises(nt.NtgError):
        ctx.tally_file(str(p), k=31)


# ------------------------------------------------------------------ chunked record scanner
def test_parse_chunks_equal_whole_parse(ctx, fixtures):
    rng = random.Random(5)
    fq = fastq(rng, 3000, 90) + b"@last\nACGT\n+\nIIII"              # last record without newline
    fa = b"".join(b">s%d d\n" % i + b"\n".join(bytes(rng.choice(b"ACGTN") for _ in range(60)) for _ in range(rng.randrange(1, 9))) + b"\n"
                  for i in range(2500))
    for data in (fq, fa, fq[:70000] + b"X" + fq[70001:], fixtures["data/28S.fasta"], fixtures["data/PRJNA271013_head.fq"]):
        whole = ctx.parse(data)
        for window in (40_000, 333_333):
            recs = list(ctx.parse_chunks(data, window))
            got = recs[-1]
            table = np.concatenate([r.table for r in recs if len(r.table)]) if any(len(r.table) for r in recs) else np.zeros((0, 10), np.uint64)
            assert got.err_kind == whole.err_kind
            assert len(table) == len(whole.table)
            assert np.array_equal(table, whole.table)
            if whole.err_kind:
                assert (got.err_line, got.err_id) == (whole.err_line, whole.err_id)


# ------------------------------------------------------------------ DEFLATE on the device (BGZF)
def bgzf_with(data, level, strategy=0, block=0xFF00):
    """BGZF members with a chosen zlib level / strategy (0 = stored blocks, Z_FIXED = fixed Huffman codes, ...)."""
    import struct, zlib
    out = []
    for o in range(0, len(data), block):
        chunk = bytes(data[o:o + block])
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        payload = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(payload) + 8
        assert bsize <= 0x10000
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + payload
                   + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


def test_device_inflate_matches_zlib(ctx):
    import zlib
    import needletail_b200 as nt
    from needletail_b200 import bgzf
    rng = random.Random(23)
    texts = [fastq(rng, 3000, 150), bytes(rng.randrange(256) for _ in range(200_000)), b"A" * 300_000, b"", b"x",
             bytes(rng.choice(b"ACGT") for _ in range(150_000)), (b"ACGTTGCA" * 40 + b"\n") * 800]
    for ti, text in enumerate(texts):
        for level, strategy, block in ((1, 0, 0xFF00), (6, 0, 0xFF00), (9, 0, 0x8000), (0, 0, 0xFF00 - 64), (6, zlib.Z_FIXED, 0xFF00),
                                       (6, zlib.Z_HUFFMAN_ONLY, 0x4000), (6, zlib.Z_RLE, 0xFF00), (1, 0, 100)):
            blob = bgzf_with(text, level, strategy, block) + bgzf.EOF_MARKER
            assert ctx.inflate_bgzf(blob) == text, (ti, level, strategy, block)
    # corrupt payloads are I/O errors, never out-of-bounds writes
    blob = bytearray(bgzf_with(texts[0], 6))
    for pos in (30, 200, 5000, len(blob) // 2):
        bad = bytearray(blob); bad[pos] ^= 0x5A
        try:
            got = ctx.inflate_bgzf(bytes(bad))
        except nt.NtgError:
            continue
        assert len(got) == len(texts[0])                          # (a flipped bit may still decode to the right length)


def test_device_inflate_in_the_tally_session(ctx):
    from needletail_b200 import bgzf
    rng = random.Random(29)
    data = fastq(rng, 9000, 150)
    exp = O.tally_fastx(data, k=31, m=21)
    bg = bgzf.compress(data)
    for cut in (None, 1 << 16, 4093):
        s = ctx.stream(k=31, m=21)
        if cut is None:
            s.feed_gz(bg, 0)
        else:
            for o in range(0, len(bg), cut):
                s.feed_gz(bg[o:o + cut], 0)
        got = s.finish()
        same(got, exp, f"device inflate cut={cut}")
    # errors inside the text keep their iterator semantics
    pos = data.index(b"@r4500 ")
    bad = data[:pos] + data[pos:].replace(b"\n+\n", b"\n-\n", 1)
    s = ctx.stream(k=31, m=21); s.feed_gz(bgzf.compress(bad), 0); got = s.finish()
    same(got, O.tally_fastx(bad, k=31, m=21), "device inflate + parse error")
    assert got["err_line"] == O.parse_fastx(bad).err_line
    # tiny and empty texts, FASTA, unknown format, truncated file
    for text in (b"", b">", b">a\nACGT\n", b"hello\n"):
        s = ctx.stream(k=3); s.feed_gz(bgzf.compress(text), 0)
        same(s.finish(), O.tally_fastx(text, k=3), repr(text))
    s = ctx.stream(k=31); s.feed_gz(bg[:len(bg) // 2], 0)
    assert s.finish()["err_kind"] == "Io"
    # a few hundred MB: several 512 MiB-text batches are not needed to cross launches — force it with many members
    L, nrec = 150, 2_000_000
    d = ctx.device_alloc(nrec * (2 * L + 16))
    ctx.synth_fastq_device(d, 0x5EED0002, 0, nrec, L, 0)
    exp = ctx.tally_device(d, nrec * (2 * L + 16), k=31, m=21)
    host = ctx.d2h(d, nrec * (2 * L + 16)); ctx.device_free(d)
    blob = bgzf.compress(host[: 64 << 20].tobytes())                # 64 MiB of text compressed once ...
    reps = (nrec * (2 * L + 16)) // (64 << 20)
    s = ctx.stream(k=31, m=21)
    for _ in range(reps * 9):                                        # ... fed often enough for two 512 MiB batches
        s.feed_gz(blob[:-28], 0)                                     # (without the EOF marker)
    got = s.finish()
    one = ctx.tally(host[: 64 << 20], k=31, m=21)
    # the 64 MiB block ends inside a record, so the concatenation is not 9*reps copies of valid text: compare with the host inflate
    s = ctx.stream(k=31, m=21)
    for _ in range(reps * 9):
        s.feed_gz(blob[:-28], 4)
    ref = s.finish()
    same(got, ref, "device vs host inflate, multi-batch")
    assert one["n_records"] > 0


def test_release_scratch_between_calls(ctx):  # ntg_release_scratch: the cached scratch can be handed back at any point
    fq = O.gen_fastq(0x5EED0002, 0, 2000, 150, 0).tobytes()
    exp = O.parse_fastx(fq)
    for _ in range(2):
        got = ctx.parse(fq)
        assert len(got.records) == len(exp.records) and np.array_equal(got.table, exp.table[:, :10])
        it = ctx.bit_kmers([fq[12:162]], 31, True)
        assert len(it.pos) == 120
        ctx.release_scratch()
