"""GPU parity of the k-mer spectrum (ntg_spectrum_*): the multiset of canonical k-mers of the records the reference's iterator
delivers, against the oracle's per-record loop (normalize(false) -> bit_kmers(k, canonical))."""
import collections
import random

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import needletail_b200 as nt
    c = nt.Context(0)
    yield c
    c.close()


def oracle_spectrum(data, k):
    cnt = collections.Counter()
    for r in O.parse_fastx(data).records:
        seq = r["raw_seq"].replace(b"\n", b"").replace(b"\r", b"")
        norm = O.normalize(seq, False)[0]
        _, km, _ = O.bit_kmers(norm, k, True)
        cnt.update(int(v) for v in km)
    return cnt


def check(sp, want):
    keys, counts = sp.items()
    assert len(keys) == len(want), (len(keys), len(want))
    assert dict(zip((int(x) for x in keys), (int(c) for c in counts))) == dict(want)
    h = sp.histogram(16)
    coc = collections.Counter(min(c, 15) for c in want.values())
    assert [int(x) for x in h] == [coc.get(i, 0) for i in range(16)]


def test_c1_readme_example(ctx, fixtures):
    """BASELINE config C1: tests/data/28S.fasta, canonical k=4, count of AAAA (8 108; k-mers 736 277)."""
    data = fixtures["data/28S.fasta"]
    sp = ctx.spectrum(4)
    t = sp.add(data)
    assert t["n_records"] == 570 and t["n_kmers"] == 736277 == sp.kmers_added()
    assert sp.count(b"AAAA") == 8108 == sp.count(b"TTTT")
    check(sp, oracle_spectrum(data, 4))
    sp.close()


def test_dense_and_hash_match_oracle(ctx, fixtures):
    rng = random.Random(3)
    fq = fixtures["data/PRJNA271013_head.fq"]
    fa = fixtures["data/28S.fasta"]
    for data, ks in ((fq, (1, 5, 11, 14, 15, 21, 31, 32)), (fa[:200000], (7, 16, 32))):
        for k in ks:
            want = oracle_spectrum(data, k)
            sp = ctx.spectrum(k, capacity=4 * len(want) + 1024)
            t = sp.add(data)
            assert t["n_kmers"] == sum(want.values())
            check(sp, want)
            # adding a second input accumulates; resident input == host input
            d = ctx.device_alloc(len(data) + 16)
            ctx.h2d(d, np.frombuffer(data, dtype=np.uint8))
            sp.add_device(d, len(data))
            ctx.device_free(d)
            check(sp, collections.Counter({key: 2 * c for key, c in want.items()}))
            some = rng.sample(sorted(want), min(5, len(want)))
            for key in some:
                kmer = bytes(b"ACGT"[(key >> (2 * (k - 1 - i))) & 3] for i in range(k))
                assert sp.count(kmer) == 2 * want[key]
                assert sp.count(O.reverse_complement(kmer)) == 2 * want[key]
            sp.clear()
            assert sp.items()[0].size == 0
            sp.close()


def test_parse_error_and_full_table(ctx, fixtures):
    import needletail_b200 as nt
    fq = fixtures["data/PRJNA271013_head.fq"]
    pos = fq.index(b"\n@", len(fq) // 2) + 1
    bad = fq[:pos] + b"X" + fq[pos + 1:]                                   # InvalidStart in the middle: records before it count
    want = oracle_spectrum(bad, 21)
    assert sum(want.values()) > 0
    sp = ctx.spectrum(21, capacity=1 << 20)
    t = sp.add(bad)
    assert t["err_kind"] is not None and t["n_kmers"] == sum(want.values())
    check(sp, want)
    sp.close()
    sp = ctx.spectrum(21, capacity=1024)                                   # far too small
    with pytest.raises(nt.NtgError):
        sp.add(fq)
    sp.close()
    with pytest.raises(nt.NtgError):
        ctx.spectrum(33)


def test_spectrum_at_scale_sums(ctx):
    """2M synthetic reads, k = 12 (dense) and k = 31 (hash, ~240M distinct would not fit a test: 200k reads): totals agree with the tallies"""
    L, nrec = 150, 2_000_000
    nb = nrec * (2 * L + 16)
    d = ctx.device_alloc(nb)
    ctx.synth_fastq_device(d, 0x5EED0002, 0, nrec, L, 655)
    sp = ctx.spectrum(12)
    t = sp.add_device(d, nb)
    ref = ctx.tally_device(d, nb, k=12)
    assert t["n_kmers"] == ref["n_kmers"] and t["kmer_sum_lo"] == ref["kmer_sum_lo"]
    keys, counts = sp.items()
    assert int(counts.astype(np.uint64).sum()) == ref["n_kmers"]
    assert int((keys * counts.astype(np.uint64)).sum() & np.uint64(0xFFFFFFFFFFFFFFFF)) == ref["kmer_sum_lo"]      # checksum from the spectrum
    sp.close()
    n31 = 200_000 * (2 * L + 16)
    sp = ctx.spectrum(31, capacity=1 << 26)
    t = sp.add_device(d, n31)
    ref = ctx.tally_device(d, n31, k=31)
    keys, counts = sp.items()
    assert int(counts.astype(np.uint64).sum()) == ref["n_kmers"] == t["n_kmers"]
    with np.errstate(over="ignore"):
        assert int((keys * counts.astype(np.uint64)).sum()) == ref["kmer_sum_lo"]
    sp.close()
    ctx.device_free(d)
