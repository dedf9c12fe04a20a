#!/usr/bin/env python3
"""A small invocation of every fused-path shape for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
FASTQ (speculative + deferred look-back, then without speculation), FASTA (general look-back), long lines, an error replay,
a streamed session; each checked against the oracle."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import needletail_b200 as nt
import oracle_lib as O

rng = random.Random(1)
ctx = nt.Context(0)
KEYS = ("n_records", "n_bases", "n_kmers", "n_not_rc", "kmer_sum_lo", "kmer_sum_hi", "n_minimizers", "minimizer_sum", "err_kind")
fq = O.gen_fastq(0x5EED0002, 0, 3000, 150, 655).tobytes()              # ~11 tiles
fa = O.gen_fasta(0x5EED0003, 0, 60, 10000, 0).tobytes()                # long lines, ~8 tiles
cases = [("fastq", fq, 31, 21), ("fastq k51", fq, 51, 0), ("fasta", fa, 21, 11), ("fastq truncated", fq[:-100], 31, 21),
         ("fastq generic", fq, 15, 9)]
for name, data, k, m in cases:
    exp = O.tally_fastx(data, k=k, m=m)
    got = ctx.tally(data, k=k, m=m)
    assert all(got[x] == exp[x] for x in KEYS), (name, got, exp)
    s = ctx.stream(k=k, m=m); s.feed(data[:100000]); s.feed(data[100000:]); got = s.finish()
    assert all(got[x] == exp[x] for x in KEYS), (name, "stream")
# record scanner (delimiter index + one thread per record), whole buffer and windows
for name, data in (("fastq", fq), ("fasta", fa), ("fasta crlf", fa.replace(b"\n", b"\r\n")), ("fastq cut", fq[:-100])):
    exp = O.parse_fastx(data)
    got = ctx.parse(data)
    assert got.err_kind == exp.err_kind and len(got.records) == len(exp.records), name
    assert (got.table[:, 8] == exp.table[:, 8]).all() and (got.table[:, 7] == exp.table[:, 7]).all(), name
    rows = sum(len(p.table) for p in ctx.parse_chunks(data, 50_001, with_records=False))
    assert rows == len(exp.records), name
ctx.tally_flags = 1                                                     # NTG_TALLY_NO_SPECULATION
got = ctx.tally(fq, k=31, m=21); exp = O.tally_fastx(fq, k=31, m=21)
assert all(got[x] == exp[x] for x in KEYS)
ctx.close()
print("sanitize_small ok")
