import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import needletail_b200 as nt
ctx = nt.Context(0)
L, nrec = 150, 700_000
nbytes = nrec * (2 * L + 16)
def run(tag):
    d = ctx.device_alloc(nbytes)
    ctx.synth_fastq_device(d, 0x5EED0002, 0, nrec, L, 655)
    t = ctx.tally_device(d, nbytes, k=31, m=21)
    print(tag, {k: t[k] for k in ("n_records", "err_kind", "err_line", "fallback", "ws_handover")}, flush=True)
    ctx.device_free(d)
run("fresh")
s = ctx.stream(k=5); s.feed(b">a\nACGTACGT\n"); print(s.finish()["n_records"])
run("after 1 stream")
for i in range(70):
    s = ctx.stream(k=5); s.feed(b">a\nACGTACGT\n"); s.finish()
run("after 71 streams")
for i in range(70):
    ctx.tally(b">a\nACGTACGT\n", k=5)
run("after 70 host tallies")
