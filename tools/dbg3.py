import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import needletail_b200 as nt
import oracle_lib as O
ctx = nt.Context(0)
fq = O.gen_fastq(0x5EED0002, 0, 3000, 150, 655).tobytes()
for name, data in (("clean", fq), ("no trailing newline", fq[:-1]), ("blank tail", fq + b"\n\n"), ("truncated", fq[:-100])):
    exp = O.tally_fastx(data, k=31, m=21)
    got = ctx.tally(data, k=31, m=21)
    print(name, all(got[x] == exp[x] for x in exp), got["fallback"], flush=True)
