import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import needletail_b200 as nt, oracle_lib as O
from conftest import load_fixtures
fx = load_fixtures()
ctx = nt.Context(0)
bad = 0
for name in ("data/28S.fasta", "data/PRJNA271013_head.fq", "data/test.fa"):
    data = fx[name]
    for k, m in ((4, 0), (21, 11), (31, 21), (31, 0), (32, 22), (51, 0)):
        try:
            t = ctx.tally(data, k=k, m=m)
        except Exception as e:
            print(name, k, m, "EXC", e); bad += 1; continue
        e = O.tally_fastx(bytes(data), k=k, m=m)
        diff = {key: (t[key], e[key]) for key in e if t[key] != e[key]}
        if diff: print(name, k, m, diff, "fallback", t.get("fallback")); bad += 1
print("bad", bad)
