import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import needletail_b200 as nt
import oracle_lib as O
from needletail_b200 import bgzf
ctx = nt.Context(0)
text = O.gen_fastq(0x5EED0005, 0, 65000, 250, 0, nthreads=8).tobytes()
for nm in (8, 64, 512):
    t = text[: nm * 0xFF00]
    blob = bgzf.compress(t, 1)
    ctx.inflate_bgzf(blob)
    t0 = time.perf_counter(); out = ctx.inflate_bgzf(blob); dt = time.perf_counter() - t0
    assert out == t
    print(f"{nm} members, {len(t)/1e6:.1f} MB text, {len(blob)/1e6:.1f} MB compressed: {dt*1e3:.1f} ms -> {len(t)/dt/1e6:.0f} MB/s", flush=True)
