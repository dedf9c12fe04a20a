#!/usr/bin/env python3
"""Throughput of the `Sequence` batch calls of the C ABI (materialising paths) on a batch of 150 bp reads held in host memory:
H2D + kernels + results D2H, per call.  Not a bench line: a side table for DESIGN.md (these paths return 13 bytes per k-mer to the
host, so they are bound by the copy back; the tally path is the one that never materialises)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import needletail_b200 as nt
from needletail_b200 import _Items

ctx = nt.Context(0)
lib, h = ctx.lib, ctx.h
nseq, L = int(os.environ.get("NT_SEQOPS_READS", "2000000")), 150
rng = np.random.default_rng(7)
cat = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=nseq * L, dtype=np.uint8)]
cat[rng.integers(0, cat.size, size=nseq // 4)] = ord("N")                     # a non-ACGT base in about every fourth read
offs = (np.arange(nseq + 1, dtype=np.uint64) * np.uint64(L))


def timed(name, fn, items_of=None, reps=3):
    best, n_items = 1e9, 0
    for _ in range(reps):
        t0 = time.perf_counter()
        n_items = fn()
        best = min(best, time.perf_counter() - t0)
    extra = f", {n_items / best / 1e6:.0f} M items/s ({n_items} items)" if n_items else ""
    print(f"{name:42s} {best * 1e3:8.1f} ms = {cat.size / best / 1e9:6.2f} Gbases/s of input{extra}", flush=True)


def items_call(fn, *args):
    def run():
        out = C.POINTER(_Items)()
        ctx._ck(fn(h, *args, C.byref(out)))
        n = int(out.contents.n_items)
        lib.ntg_items_free(out)
        return n
    return run


out = np.empty(cat.size, dtype=np.uint8); ooffs = np.zeros(nseq + 1, dtype=np.uint64); ch = np.zeros(nseq, dtype=np.uint8)
timed("ntg_normalize", lambda: ctx._ck(lib.ntg_normalize(h, cat.ctypes.data, offs.ctypes.data, nseq, 0, out.ctypes.data, ooffs.ctypes.data, ch.ctypes.data)) or 0)
timed("ntg_reverse_complement", lambda: ctx._ck(lib.ntg_reverse_complement(h, cat.ctypes.data, offs.ctypes.data, nseq, out.ctypes.data)) or 0)
timed("ntg_canonical_kmers k=31", items_call(lib.ntg_canonical_kmers, cat.ctypes.data, None, offs.ctypes.data, nseq, 31))
timed("ntg_bit_kmers k=31 canonical", items_call(lib.ntg_bit_kmers, cat.ctypes.data, offs.ctypes.data, nseq, 31, 1))
timed("ntg_bit_minimizers k=31 m=21", items_call(lib.ntg_bit_minimizers, cat.ctypes.data, offs.ctypes.data, nseq, 31, 21))
ctx.close()
