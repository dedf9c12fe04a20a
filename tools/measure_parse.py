#!/usr/bin/env python3
"""Throughput of the incremental record scanner (ntg_parse_fastx_chunk) over host text, window by window: GB/s of input and
records per second, rows checked against arithmetic expectation.  Not a bench line: a side number for DESIGN.md."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import needletail_b200 as nt

ctx = nt.Context(0)
L, nrec = 150, 12_000_000                                   # 3.8 GB of FASTQ
rb = 2 * L + 16
d = ctx.device_alloc(nrec * rb)
ctx.synth_fastq_device(d, 0x5EED0002, 0, nrec, L, 0)
host = ctx.d2h(d, nrec * rb)
ctx.device_free(d)
for window in (256 << 20, 1 << 30):
    t0 = time.perf_counter()
    rows = 0
    for p in ctx.parse_chunks(host, window, with_records=False):
        rows += len(p.table)
        assert p.err_kind is None
        if len(p.table):
            assert int(p.table[-1, 9]) == 4 * (rows - 1) + 1      # line of the last row
    dt = time.perf_counter() - t0
    assert rows == nrec
    print(f"window {window >> 20} MiB: {rows} rows in {dt:.2f} s = {nrec * rb / dt / 1e9:.2f} GB/s of text, {rows / dt / 1e6:.1f} M records/s (H2D + 3-pass scanner + rows D2H)")
ctx.close()
