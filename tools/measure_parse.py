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
import ctypes as C
# the same text in pinned host memory (what a feeder thread would hand over): H2D at PCIe speed instead of the driver's staging copy
hp = C.c_void_p()
assert ctx.lib.ntg_alloc_pinned(host.size, C.byref(hp)) == 0
pinned = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(host.size,))
pinned[:] = host
for label, src in (("pageable", host), ("pinned", pinned)):
    for window in (256 << 20, 1 << 30):
        for rep in range(2):                                  # (the second pass runs on the context's warm scratch buffers)
            ctx.parse_seconds = 0.0
            t0 = time.perf_counter()
            rows = 0
            for p in ctx.parse_chunks(src, window, with_records=False):
                rows += len(p.table)
                assert p.err_kind is None
                if len(p.table):
                    assert int(p.table[-1, 9]) == 4 * (rows - 1) + 1      # line of the last row
                    assert int(p.table[-1, 7]) == rows * rb - 1           # all_e of the last row: arithmetic expectation
            dt = time.perf_counter() - t0
            assert rows == nrec
        print(f"{label} host text, window {window >> 20} MiB: {rows} rows; inside the C ABI (H2D + index + rows + table D2H) "
              f"{ctx.parse_seconds:.3f} s = {nrec * rb / ctx.parse_seconds / 1e9:.2f} GB/s of text, {rows / ctx.parse_seconds / 1e6:.1f} M records/s; "
              f"with the Python generator's table copies {dt:.2f} s = {nrec * rb / dt / 1e9:.2f} GB/s", flush=True)
# ---- the headline text at its full size: 100 M records (31.6 GB) as consecutive windows of one stream, each generated on the device,
# copied to the pinned buffer and scanned; every row of every window is checked against arithmetic expectation
if os.environ.get("NT_PARSE_FULL", "1") != "0":
    total, slab = 100_000_000, nrec
    dslab = ctx.device_alloc(slab * rb)
    spent, rows_total = 0.0, 0
    col = np.arange(slab, dtype=np.uint64) * np.uint64(rb)
    for rec0 in range(0, total, slab):
        cnt = min(slab, total - rec0)
        ctx.synth_fastq_device(dslab, 0x5EED0002, rec0, cnt, L, 0)
        ctx.lib.ntg_memcpy_d2h(ctx.h, hp, dslab, cnt * rb)
        out = C.POINTER(nt._Records)(); consumed = C.c_uint64()
        t0 = time.perf_counter()
        ctx._ck(ctx.lib.ntg_parse_fastx_chunk(ctx.h, hp, cnt * rb, 2, int(rec0 + cnt == total), C.byref(out), C.byref(consumed)))
        spent += time.perf_counter() - t0
        rs = out.contents
        assert rs.error.kind == 0 and int(rs.n_records) == cnt and (rec0 + cnt == total or consumed.value == cnt * rb)
        t = np.ctypeslib.as_array(C.cast(rs.records, C.POINTER(C.c_uint64)), shape=(cnt, 10))
        s0 = col[:cnt]
        # '@r%09d' header (11 bytes), L bases, '+', L quality bytes, LF everywhere: start,id_b,id_e,seq_b,seq_e,qual_b,qual_e,all_e,num_bases,line
        exp = (s0, s0 + 1, s0 + 11, s0 + 12, s0 + 12 + L, s0 + 15 + L, s0 + 15 + 2 * L, s0 + 15 + 2 * L, None, None)
        for c, e in enumerate(exp):
            if e is not None:
                assert np.array_equal(t[:, c], e), (rec0, c)
        assert (t[:, 8] == L).all() and np.array_equal(t[:, 9], np.arange(cnt, dtype=np.uint64) * np.uint64(4) + np.uint64(1))
        rows_total += cnt
        ctx.lib.ntg_records_free(out)
    ctx.device_free(dslab)
    print(f"full headline text: {rows_total} rows over {total * rb / 1e9:.1f} GB in {-(-total // slab)} windows, all rows equal to arithmetic expectation; "
          f"inside the C ABI {spent:.2f} s = {total * rb / spent / 1e9:.2f} GB/s of text, {total / spent / 1e6:.1f} M records/s", flush=True)
del pinned
ctx.lib.ntg_free_pinned(hp)
ctx.close()
