#!/usr/bin/env python3
"""Kernel time of the fused path on the C2 shape, results ignored (for timing-only experiment builds whose tallies are wrong)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import needletail_b200 as nt
ctx = nt.Context(0)
reads, L = 100_000_000, 150
nbytes = reads * 316
d = ctx.device_alloc(nbytes)
ctx.synth_fastq_device(d, 0x5EED0002, 0, reads, L, 0)
ctx.sync()
ms = []
for i in range(6):
    ctx.event_record(0)
    ctx.tally_device_enqueue(d, nbytes, k=31, m=21)
    ctx.event_record(1)
    ms.append(ctx.event_elapsed_ms(0, 1))
    try:
        ctx.tally_device_collect()
    except Exception as e:
        pass
print("ms per pass", [round(x, 2) for x in ms], "Gbases/s", reads * L / (sum(ms[2:]) / 4) / 1e6)
