#!/usr/bin/env python3
"""Kernel-resident throughput of the fused path on the BASELINE.json config shapes (one GPU).
Not the bench line: a side table for DESIGN.md.  Sizes are reduced where noted so the run stays short.

    python tools/measure_configs.py > gpurun_out/configs.json
    NT_MC_ONLY=0,1,3 NTGPU_SO=needletail_b200/libntgpu_base.so python tools/measure_configs.py     # A/B of an experiment build
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import needletail_b200 as nt

CONFIGS = [
    # name, kind, reads, L, k, m, n_thresh, seed
    ("C2 100M x 150bp FASTQ k=31 m=21", "fastq", 100_000_000, 150, 31, 21, 0, 0x5EED0002),
    ("C2' same, canonical k-mers only (m=0)", "fastq", 100_000_000, 150, 31, 0, 0, 0x5EED0002),
    ("C4 100M x 150bp FASTQ 1% N k=31 m=21 (one GPU's shard)", "fastq", 100_000_000, 150, 31, 21, 655, 0x5EED0004),
    ("C3 shape, 4M x 10kbp FASTA k=21 m=11 (40 of the 100 Gbases)", "fasta", 4_000_000, 10_000, 21, 11, 0, 0x5EED0003),
    ("C5 shape, 50M x 250bp FASTQ k=51 (text resident; gzip inflate is host work)", "fastq", 50_000_000, 250, 51, 0, 0, 0x5EED0005),
]


def main():
    ctx = nt.Context(0)
    out = []
    only = os.environ.get("NT_MC_ONLY")
    configs = [CONFIGS[int(i)] for i in only.split(",")] if only else CONFIGS
    for name, kind, reads, L, k, m, nth, seed in configs:
        rb = 2 * L + 16 if kind == "fastq" else L + 12
        nbytes = reads * rb
        d = ctx.device_alloc(nbytes)
        (ctx.synth_fastq_device if kind == "fastq" else ctx.synth_fasta_device)(d, seed, 0, reads, L, nth)
        ctx.sync()
        ms = []
        t = None
        for i in range(3 + 5):
            ctx.tally_device_enqueue(d, nbytes, k=k, m=m)
            t = ctx.tally_device_collect()
            if i >= 3:
                ms.append(t["fused_kernel_ms"])
        ctx.device_free(d)
        avg = sum(ms) / len(ms)
        assert t["err_kind"] is None and t["fallback"] == 0 and t["n_records"] == reads and t["n_bases"] == reads * L
        out.append({"config": name, "bytes": nbytes, "kernel_ms": avg, "gbases_per_s": reads * L / avg / 1e6,
                    "gb_per_s": nbytes / avg / 1e6, "n_kmers": t["n_kmers"], "n_not_rc": t["n_not_rc"],
                    "kmer_sum_lo": t["kmer_sum_lo"], "kmer_sum_hi": t["kmer_sum_hi"], "minimizer_sum": t["minimizer_sum"]})
        if any(t["ws_cycles"].values()):          # NTG_STATS build: per-CTA cycle accounting (see fused.cuh)
            c = t["ws_cycles"]
            out[-1]["stats"] = {"cta_cycles_sum": c["claim"], "lookback_cycles": c["scan"], "lookbacks": c["lookback_retry"],
                                "barrier_wait_cycles_t0": c["walker_wait"], "walk_cycles_t0": c["walker_work"]}
        print(json.dumps(out[-1]), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
