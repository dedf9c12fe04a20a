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
    ("C3 10M x 10kbp FASTA k=21 m=11 (the named size: 100 GB of text resident in HBM)", "fasta", 10_000_000, 10_000, 21, 11, 0, 0x5EED0003),
    # not a BASELINE config: the same FASTA shape wrapped at 70 columns (what genome FASTA files look like); generated on the host
    ("C3w 120k x 10kbp FASTA wrapped at 70 columns, k=21 m=11 (host-generated, 1.2 GB)", "fasta_wrapped", 120_000, 10_000, 21, 11, 0, 0x5EED0013),
]


def wrapped_fasta(reads, L, width, seed):
    """reads records '>r%08d\\n' + L random ACGT bases wrapped at `width` columns, as one uint8 array (numpy, host)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    full, rest = divmod(L, width)
    body = full * (width + 1) + (rest + 1 if rest else 0)
    ids = np.char.add(">r", np.char.zfill(np.arange(reads).astype(str), 8))
    hdr = np.frombuffer("".join(i + "\n" for i in ids).encode(), dtype=np.uint8).reshape(reads, 11)
    out = np.empty((reads, 11 + body), dtype=np.uint8)
    out[:, :11] = hdr
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(reads, L), dtype=np.uint8)]
    if full:
        blk = out[:, 11:11 + full * (width + 1)].reshape(reads, full, width + 1)
        blk[:, :, :width] = bases[:, :full * width].reshape(reads, full, width)
        blk[:, :, width] = 10
    if rest:
        out[:, 11 + full * (width + 1):11 + full * (width + 1) + rest] = bases[:, full * width:]
        out[:, -1] = 10
    return out.reshape(-1)


def main():
    ctx = nt.Context(0)
    ctx.tally_flags = int(os.environ.get("NT_MC_FLAGS", "0"))     # 4 = NTG_TALLY_NO_FASTPATH: the tile kernel on short-read FASTQ too
    out = []
    only = os.environ.get("NT_MC_ONLY")
    configs = [CONFIGS[int(i)] for i in only.split(",")] if only else CONFIGS
    for name, kind, reads, L, k, m, nth, seed in configs:
        if kind == "fasta_wrapped":
            host = wrapped_fasta(reads, L, 70, seed)
            nbytes = int(host.size)
            d = ctx.device_alloc(nbytes)
            ctx.h2d(d, host)
            del host
        else:
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
        if any(t["stats"].values()):              # NTG_STATS build: per-CTA cycle accounting (see fused.cuh)
            out[-1]["stats"] = dict(t["stats"], n_query_slot=t["n_query"])
        print(json.dumps(out[-1]), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
