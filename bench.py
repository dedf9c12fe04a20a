#!/usr/bin/env python3
"""bench.py — the headline metric of BASELINE.json on B200:

    Gbases/s, k=31 canonical k-mer + minimizer (m=21, w=11) over 100M x 150 bp synthetic FASTQ.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU; STRONG scaling: the 100M-read shape is split into N
   contiguous record shards, and the tallies are summed with one ncclAllReduce per step, enqueued on the compute
   stream right behind the kernel.)
  python bench.py --workload gz --reads 50000000 --read-len 250 --k 51 --m 0 [--gpus N]     # BASELINE config C5 shape

A "step" is one pass of the fused hot path over the whole 31.6 GB of FASTQ text (1/N of it per GPU) resident in HBM
(input >> L2, so no flush is needed between steps).  `e2e` is the same metric through the host-facing
C-ABI call (ntg_tally_fastx: pinned host bytes, H2D copies inside the timed region).  The CPU arm
(`--impl reference`, and `cpu_baseline` inside the default line) times the C++ oracle — a literal
restatement of the reference's Rust code, which cannot be built in this image (no rustc) — on the
box's host cores; it is a reported baseline, not the target.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Gbases/s k=31 canonical k-mer+minimizer over 100M x 150bp FASTQ"
SEED = 0x5EED0002


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic_ratio(kernel=None, fasta=False):
    """DRAM bytes per input byte of the named kernel from the committed ncu --set full captures (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        ks = t.get("kernels", {})
        for key in ((kernel + " (FASTA, C3)") if (kernel and fasta) else None, kernel):
            if key and key in ks:
                return float(ks[key]["dram_bytes_per_input_byte"])
        return float(t["dram_bytes_per_input_byte"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(L, k, m, budget_cpu_s=20.0, max_bytes=4 << 30):
    """Time the oracle (the reference's per-record loop, restated) on all host cores over a bounded sample."""
    import oracle_lib as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)      # the cores this process may use
    rb = 2 * L + 16
    probe = O.gen_fastq(SEED, 0, 20000, L, 0, nthreads=min(cores, 8))
    _, secs = O.bench_fastq(probe, rb, 1, k, m)
    rate1 = 20000 * L / max(secs, 1e-6)                      # bases/s on one thread
    nrec = int(min(budget_cpu_s * rate1 / L, max_bytes // rb))
    nrec = max(nrec - nrec % cores, cores * 1000)
    buf = O.gen_fastq(SEED, 0, nrec, L, 0, nthreads=cores)
    O.bench_fastq(buf, rb, cores, k, m)                      # warm-up (page faults, caches)
    best = None
    for _ in range(3):
        t, secs = O.bench_fastq(buf, rb, cores, k, m)
        best = secs if best is None else min(best, secs)
    return {"value": nrec * L / best / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "port",
            "sample": f"{nrec} records x {L} bp ({nrec * rb / 1e6:.0f} MB synthetic FASTQ in memory), best of 3, "
                      f"{cores} threads, C++ oracle (g++ -O3) of the reference loop; Rust reference not buildable here",
            "single_thread_value": rate1 / 1e9}, t


def cpu_baseline_fasta(L, k, m, seed, budget_s=15.0):
    """FASTA shapes: the oracle's whole-input loop (ntref_tally_fastx) on one shard per host thread (ctypes drops the GIL)."""
    import oracle_lib as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    t0 = time.perf_counter(); O.tally_fastx(O.gen_fasta(seed, 0, 20, L, 0).tobytes(), k=k, m=m); one = time.perf_counter() - t0
    nrec = max(4, int(budget_s / 3 / max(one / 20, 1e-9)))
    shards = [O.gen_fasta(seed, i * nrec, nrec, L, 0).tobytes() for i in range(cores)]
    best = None
    for _ in range(2):
        ths = [threading.Thread(target=O.tally_fastx, args=(sh,), kwargs=dict(k=k, m=m)) for sh in shards]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": cores * nrec * L / best / 1e9, "unit": "Gbases/s", "cores": cores, "kind": "port",
            "sample": f"{cores} x {nrec} records x {L} bp of synthetic FASTA, one shard per host thread, best of 2, C++ oracle of the reference loop"}, None


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import oracle_lib as O
    L, k, m = args.read_len, args.k, args.m
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rb = 2 * L + 16
    # size one step from a short all-threads probe so that the whole --steps/--warmup run takes about 90 s
    nprobe = cores * 4000
    probe = O.gen_fastq(SEED, 0, nprobe, L, 0, nthreads=cores)
    O.bench_fastq(probe, rb, cores, k, m)
    _, secs = O.bench_fastq(probe, rb, cores, k, m)
    rate_n = nprobe * L / max(secs, 1e-6)                    # bases/s on all host threads
    total_steps = args.steps + args.warmup
    nrec = int(min(90.0 / total_steps * rate_n / L, (2 << 30) // rb))
    nrec = max(nrec - nrec % cores, cores * 1000)
    buf = O.gen_fastq(SEED, 0, nrec, L, 0, nthreads=cores)
    for _ in range(args.warmup):
        O.bench_fastq(buf, rb, cores, k, m)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t, _ = O.bench_fastq(buf, rb, cores, k, m)
    dt = time.perf_counter() - t0
    value = args.steps * nrec * L / dt / 1e9
    sample = f"{nrec} records x {L} bp per step, {cores} host threads, C++ oracle port of the reference loop (Rust toolchain absent)"
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbases/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"synthetic {args.reads} x {L}bp FASTQ, k={k} canonical k-mers + m={m} minimizers (bounded sample per step)",
                   "k": k, "m": m, "w": k - m + 1, "read_len": L},
        "cpu_baseline": {"value": value, "unit": "Gbases/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def rec_bytes(L, fmt):
    return 2 * L + 16 if fmt == "fastq" else L + 12


def verify_full_scale(ctx, dbuf, nrec, rec0, L, k, m, full, n_thresh=0, fmt="fastq", seed=SEED):
    """Checks, outside the timed region, that the tallies of the whole resident shard are right — not only its counts:
    (1) the oracle (CPU) on sampled sub-shards of the very bytes resident in HBM, against the kernel on the same sub-shards;
    (2) the whole-shard tallies equal the sum over a partition into 7 parts (other tile alignments, other look-back chains).
    Checksums are wrapping u64 sums, so both properties are exact."""
    import numpy as np
    import oracle_lib as O
    rb = rec_bytes(L, fmt)
    gen = O.gen_fastq if fmt == "fastq" else O.gen_fasta
    keys = ("n_records", "n_bases", "n_kmers", "n_not_rc", "kmer_sum_lo", "n_minimizers", "minimizer_sum")
    sub = 20_000 if fmt == "fastq" else 400
    step4 = 4 if fmt == "fastq" else 16                          # record steps that keep sub-shards 16-byte aligned
    assert (rb * step4) % 16 == 0
    starts = sorted({(nrec - sub) * i // 4 // step4 * step4 for i in range(5)}) if nrec > sub else [0]
    for r0 in starts:
        n = min(sub, nrec - r0)
        got = ctx.tally_device(dbuf + r0 * rb, n * rb, k=k, m=m)
        exp = O.tally_fastx(gen(seed, rec0 + r0, n, L, n_thresh).tobytes(), k=k, m=m)
        for key in keys:
            assert got[key] == exp[key], ("oracle sample", r0, key, got[key], exp[key])
    parts, acc = 7, {key: 0 for key in keys}
    for i in range(parts):
        a, b = nrec * i // parts // step4 * step4, (nrec * (i + 1) // parts // step4 * step4 if i + 1 < parts else nrec)
        t = ctx.tally_device(dbuf + a * rb, (b - a) * rb, k=k, m=m)
        assert t["err_kind"] is None and t["fallback"] == 0
        for key in keys:
            acc[key] = (acc[key] + t[key]) & 0xFFFFFFFFFFFFFFFF
    for key in keys:
        assert acc[key] == full[key], ("partition sum", key, acc[key], full[key])
    return {"oracle_samples": len(starts), "records_per_sample": min(sub, nrec), "partition_parts": parts, "ok": True}


def run_gz_pipeline(args, ctx, dist, world, rank):
    """BASELINE config C5 shape: gzip-compressed FASTQ -> host inflate (BGZF members on `--gz-threads` workers, straight into
    pinned staging) -> H2D -> fused kernel, one stream session per rank, tallies reduced with NCCL at the end.
    The compressed input is a block of synthetic records compressed once and fed repeatedly until the rank's share is covered."""
    import numpy as np
    import oracle_lib as O
    from needletail_b200 import bgzf, shard
    L, k, m = args.read_len, args.k, args.m
    rb = 2 * L + 16
    first, nrec = shard.shard_records(args.reads, world, rank)
    block_rec = min(nrec, (32 << 20) // rb)
    text = O.gen_fastq(SEED + 3, first, block_rec, L, 0, nthreads=8).tobytes()
    comp = bgzf.compress(text, level=1, eof_marker=False)
    import ctypes as C
    hp = C.c_void_p()
    assert ctx.lib.ntg_alloc_pinned(len(comp), C.byref(hp)) == 0       # pinned: the device-inflate path copies straight from it
    blob = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(len(comp),))
    blob[:] = np.frombuffer(comp, dtype=np.uint8)
    exp = O.tally_fastx(text[: 2000 * rb], k=k, m=m)
    reps = max(1, nrec // block_rec)
    # --gz-threads: -1 = all host cores of this rank's share, 0 = inflate on the device (NTG_GZ_DEVICE), n = n host threads
    threads = args.gz_threads if args.gz_threads >= 0 else max(1, (os.cpu_count() or 8) // max(world, 1))

    def one_pass():
        s = ctx.stream(k=k, m=m)
        for _ in range(reps):
            s.feed_gz_ptr(blob.ctypes.data, blob.size, threads)
        return s.finish()

    s = ctx.stream(k=k, m=m); s.feed(text[: 2000 * rb]); t = s.finish()
    for key in ("n_records", "n_kmers", "kmer_sum_lo", "kmer_sum_hi", "n_not_rc"):
        assert t[key] == exp[key], (key, t[key], exp[key])
    one_pass()                                                  # warm-up
    ctx.sync()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        t = one_pass()
    ctx.sync()
    dt = (time.perf_counter() - t0) / steps
    assert t["err_kind"] is None and t["n_records"] == reps * block_rec, t
    if dist is not None:
        import torch
        tt = torch.tensor([dt], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt = float(tt.item())
        t = shard.allreduce_tallies({f: t[f] for f in nt_fields()}, ctx)
    total_rec = reps * block_rec * world if dist is not None else reps * block_rec
    if rank == 0:
        emit({"metric": f"Gbases/s k={k} canonical k-mers over gzip FASTQ, host inflate -> pinned -> H2D -> fused kernel", "value": total_rec * L / dt / 1e9,
              "unit": "Gbases/s", "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "u64", "data": "synthetic",
              "config": {"workload": f"BGZF-compressed synthetic {total_rec} x {L}bp FASTQ (a {block_rec}-record block fed {reps}x per rank), k={k}, m={m}",
                         "compressed_bytes_per_rank": int(blob.size) * reps, "text_bytes_per_rank": reps * block_rec * rb, "inflate": "device (gzdev::k_inflate)" if threads == 0 else f"host zlib, {threads} threads per rank",
                         "host_cores": os.cpu_count()},
              "e2e": {"value": total_rec * L / dt / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": (int(blob.size) * reps) if threads == 0 else reps * block_rec * rb, "d2h_bytes_per_step": 192 * reps,
                      "note": "device inflate: compressed bytes cross PCIe, text never exists on the host" if threads == 0 else "host-inflate bound: the kernel runs at ~1 TB/s of text"},
              "gpu_launches": None, "tallies": t})


def nt_fields():
    import needletail_b200 as nt
    return nt.TALLY_FIELDS


def run_ours(args):
    import numpy as np
    import needletail_b200 as nt
    from needletail_b200 import shard

    world, rank, local = env_int("WORLD_SIZE", 1), env_int("RANK", 0), env_int("LOCAL_RANK", 0)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = nt.Context(local)
    if world > 1:
        shard.init_nccl_comm(ctx)
    if args.workload == "gz":
        run_gz_pipeline(args, ctx, dist, world, rank)
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    L, k, m, fmt, seed = args.read_len, args.k, args.m, args.format, args.seed
    rb = rec_bytes(L, fmt)
    total_rec = args.reads                              # strong scaling: the named shape is split over the ranks
    rec0, nrec = shard.shard_records(total_rec, world, rank)
    nbytes = nrec * rb
    dbuf = ctx.device_alloc(nbytes)
    (ctx.synth_fastq_device if fmt == "fastq" else ctx.synth_fasta_device)(dbuf, seed, rec0, nrec, L, args.n_thresh)
    ctx.sync()

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def step():
        # one pass of the shard; with several ranks the tallies are summed by one ncclAllReduce enqueued on the compute
        # stream right behind the kernel (NTG_TALLY_ALLREDUCE): a single host wait per step
        ctx.tally_device_enqueue(dbuf, nbytes, k=k, m=m, allreduce=world > 1)
        t = ctx.tally_device_collect()
        kms = t.pop("fused_kernel_ms")
        err = t.pop("err_kind"); t.pop("err_line")
        assert err is None and t.pop("fallback") == 0 and not t.pop("not_reduced"), (err, "the fused single-pass kernel must produce the tallies")
        return t, kms

    for _ in range(args.warmup):
        tallies, _ = step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    ctx.event_record(0)
    kernel_ms = []
    for _ in range(args.steps):
        tallies, kms = step()
        kernel_ms.append(kms)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    if dist is not None:
        import torch
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    # size-independent checks on the full shape (clean synthetic reads: every window is a k-mer)
    assert tallies["n_records"] == total_rec, tallies
    assert tallies["n_bases"] == total_rec * L
    if args.n_thresh == 0:
        assert tallies["n_kmers"] == total_rec * (L - k + 1)
    assert tallies["n_minimizers"] == (tallies["n_kmers"] if m else 0)
    ms_per_step = ms / args.steps
    value = total_rec * L / (ms_per_step * 1e-3) / 1e9
    # ---- the checksums of the full shape, not only its counts (outside the timed region)
    verified = None
    if not args.no_verify:
        local_t = ctx.tally_device(dbuf, nbytes, k=k, m=m)
        verified = verify_full_scale(ctx, dbuf, nrec, rec0, L, k, m, local_t, args.n_thresh, fmt, seed)
        if world > 1:
            summed = shard.allreduce_tallies({f: local_t[f] for f in nt.TALLY_FIELDS}, ctx)     # host-staged reduce of the same shards
            for f in nt.TALLY_FIELDS:
                assert summed[f] == tallies[f], ("in-stream all-reduce vs host-staged all-reduce", f)

    # ---- roofline of the dominant kernel (k_fused): algorithmic bytes = the FASTQ text read once
    peak, peak_src = load_peak()
    kavg = sum(kernel_ms) / len(kernel_ms)
    achieved = nbytes / (kavg * 1e-3) / 1e9
    kname = "fqw::k_records" if tallies.get("fast_path") else "fused::k_fused"
    ratio = load_traffic_ratio(kname, fasta=(fmt == "fasta"))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (ratio * nbytes) if ratio else None,
                "kernel": "fqw::k_records (+ k_verify, fix-up launch)" if tallies.get("fast_path") else "fused::k_fused", "kernel_ms": kavg,
                "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src,
                "traffic_note": "DRAM bytes per launch = ncu dram read+write bytes per input byte (profiles/traffic.json) x algorithmic bytes",
                "note": ("single pass (DRAM traffic = 1.01 x algorithmic bytes) but integer-pipe bound, not HBM bound: ~41 SASS thread-instructions per base, ~25 of them on the 16-lane INT pipe, which is 84 % busy in the record-owned kernel (ncu profiles/r2h_*; 63 % in the tile kernel it replaces for short-read FASTQ); see DESIGN.md"
                         if tallies.get("fast_path") else
                         "single pass over the text (tile kernel: TMA-staged tiles, decoupled look-back) but integer-pipe bound, not HBM bound: the walker costs ~24 SASS instructions per base at k=21 m=11, ~31 at k=31 m=21; see DESIGN.md 3.1 / 6")}

    # ---- end to end through the host-facing C-ABI call: pinned host FASTQ -> H2D -> fused kernel -> tallies
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        gib = args.e2e_host_gib
        while True:                                              # pinned host buffer: halve until the box grants it
            host_rec = min(nrec, (gib << 30) // rb)
            hbytes = host_rec * rb
            hp = C.c_void_p()
            if ctx.lib.ntg_alloc_pinned(hbytes, C.byref(hp)) == 0:
                break
            assert gib > 1, "cannot pin even 1 GiB of host memory"
            gib //= 2
        ctx.lib.ntg_memcpy_d2h(ctx.h, hp, dbuf, hbytes)          # the host copy of the first host_rec records
        ctx.device_free(dbuf); dbuf = None                       # the call streams through its own three device segments
        calls = (nrec + host_rec - 1) // host_rec                # the host set is fed repeatedly until the shard is covered
        ctx.tally_ptr(hp.value, min(hbytes, 256 << 20) // rb * rb, k=k, m=m)      # warm-up (allocations)
        best = None
        for _ in range(2):
            barrier()
            t0 = time.perf_counter()
            done = 0
            for _ in range(calls):
                n_this = min(host_rec, nrec - done)
                t = ctx.tally_ptr(hp.value, n_this * rb, k=k, m=m)
                assert t["n_records"] == n_this and t["err_kind"] is None
                done += n_this
            ctx.sync()
            dt = time.perf_counter() - t0
            if dist is not None:
                import torch
                tt = torch.tensor([dt], device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); dt = float(tt.item())
            best = dt if best is None else min(best, dt)
        e2e = {"value": total_rec * L / best / 1e9, "unit": "Gbases/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 192 * ((nbytes >> 26) + 1),
               "seconds": best, "host_buffer_bytes": hbytes, "calls_per_step": calls, "repeats": 2,
               "note": "ntg_tally_fastx on pinned host text, streamed through three 64 MiB device segments (bounded device memory); PCIe H2D bound"}
        ctx.lib.ntg_free_pinned(hp)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = cpu_baseline(L, k, m) if fmt == "fastq" else cpu_baseline_fasta(L, k, m, seed)
    if dist is not None:
        dist.barrier()
    if rank == 0:
        emit({
            "metric": args.metric, "value": value, "unit": "Gbases/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: synthetic {total_rec} x {L}bp {fmt.upper()} ({total_rec * rb / 1e9:.1f} GB text) split over {world} GPU(s), resident in HBM, "
                                   f"k={k} canonical k-mers + m={m} minimizers, tallies" + (f", N bases at {args.n_thresh}/65536" if args.n_thresh else ""),
                       "k": k, "m": m, "w": k - m + 1, "read_len": L, "reads_per_gpu": nrec,
                       "l2": f"input per GPU ({nbytes / 1e9:.1f} GB) >> L2 (126 MB): no flush needed",
                       "parallelism": f"records sharded x{world}; one in-stream ncclAllReduce of the tallies per step" if world > 1 else "1 GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "tallies": tallies, "verified": verified,
        })
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print on fd 1 (e.g. NCCL's version banner) goes to stderr; the one JSON line is written
    to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--m", "--minimizer-len", dest="m", type=int, default=21)
    ap.add_argument("--e2e-host-gib", type=int, default=8)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--n-thresh", type=int, default=0, help="N bases: threshold / 65536 per base (655 = 1 %%, BASELINE config C4)")
    ap.add_argument("--workload", default="resident", choices=["resident", "gz"], help="gz: the compressed-input pipeline (BASELINE config C5 shape)")
    ap.add_argument("--gz-threads", type=int, default=-1, help="-1: host cores / ranks; 0: inflate on the device; n: n host threads")
    ap.add_argument("--config", default="C2", choices=["C2", "C3", "C4", "C5"],
                    help="BASELINE.json configs at their named sizes: C2 (default, the headline), C3 10M x 10kbp FASTA k=21 m=11, "
                         "C4 = C2 with 1 %% N, C5 = BGZF 50M x 250bp k=51 through the inflate pipeline")
    args = ap.parse_args()
    args.format, args.seed, args.metric = "fastq", SEED, METRIC
    if args.config == "C3":
        args.format, args.seed, args.reads, args.read_len, args.k, args.m = "fasta", 0x5EED0003, 10_000_000, 10_000, 21, 11
        args.metric = "Gbases/s k=21 bit k-mers + w=11 minimizers over 10M x 10kbp FASTA"
    elif args.config == "C4":
        args.n_thresh, args.seed = 655, 0x5EED0004
        args.metric = METRIC + " with 1% N bases"
    elif args.config == "C5":
        args.workload, args.reads, args.read_len, args.k, args.m = "gz", 50_000_000, 250, 51, 0
        if args.gz_threads < 0:
            args.gz_threads = 0                                  # inflate on the device
    if args.impl == "ours":
        args.warmup = max(args.warmup, 3)          # timing hygiene: at least three untimed passes
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
