// seqops.cuh — batch form of the reference's `Sequence` trait (materialising paths).
// One thread per input byte / per k-mer start; compaction through scan.cuh.  Inputs are a
// CSR batch (concatenated bytes + offsets).  Part of the unity build (ntgpu.cu).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace seqops {
constexpr int BLOCK = 256;
static inline unsigned grid_for(size_t n) { return (unsigned)((n + BLOCK - 1) / BLOCK); }

// largest s with offs[s] <= g  (offs has nseq+1 ascending entries, offs[nseq] > g)
__device__ __forceinline__ uint32_t seq_of(const uint32_t* __restrict__ offs, uint32_t nseq, uint32_t g) {
    uint32_t lo = 0, hi = nseq;   // invariant: offs[lo] <= g < offs[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(offs + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- normalize / strip_returns: flag + changed, then scatter ---------------------------------
// mode 0/1: sequence::normalize(iupac = mode)  (src/sequence.rs:19-62)
// mode 2  : Sequence::strip_returns            (src/sequence.rs:165-191)
// returns the output byte; *keep = false when the byte is deleted (mode 2 deletes only \r and \n: a NUL byte is kept)
__device__ __forceinline__ uint8_t xform(uint8_t b, int mode, bool* keep) {
    if (mode == 2) { *keep = !(b == '\r' || b == '\n'); return b; }
    const uint8_t o = __ldg(&c_norm[mode][b]);
    *keep = o != 0;
    return o;
}
__global__ void __launch_bounds__(BLOCK) k_xform_flags(const uint8_t* __restrict__ seqs, const uint32_t* __restrict__ offs,
                                                       uint32_t nseq, uint32_t total, int mode,
                                                       uint8_t* __restrict__ keep, uint8_t* __restrict__ changed) {
    uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g >= total) return;
    bool kp;
    uint8_t b = seqs[g], o = xform(b, mode, &kp);
    keep[g] = kp;
    if (!kp || o != b) changed[seq_of(offs, nseq, g)] = 1;   // benign race: everyone writes 1
}
__global__ void __launch_bounds__(BLOCK) k_xform_scatter(const uint8_t* __restrict__ seqs, uint32_t total, int mode,
                                                         const uint32_t* __restrict__ idx, uint8_t* __restrict__ out) {
    uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g >= total) return;
    bool kp;
    uint8_t o = xform(seqs[g], mode, &kp);
    if (kp) out[idx[g]] = o;
}
__global__ void __launch_bounds__(BLOCK) k_gather_offsets(const uint32_t* __restrict__ offs, uint32_t nseq,
                                                          const uint32_t* __restrict__ idx, uint64_t* __restrict__ out_offs) {
    uint32_t s = blockIdx.x * BLOCK + threadIdx.x;
    if (s <= nseq) out_offs[s] = idx[offs[s]];
}

// ---- reverse_complement / quality_mask --------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_revcomp(const uint8_t* __restrict__ seqs, const uint32_t* __restrict__ offs,
                                                   uint32_t nseq, uint32_t total, uint8_t* __restrict__ out) {
    uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g >= total) return;
    uint32_t s = seq_of(offs, nseq, g);
    uint32_t b = offs[s], e = offs[s + 1];
    out[b + (e - 1 - g)] = __ldg(&c_comp[seqs[g]]);      // complement(): src/sequence.rs:67-105
}
__global__ void __launch_bounds__(BLOCK) k_qmask(const uint8_t* __restrict__ seqs, const uint8_t* __restrict__ quals,
                                                 uint32_t total, uint8_t score, uint8_t* __restrict__ out) {
    uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g < total) out[g] = quals[g] < score ? (uint8_t)'N' : seqs[g];   // src/sequence.rs:280-297
}

// ---- k-mer start validity: window [g, g+k) inside its sequence and all good bases ------------
// (the set of windows both CanonicalKmers, src/kmer.rs:84-129, and BitNuclKmer,
//  src/bitkmer.rs:39-109, emit: every start whose k bytes are all in ACGTacgt)
__global__ void __launch_bounds__(BLOCK) k_kmer_valid(const uint8_t* __restrict__ seqs, const uint32_t* __restrict__ offs,
                                                      uint32_t nseq, uint32_t total, uint32_t k, int any_base, uint8_t* __restrict__ valid) {
    uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g >= total) return;
    uint32_t s = seq_of(offs, nseq, g);
    uint32_t e = offs[s + 1];
    uint8_t v = 0;
    if (g + k <= e) {
        v = 1;
        if (!any_base)
            for (uint32_t i = 0; i < k; i++)
                if (__ldg(&c_code[seqs[g + i]]) > 3) { v = 0; break; }
    }
    valid[g] = v;
}

// mode 0: canonical_kmers (byte compare, ties -> rc, was_rc = 1; val = 2-bit pack of the chosen slice, k <= 64)
// mode 1: bit_kmers(canonical = false)   mode 2: bit_kmers(canonical = true)   (k <= 32; ties -> original)
// mode 3: bit minimizers (m)             (k <= 32)
// mode 4: kmer::Kmers — every window, whatever its bytes (src/kmer.rs:13-41): positions only, the item IS the input slice
__global__ void __launch_bounds__(BLOCK) k_kmer_emit(const uint8_t* __restrict__ seqs, const uint8_t* __restrict__ rcs,
                                                     const uint32_t* __restrict__ offs, uint32_t nseq, uint32_t total,
                                                     uint32_t k, uint32_t m, int mode,
                                                     const uint8_t* __restrict__ valid, const uint32_t* __restrict__ idx,
                                                     uint32_t* __restrict__ pos, uint8_t* __restrict__ was_rc,
                                                     uint64_t* __restrict__ val_lo, uint64_t* __restrict__ val_hi) {
    uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g >= total || !valid[g]) return;
    uint32_t s = seq_of(offs, nseq, g);
    uint32_t b = offs[s], e = offs[s + 1];
    uint32_t p = g - b, len = e - b;
    uint32_t o = idx[g];
    pos[o] = p;
    if (mode == 4) return;
    if (mode == 0) {
        // result = buffer[pos..pos+k]; rc_result = rc_buffer[len-pos-k .. len-pos]   (src/kmer.rs:121-123)
        uint32_t rbase = b + (len - p - k);
        bool lt = false;     // result < rc_result ?
        for (uint32_t i = 0; i < k; i++) {
            uint8_t f = seqs[g + i];
            uint8_t r = rcs ? rcs[rbase + i] : __ldg(&c_comp[seqs[g + k - 1 - i]]);
            if (f != r) { lt = f < r; break; }
        }
        bool rc = !lt;       // ties => rc slice, was_rc = true   (src/kmer.rs:124-128)
        was_rc[o] = rc ? 1 : 0;
        uint64_t hi = 0, lo = 0;
        for (uint32_t i = 0; i < k; i++) {
            uint8_t c = rc ? (rcs ? rcs[rbase + i] : __ldg(&c_comp[seqs[g + k - 1 - i]])) : seqs[g + i];
            uint64_t code = __ldg(&c_code[c]) & 3;
            hi = (hi << 2) | (lo >> 62);
            lo = (lo << 2) | code;
        }
        val_lo[o] = lo;
        if (val_hi) val_hi[o] = hi;
    } else {
        uint64_t v = 0;
        for (uint32_t i = 0; i < k; i++) v = (v << 2) | (uint64_t)__ldg(&c_code[seqs[g + i]]);   // extend_kmer, bitkmer.rs:26-36
        if (mode == 1) { val_lo[o] = v; was_rc[o] = 0; }
        else if (mode == 2) {
            uint64_t r = bit_rc(v, k);                    // canonical(): bitkmer.rs:136-143
            bool rc = v > r;
            val_lo[o] = rc ? r : v; was_rc[o] = rc ? 1 : 0;
        } else {
            val_lo[o] = bit_minimizer_slow(v, k, m);      // bitkmer.rs:146-162
        }
    }
}

// ---- record writers: write_fasta / write_fastq (src/parser/record.rs:207-247) for a batch of kept records ------------------
// One warp per record: '>' / '@', id, line ending, raw_seq, line ending [, '+', line ending, qual (or 'I' x seq length when the
// record has none), line ending].  `src` holds the text the record table indexes; `dst_off[r]` is where record r's text starts.
struct WriteRow { uint64_t id_b, id_e, seq_b, seq_e, qual_b, qual_e, dst; uint32_t fastq, has_qual; };
__global__ void __launch_bounds__(BLOCK) k_write_records(const uint8_t* __restrict__ src, const WriteRow* __restrict__ rows, uint32_t n_rows,
                                                         uint32_t le_len, uint8_t* __restrict__ dst) {
    const uint32_t w = (blockIdx.x * BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const WriteRow r = rows[w];
    const uint64_t idl = r.id_e - r.id_b, sl = r.seq_e - r.seq_b, ql = r.fastq ? (r.has_qual ? r.qual_e - r.qual_b : sl) : 0;
    // segments of the output: [0] start byte, [1] id, [2] ending, [3] seq, [4] ending, (fastq) [5] '+', [6] ending, [7] qual, [8] ending
    const uint64_t e1 = 1 + idl, e2 = e1 + le_len, e3 = e2 + sl, e4 = e3 + le_len;
    const uint64_t e5 = e4 + (r.fastq ? 1 : 0), e6 = e5 + (r.fastq ? le_len : 0), e7 = e6 + ql, e8 = e7 + (r.fastq ? le_len : 0);
    auto ending = [&](uint64_t k) -> uint8_t { return (le_len == 2 && k == 0) ? (uint8_t)'\r' : (uint8_t)'\n'; };
    uint8_t* o = dst + r.dst;
    for (uint64_t p = lane; p < e8; p += 32) {
        uint8_t b;
        if (p == 0) b = r.fastq ? '@' : '>';
        else if (p < e1) b = src[r.id_b + (p - 1)];
        else if (p < e2) b = ending(p - e1);
        else if (p < e3) b = src[r.seq_b + (p - e2)];
        else if (p < e4) b = ending(p - e3);
        else if (p < e5) b = '+';
        else if (p < e6) b = ending(p - e5);
        else if (p < e7) b = r.has_qual ? src[r.qual_b + (p - e6)] : (uint8_t)'I';
        else b = ending(p - e7);
        o[p] = b;
    }
}

// ---- element-wise BitKmer helpers -------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_bitkmer_elem(const uint64_t* __restrict__ in, size_t n, uint32_t k, uint32_t m, int mode,
                                                        uint64_t* __restrict__ out, uint8_t* __restrict__ was_rc) {
    size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x;
    if (i >= n) return;
    uint64_t v = in[i];
    if (mode == 0) out[i] = bit_rc(v, k);
    else if (mode == 1) { uint64_t r = bit_rc(v, k); bool rc = v > r; out[i] = rc ? r : v; was_rc[i] = rc ? 1 : 0; }
    else out[i] = bit_minimizer_slow(v, k, m);
}
}  // namespace seqops

// ================================================================================== host side
struct BatchOnDevice {
    DevBuf<uint8_t> seqs;
    DevBuf<uint32_t> offs;
    uint32_t nseq = 0, total = 0;
};
static int upload_batch(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, BatchOnDevice& b) {
    if (!offs || (n && !seqs && offs[n] > 0)) return ntg_set_error(ctx, NTG_EINVAL, "null batch pointer");
    if (offs[0] != 0) return ntg_set_error(ctx, NTG_EINVAL, "offs[0] must be 0");
    for (size_t i = 0; i < n; i++)
        if (offs[i + 1] < offs[i]) return ntg_set_error(ctx, NTG_EINVAL, "offsets must be ascending");
    uint64_t total = offs[n];
    if (total >= 0xFFFFFFF0ull || n >= 0xFFFFFFF0ull)
        return ntg_set_error(ctx, NTG_EUNSUPPORTED, "a single sequence of %llu bytes: sequences are limited to 4 GiB", (unsigned long long)total);
    b.nseq = (uint32_t)n; b.total = (uint32_t)total;
    if (b.seqs.alloc_pooled(ctx->scratch, ScratchPool::SQ_SEQS, total) != cudaSuccess || b.offs.alloc_pooled(ctx->scratch, ScratchPool::SQ_OFFS, n + 1) != cudaSuccess)
        return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    std::vector<uint32_t> o32(n + 1);
    for (size_t i = 0; i <= n; i++) o32[i] = (uint32_t)offs[i];
    NTG_CUDA(ctx, cudaMemcpyAsync(b.seqs.p, seqs, total, cudaMemcpyHostToDevice, ctx->stream));
    NTG_CUDA(ctx, cudaMemcpyAsync(b.offs.p, o32.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // o32 is a stack-scoped staging vector
    return NTG_OK;
}

// Batches beyond the 32-bit indices of one device pass are cut into runs of whole sequences (at most SUB_BATCH bytes each).
constexpr uint64_t SUB_BATCH = uint64_t(1) << 31;
static int check_batch(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n) {
    if (!offs || (n && !seqs && offs[n] > 0)) return ntg_set_error(ctx, NTG_EINVAL, "null batch pointer");
    if (offs[0] != 0) return ntg_set_error(ctx, NTG_EINVAL, "offs[0] must be 0");
    for (size_t i = 0; i < n; i++)
        if (offs[i + 1] < offs[i]) return ntg_set_error(ctx, NTG_EINVAL, "offsets must be ascending");
    return NTG_OK;
}
// next run [a, b) of sequences starting at a: as many whole sequences as fit (at least one)
static size_t next_run(const uint64_t* offs, size_t n, size_t a) {
    size_t b = a + 1;
    while (b < n && offs[b + 1] - offs[a] <= SUB_BATCH) b++;
    return b;
}

static int run_xform_one(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, int mode,
                         uint8_t* out, uint64_t* out_offs, uint8_t* changed);
static int run_xform(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, int mode,
                     uint8_t* out, uint64_t* out_offs, uint8_t* changed) {
    if (!out || !out_offs || !changed) return ntg_set_error(ctx, NTG_EINVAL, "null output pointer");
    NTG_TRY(check_batch(ctx, seqs, offs, n));
    if (offs[n] <= SUB_BATCH) return run_xform_one(ctx, seqs, offs, n, mode, out, out_offs, changed);
    uint64_t out_total = 0;
    std::vector<uint64_t> sub, sub_out;
    for (size_t a = 0; a < n;) {
        const size_t b = next_run(offs, n, a);
        sub.resize(b - a + 1); sub_out.resize(b - a + 1);
        for (size_t i = a; i <= b; i++) sub[i - a] = offs[i] - offs[a];
        NTG_TRY(run_xform_one(ctx, seqs + offs[a], sub.data(), b - a, mode, out + out_total, sub_out.data(), changed + a));
        for (size_t i = a; i <= b; i++) out_offs[i] = out_total + sub_out[i - a];
        out_total += sub_out[b - a];
        a = b;
    }
    return NTG_OK;
}
static int run_xform_one(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, int mode,
                         uint8_t* out, uint64_t* out_offs, uint8_t* changed) {
    using namespace seqops;
    BatchOnDevice b;
    NTG_TRY(upload_batch(ctx, seqs, offs, n, b));
    DevBuf<uint8_t> keep, dout, dchg; DevBuf<uint32_t> idx, tmp; DevBuf<uint64_t> dooffs;
    ScratchPool& sp = ctx->scratch;
    if (keep.alloc_pooled(sp, ScratchPool::SQ_FLAGS, b.total) || dout.alloc_pooled(sp, ScratchPool::SQ_OUT, b.total) || dchg.alloc_pooled(sp, ScratchPool::SQ_CHG, n) ||
        idx.alloc_pooled(sp, ScratchPool::SQ_IDX, (size_t)b.total + 1) || tmp.alloc_pooled(sp, ScratchPool::SQ_TMP, scan_tmp_count(b.total)) ||
        dooffs.alloc_pooled(sp, ScratchPool::SQ_OOFFS, n + 1))
        return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    NTG_CUDA(ctx, cudaMemsetAsync(dchg.p, 0, n ? n : 1, ctx->stream));
    if (b.total) {
        k_xform_flags<<<grid_for(b.total), BLOCK, 0, ctx->stream>>>(b.seqs.p, b.offs.p, b.nseq, b.total, mode, keep.p, dchg.p);
        ctx->launches++;
    }
    NTG_TRY(exclusive_scan_u8(ctx, keep.p, idx.p, b.total, tmp.p));
    if (b.total) {
        k_xform_scatter<<<grid_for(b.total), BLOCK, 0, ctx->stream>>>(b.seqs.p, b.total, mode, idx.p, dout.p);
        ctx->launches++;
    }
    k_gather_offsets<<<grid_for(n + 1), BLOCK, 0, ctx->stream>>>(b.offs.p, b.nseq, idx.p, dooffs.p);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    NTG_CUDA(ctx, cudaMemcpyAsync(out_offs, dooffs.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (n) NTG_CUDA(ctx, cudaMemcpyAsync(changed, dchg.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t out_total = out_offs[n];
    if (out_total) NTG_CUDA(ctx, cudaMemcpy(out, dout.p, out_total, cudaMemcpyDeviceToHost));
    return NTG_OK;
}

struct ItemsPriv {
    PinBuf<uint64_t> item_offs, val_lo, val_hi;
    PinBuf<uint32_t> pos;
    PinBuf<uint8_t> was_rc;
};

// mode: 0 canonical_kmers, 1 bit_kmers, 2 bit_kmers canonical, 3 minimizers, 4 plain windows (Kmers)
static int run_kmers_one(ntg_ctx* ctx, const uint8_t* seqs, const uint8_t* rc, const uint64_t* offs, size_t n,
                     uint32_t k, uint32_t m, int mode, ntg_items** out) {
    using namespace seqops;
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output pointer");
    *out = nullptr;
    if (k == 0) return ntg_set_error(ctx, NTG_EINVAL, "k must be >= 1 (k = 0 panics in the reference, src/kmer.rs:91)");
    if (mode == 0 && k > 64) return ntg_set_error(ctx, NTG_EINVAL, "canonical_kmers: k <= 64 supported");
    if (mode == 4 && k > 255) return ntg_set_error(ctx, NTG_EINVAL, "k is a u8 in the reference (src/kmer.rs:14)");
    if (mode != 0 && mode != 4 && k > 32) return ntg_set_error(ctx, NTG_EINVAL, "bit k-mers hold at most 32 bases (u64, src/bitkmer.rs:2-3)");
    if (mode == 3 && (m == 0 || m > k)) return ntg_set_error(ctx, NTG_EINVAL, "minimizer needs 1 <= m <= k");
    BatchOnDevice b;
    NTG_TRY(upload_batch(ctx, seqs, offs, n, b));
    DevBuf<uint8_t> drc, valid, dwas; DevBuf<uint32_t> idx, tmp, dpos; DevBuf<uint64_t> dlo, dhi, dioffs;
    if (rc) {
        if (drc.alloc_pooled(ctx->scratch, ScratchPool::SQ_RC, b.total)) return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
        NTG_CUDA(ctx, cudaMemcpyAsync(drc.p, rc, b.total, cudaMemcpyHostToDevice, ctx->stream));
    }
    ScratchPool& sp = ctx->scratch;
    if (valid.alloc_pooled(sp, ScratchPool::SQ_FLAGS, b.total) || idx.alloc_pooled(sp, ScratchPool::SQ_IDX, (size_t)b.total + 1) ||
        tmp.alloc_pooled(sp, ScratchPool::SQ_TMP, scan_tmp_count(b.total)) || dioffs.alloc_pooled(sp, ScratchPool::SQ_OOFFS, n + 1))
        return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    if (b.total) {
        k_kmer_valid<<<grid_for(b.total), BLOCK, 0, ctx->stream>>>(b.seqs.p, b.offs.p, b.nseq, b.total, k, mode == 4 ? 1 : 0, valid.p);
        ctx->launches++;
    }
    NTG_TRY(exclusive_scan_u8(ctx, valid.p, idx.p, b.total, tmp.p));
    k_gather_offsets<<<grid_for(n + 1), BLOCK, 0, ctx->stream>>>(b.offs.p, b.nseq, idx.p, dioffs.p);
    ctx->launches++;
    uint32_t n_items = 0;
    NTG_CUDA(ctx, cudaMemcpyAsync(&n_items, idx.p + b.total, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const bool want_val = mode != 4;
    bool want_hi = (mode == 0 && k > 32), want_rc = (mode != 3 && mode != 4);
    if (dpos.alloc_pooled(sp, ScratchPool::SQ_OUT, n_items) || (want_val && dlo.alloc_pooled(sp, ScratchPool::SQ_LO, n_items)) ||
        (want_hi && dhi.alloc_pooled(sp, ScratchPool::SQ_HI, n_items)) || dwas.alloc_pooled(sp, ScratchPool::SQ_CHG, n_items))
        return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    if (b.total && n_items) {
        k_kmer_emit<<<grid_for(b.total), BLOCK, 0, ctx->stream>>>(b.seqs.p, rc ? drc.p : nullptr, b.offs.p, b.nseq, b.total, k, m,
                                                                 mode, valid.p, idx.p, dpos.p, dwas.p, dlo.p, want_hi ? dhi.p : nullptr);
        ctx->launches++;
    }
    NTG_CUDA(ctx, cudaGetLastError());
    auto* priv = new ItemsPriv();
    auto* it = new ntg_items();
    auto fail = [&](int st, const char* msg) { delete priv; delete it; return ntg_set_error(ctx, st, "%s", msg); };
    if (priv->item_offs.alloc_pooled(ctx->pinpool, n + 1) || priv->pos.alloc_pooled(ctx->pinpool, n_items) || (want_val && priv->val_lo.alloc_pooled(ctx->pinpool, n_items)) ||
        (want_hi && priv->val_hi.alloc_pooled(ctx->pinpool, n_items)) || (want_rc && priv->was_rc.alloc_pooled(ctx->pinpool, n_items)))
        return fail(NTG_ENOMEM, "pinned allocation failed");
    cudaError_t e = cudaMemcpyAsync(priv->item_offs.p, dioffs.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (!e && n_items) {
        e = cudaMemcpyAsync(priv->pos.p, dpos.p, (size_t)n_items * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (!e && want_val) e = cudaMemcpyAsync(priv->val_lo.p, dlo.p, (size_t)n_items * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (!e && want_hi) e = cudaMemcpyAsync(priv->val_hi.p, dhi.p, (size_t)n_items * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (!e && want_rc) e = cudaMemcpyAsync(priv->was_rc.p, dwas.p, (size_t)n_items, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (!e) e = cudaStreamSynchronize(ctx->stream);
    if (e) return fail(NTG_ECUDA, cudaGetErrorString(e));
    it->n_seqs = n; it->n_items = n_items;
    it->item_offs = priv->item_offs.p; it->pos = priv->pos.p;
    it->was_rc = want_rc ? priv->was_rc.p : nullptr;
    it->val_lo = want_val ? priv->val_lo.p : nullptr; it->val_hi = want_hi ? priv->val_hi.p : nullptr;
    it->_priv = priv;
    *out = it;
    return NTG_OK;
}

static int run_kmers(ntg_ctx* ctx, const uint8_t* seqs, const uint8_t* rc, const uint64_t* offs, size_t n,
                     uint32_t k, uint32_t m, int mode, ntg_items** out) {
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output pointer");
    *out = nullptr;
    NTG_TRY(check_batch(ctx, seqs, offs, n));
    if (offs[n] <= SUB_BATCH) return run_kmers_one(ctx, seqs, rc, offs, n, k, m, mode, out);
    // runs of whole sequences, merged into one result (item offsets shifted by the items of the runs before)
    std::vector<ntg_items*> parts;
    auto drop = [&]() { for (auto* p : parts) ntg_items_free(p); };
    std::vector<uint64_t> sub;
    uint64_t n_items = 0;
    for (size_t a = 0; a < n;) {
        const size_t b = next_run(offs, n, a);
        sub.resize(b - a + 1);
        for (size_t i = a; i <= b; i++) sub[i - a] = offs[i] - offs[a];
        ntg_items* part = nullptr;
        const int st = run_kmers_one(ctx, seqs + offs[a], rc ? rc + offs[a] : nullptr, sub.data(), b - a, k, m, mode, &part);
        if (st != NTG_OK) { drop(); return st; }
        parts.push_back(part); n_items += part->n_items;
        a = b;
    }
    auto* priv = new ItemsPriv();
    auto* it = new ntg_items();
    const bool want_val = parts[0]->val_lo != nullptr, want_hi = parts[0]->val_hi != nullptr, want_rc = parts[0]->was_rc != nullptr;
    if (priv->item_offs.alloc_pooled(ctx->pinpool, n + 1) || priv->pos.alloc_pooled(ctx->pinpool, n_items) || (want_val && priv->val_lo.alloc_pooled(ctx->pinpool, n_items)) ||
        (want_hi && priv->val_hi.alloc_pooled(ctx->pinpool, n_items)) || (want_rc && priv->was_rc.alloc_pooled(ctx->pinpool, n_items))) {
        delete priv; delete it; drop();
        return ntg_set_error(ctx, NTG_ENOMEM, "pinned allocation failed");
    }
    uint64_t io = 0; size_t so = 0;
    for (auto* p : parts) {
        for (uint64_t i = 0; i <= p->n_seqs; i++) priv->item_offs.p[so + i] = io + p->item_offs[i];
        std::memcpy(priv->pos.p + io, p->pos, p->n_items * sizeof(uint32_t));
        if (want_val) std::memcpy(priv->val_lo.p + io, p->val_lo, p->n_items * sizeof(uint64_t));
        if (want_hi) std::memcpy(priv->val_hi.p + io, p->val_hi, p->n_items * sizeof(uint64_t));
        if (want_rc) std::memcpy(priv->was_rc.p + io, p->was_rc, p->n_items);
        io += p->n_items; so += p->n_seqs;
    }
    drop();
    it->n_seqs = n; it->n_items = n_items;
    it->item_offs = priv->item_offs.p; it->pos = priv->pos.p;
    it->was_rc = want_rc ? priv->was_rc.p : nullptr;
    it->val_lo = want_val ? priv->val_lo.p : nullptr; it->val_hi = want_hi ? priv->val_hi.p : nullptr;
    it->_priv = priv;
    *out = it;
    return NTG_OK;
}

// Text of the kept records of a record table, in table order.  keep == nullptr: all.  line_ending: NTG_LE_UNIX / NTG_LE_WINDOWS.
static int run_write_records(ntg_ctx* ctx, const uint8_t* bytes, size_t n, int format, const ntg_record* recs, size_t n_recs, const uint8_t* keep,
                             int line_ending, uint8_t* out, size_t out_cap, size_t* out_len) {
    using namespace seqops;
    if (!out_len || (n_recs && !recs) || (n && !bytes)) return ntg_set_error(ctx, NTG_EINVAL, "null pointer");
    if (format != NTG_FMT_FASTA && format != NTG_FMT_FASTQ) return ntg_set_error(ctx, NTG_EINVAL, "format must be FASTA or FASTQ");
    if (line_ending != NTG_LE_UNIX && line_ending != NTG_LE_WINDOWS) return ntg_set_error(ctx, NTG_EINVAL, "line ending must be unix or windows");
    const uint32_t le = line_ending == NTG_LE_WINDOWS ? 2 : 1;
    const bool fq = format == NTG_FMT_FASTQ;
    std::vector<WriteRow> rows;
    uint64_t total = 0;
    for (size_t i = 0; i < n_recs; i++) {
        if (keep && !keep[i]) continue;
        const ntg_record& r = recs[i];
        if (r.id_e > n || r.seq_e > n || (fq && r.qual_e > n) || r.id_b > r.id_e || r.seq_b > r.seq_e) return ntg_set_error(ctx, NTG_EINVAL, "record %zu points outside the buffer", i);
        const uint64_t sl = r.seq_e - r.seq_b;
        rows.push_back(WriteRow{r.id_b, r.id_e, r.seq_b, r.seq_e, r.qual_b, r.qual_e, total, fq ? 1u : 0u, 1u});
        total += 1 + (r.id_e - r.id_b) + le + sl + le + (fq ? 1 + le + (r.qual_e - r.qual_b) + le : 0);
    }
    *out_len = (size_t)total;
    if (total > out_cap) return ntg_set_error(ctx, NTG_EINVAL, "output buffer too small: %llu bytes needed", (unsigned long long)total);
    if (rows.empty()) return NTG_OK;
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output");
    if (rows.size() >= 0x07FFFFFFull) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "write at most 2^27 records per call");
    DevBuf<uint8_t> dsrc, ddst; DevBuf<WriteRow> drows;
    if (dsrc.alloc(n) || ddst.alloc(total) || drows.alloc(rows.size())) return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    NTG_CUDA(ctx, cudaMemcpyAsync(dsrc.p, bytes, n, cudaMemcpyHostToDevice, ctx->stream));
    NTG_CUDA(ctx, cudaMemcpyAsync(drows.p, rows.data(), rows.size() * sizeof(WriteRow), cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t threads = (uint64_t)rows.size() * 32;
    k_write_records<<<(unsigned)((threads + BLOCK - 1) / BLOCK), BLOCK, 0, ctx->stream>>>(dsrc.p, drows.p, (uint32_t)rows.size(), le, ddst.p);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    NTG_CUDA(ctx, cudaMemcpyAsync(out, ddst.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NTG_OK;
}

static int run_bitkmer_elem(ntg_ctx* ctx, const uint64_t* in, size_t n, uint32_t k, uint32_t m, int mode, uint64_t* out, uint8_t* was_rc) {
    using namespace seqops;
    if (!in || !out || (mode == 1 && !was_rc)) return ntg_set_error(ctx, NTG_EINVAL, "null pointer");
    if (k == 0 || k > 32) return ntg_set_error(ctx, NTG_EINVAL, "BitKmer needs 1 <= k <= 32 (src/bitkmer.rs:130)");
    if (mode == 2 && (m == 0 || m > k)) return ntg_set_error(ctx, NTG_EINVAL, "minimizer needs 1 <= m <= k");
    if (n == 0) return NTG_OK;
    DevBuf<uint64_t> din, dout; DevBuf<uint8_t> dw;
    if (din.alloc(n) || dout.alloc(n) || dw.alloc(n)) return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    NTG_CUDA(ctx, cudaMemcpyAsync(din.p, in, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_bitkmer_elem<<<grid_for(n), BLOCK, 0, ctx->stream>>>(din.p, n, k, m, mode, dout.p, dw.p);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    NTG_CUDA(ctx, cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (mode == 1) NTG_CUDA(ctx, cudaMemcpyAsync(was_rc, dw.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NTG_OK;
}
