// parse.cuh — FASTX record scanner producing the record table (materialising path).
//
// Whole-buffer formulation of the reference's readers in two passes over the window:
//   k_index_count : per 8 KiB tile, the number of newlines (FASTA also: record starts — '>' at a line start — and '\r')
//   k_index_scan  : exclusive scan of the tile counts (one CTA)
//   k_index_emit  : the ordered lists — position of every newline; FASTA: position, newline ordinal and '\r' ordinal of every
//                   record start (what num_bases needs: fasta.rs:102-107 without a per-byte index)
// then one thread per record fills the ntg_record row and validates it.
//   FASTQ: record r owns newlines 4r..4r+3          (src/parser/fastq.rs:155-187)
//   FASTA: a record starts at byte 0 and after every "\n>"   (src/parser/fasta.rs:220-243)
// Round 1 wrote a flag byte and a 32-bit ordinal per input byte (12 bytes of traffic and 9 bytes of scratch per byte, allocated
// per call); this form reads the window twice and writes 4 bytes per newline, from buffers the context keeps (ScratchPool).
// Only the O(1) end-of-stream rules (fastq.rs:337-356, fasta.rs:200-216) run on the host, on the
// few newline positions the device hands back.  Part of the unity build (ntgpu.cu).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace parse {
constexpr int BLOCK = 256;
static inline unsigned grid_for(size_t n) { return (unsigned)((n + BLOCK - 1) / BLOCK); }

constexpr int IDX_PER = 32;                       // bytes per thread: one bit mask per delimiter
constexpr int IDX_TILE = BLOCK * IDX_PER;         // 8 KiB per CTA
struct Masks { uint32_t nl, cr, gt; };           // bit i = byte i of the thread's span

// 4-bit mask of the bytes of w equal to the byte replicated in pat (exact: no false positives from borrows)
__device__ __forceinline__ uint32_t eq4(uint32_t w, uint32_t pat) {
    const uint32_t x = w ^ pat;
    const uint32_t t = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);      // 0x80 in the bytes that are zero
    return (((t >> 7) * 0x00204081u) >> 21) & 0xFu;
}
__device__ __forceinline__ Masks span_masks(const uint8_t* __restrict__ bytes, uint32_t g0, uint32_t n, bool fasta, bool aligned) {
    Masks m{0, 0, 0};
    if (g0 >= n) return m;
    uint32_t w[8];
    if (aligned && g0 + IDX_PER <= n) {
        const uint4 a = *reinterpret_cast<const uint4*>(bytes + g0), b = *reinterpret_cast<const uint4*>(bytes + g0 + 16);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) { const uint32_t g = g0 + 4 * q + j; if (g < n) v |= (uint32_t)bytes[g] << (8 * j); }
            w[q] = v;                                 // (bytes behind the window read as 0: no delimiter)
        }
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        m.nl |= eq4(w[q], 0x0A0A0A0Au) << (4 * q);
        if (fasta) { m.cr |= eq4(w[q], 0x0D0D0D0Du) << (4 * q); m.gt |= eq4(w[q], 0x3E3E3E3Eu) << (4 * q); }
    }
    return m;
}
// FASTA record starts of the span: '>' at byte 0 of the window or right behind a newline (fasta.rs:220-243)
__device__ __forceinline__ uint32_t start_mask(const uint8_t* __restrict__ bytes, uint32_t g0, uint32_t n, const Masks& m) {
    if (g0 >= n || !m.gt) return 0;
    const uint32_t prev_nl = (g0 == 0 || bytes[g0 - 1] == '\n') ? 1u : 0u;
    return m.gt & ((m.nl << 1) | prev_nl);
}
// per tile: counts of newlines / record starts / '\r'
__global__ void __launch_bounds__(BLOCK) k_index_count(const uint8_t* __restrict__ bytes, uint32_t n, int fasta, int aligned,
                                                       uint32_t* __restrict__ cnt_nl, uint32_t* __restrict__ cnt_st, uint32_t* __restrict__ cnt_cr) {
    __shared__ uint32_t red[BLOCK / 32][3];
    const uint32_t g0 = blockIdx.x * IDX_TILE + threadIdx.x * IDX_PER;
    const Masks m = span_masks(bytes, g0, n, fasta != 0, aligned != 0);
    uint32_t c0 = __popc(m.nl), c1 = fasta ? __popc(start_mask(bytes, g0, n, m)) : 0u, c2 = __popc(m.cr);
    c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = c0; red[threadIdx.x >> 5][1] = c1; red[threadIdx.x >> 5][2] = c2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t s = 0;
        for (int wi = 0; wi < BLOCK / 32; wi++) s += red[wi][threadIdx.x];
        (threadIdx.x == 0 ? cnt_nl : threadIdx.x == 1 ? cnt_st : cnt_cr)[blockIdx.x] = s;
    }
}
// in-place exclusive scan of the three count arrays (n_tiles entries each), totals at [n_tiles]; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_index_scan(uint32_t* __restrict__ a0, uint32_t* __restrict__ a1, uint32_t* __restrict__ a2, uint32_t n_tiles, int n_arrays) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    for (int q = 0; q < n_arrays; q++) {
        uint32_t* a = q == 0 ? a0 : q == 1 ? a1 : a2;
        if (threadIdx.x == 0) carry_s = 0;
        __syncthreads();
        for (uint32_t base = 0; base < n_tiles; base += 1024) {
            const uint32_t i = base + threadIdx.x;
            const uint32_t v = i < n_tiles ? a[i] : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= d) incl += t; }
            if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
            __syncthreads();
            if (threadIdx.x < 32) {
                uint32_t ws = wsum[threadIdx.x];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, ws, d); if (threadIdx.x >= d) ws += t; }
                wsum[threadIdx.x] = ws;                   // inclusive warp totals
            }
            __syncthreads();
            const uint32_t carry = carry_s;
            const uint32_t before = carry + ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0u) + incl - v;
            if (i < n_tiles) a[i] = before;
            __syncthreads();
            if (threadIdx.x == 1023) carry_s = before + v;
            __syncthreads();
        }
        if (threadIdx.x == 0) a[n_tiles] = carry_s;
        __syncthreads();
    }
}
// ordered lists: nlpos[] (every newline); FASTA: stpos[] / stord[] / stcr[] (record start, newlines before it, '\r' before it)
__global__ void __launch_bounds__(BLOCK) k_index_emit(const uint8_t* __restrict__ bytes, uint32_t n, int fasta, int aligned,
                                                      const uint32_t* __restrict__ pre_nl, const uint32_t* __restrict__ pre_st, const uint32_t* __restrict__ pre_cr,
                                                      uint32_t* __restrict__ nlpos, uint32_t* __restrict__ stpos, uint32_t* __restrict__ stord,
                                                      uint32_t* __restrict__ stcr) {
    __shared__ unsigned long long wsum[BLOCK / 32];
    const uint32_t g0 = blockIdx.x * IDX_TILE + threadIdx.x * IDX_PER;
    const Masks m = span_masks(bytes, g0, n, fasta != 0, aligned != 0);
    const uint32_t st = fasta ? start_mask(bytes, g0, n, m) : 0u;
    // one block scan of the three per-thread counts (<= 32 each, <= 8192 per tile: 14 bits per field)
    const unsigned long long v = (unsigned long long)__popc(m.nl) | ((unsigned long long)__popc(st) << 14) | ((unsigned long long)__popc(m.cr) << 28);
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= d) incl += t; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned long long before = incl - v;
    for (int wi = 0; wi < (int)(threadIdx.x >> 5); wi++) before += wsum[wi];
    uint32_t o_nl = pre_nl[blockIdx.x] + (uint32_t)(before & 0x3FFFu);
    uint32_t mk = m.nl;
    while (mk) { const int b = __ffs((int)mk) - 1; mk &= mk - 1; nlpos[o_nl++] = g0 + b; }
    if (st) {
        const uint32_t nl0 = pre_nl[blockIdx.x] + (uint32_t)(before & 0x3FFFu), cr0 = pre_cr[blockIdx.x] + (uint32_t)((before >> 28) & 0x3FFFu);
        uint32_t o_st = pre_st[blockIdx.x] + (uint32_t)((before >> 14) & 0x3FFFu);
        mk = st;
        while (mk) {
            const int b = __ffs((int)mk) - 1; mk &= mk - 1;
            const uint32_t below = (1u << b) - 1u;
            stpos[o_st] = g0 + b; stord[o_st] = nl0 + __popc(m.nl & below); stcr[o_st] = cr0 + __popc(m.cr & below);
            o_st++;
        }
    }
}

__device__ __forceinline__ uint32_t trim_cr_end(const uint8_t* __restrict__ bytes, uint32_t b, uint32_t e) {
    return (e > b && bytes[e - 1] == '\r') ? e - 1 : e;      // parser/utils.rs:12-18
}

// one thread per complete FASTQ record (4 newlines): BufferPosition + validate (fastq.rs:17-65,240-285)
__global__ void __launch_bounds__(BLOCK) k_fastq_records(const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ nlpos,
                                                         uint32_t n_complete, ntg_record* __restrict__ recs,
                                                         uint8_t* __restrict__ errkind, uint32_t* __restrict__ first_err) {
    uint32_t r = blockIdx.x * BLOCK + threadIdx.x;
    if (r >= n_complete) return;
    uint32_t start = r ? nlpos[4 * r - 1] + 1 : 0;
    uint32_t seq = nlpos[4 * r] + 1, sep = nlpos[4 * r + 1] + 1, qual = nlpos[4 * r + 2] + 1, end = nlpos[4 * r + 3];
    ntg_record o;
    o.start = start;
    o.id_b = start + 1; o.id_e = trim_cr_end(bytes, start + 1, seq - 1);
    o.seq_b = seq; o.seq_e = trim_cr_end(bytes, seq, sep - 1);
    o.qual_b = qual; o.qual_e = trim_cr_end(bytes, qual, end);
    o.all_e = end;
    o.num_bases = o.seq_e - o.seq_b;
    o.line = 1 + 4ull * r;                                    // fastq.rs:116,411-415
    recs[r] = o;
    uint8_t ek = 0;
    if (bytes[start] != '@') ek = NTG_EINVALID_START;
    else if (bytes[sep] != '+') ek = NTG_EINVALID_SEPARATOR;
    else if (o.seq_e - o.seq_b != o.qual_e - o.qual_b) ek = NTG_EUNEQUAL_LENGTHS;
    errkind[r] = ek;
    if (ek) atomicMin(first_err, r);
}

// one thread per FASTA record start (fasta.rs:16-108,190-195).  stord / stcr: newlines / '\r' bytes in front of each start.
__global__ void __launch_bounds__(BLOCK) k_fasta_records(const uint8_t* __restrict__ bytes, uint32_t n,
                                                         const uint32_t* __restrict__ stpos, const uint32_t* __restrict__ stord,
                                                         const uint32_t* __restrict__ stcr, uint32_t n_starts,
                                                         const uint32_t* __restrict__ nlpos, uint32_t n_nl, uint32_t n_cr,
                                                         ntg_record* __restrict__ recs, uint32_t* __restrict__ last_bad,
                                                         uint32_t* __restrict__ first_le) {
    uint32_t r = blockIdx.x * BLOCK + threadIdx.x;
    if (r >= n_starts) return;
    uint32_t start = stpos[r];
    bool is_last = (r + 1 == n_starts);
    uint32_t ord = stord[r];                                  // newlines before `start` (none at start: it is '>')
    uint32_t first_nl = ord < n_nl ? nlpos[ord] : 0xFFFFFFFFu;
    uint32_t last;
    if (!is_last) last = stpos[r + 1] - 1;                    // the '\n' in front of the next '>'
    else {
        // EOF: a newline that is the final byte is not pushed; seq_pos empty => UnexpectedEnd (fasta.rs:205-213,348-356)
        if (first_nl == 0xFFFFFFFFu || first_nl == n - 1) { *last_bad = 1; first_nl = n; }
        last = (bytes[n - 1] == '\n') ? n - 1 : n;
    }
    ntg_record o;
    o.start = start;
    o.id_b = start + 1; o.id_e = trim_cr_end(bytes, start + 1, first_nl < n ? first_nl : n);
    if (last > first_nl) { o.seq_b = first_nl + 1; o.seq_e = trim_cr_end(bytes, first_nl + 1, last); }
    else { o.seq_b = o.seq_e = (first_nl < n ? first_nl : n); }
    o.qual_b = o.qual_e = 0;
    o.all_e = last;
    uint32_t sb = (uint32_t)o.seq_b, se = (uint32_t)o.seq_e;
    // num_bases = bytes of [sb, se) that are neither '\n' nor '\r' (fasta.rs:102-107), from the ordinals at the record starts:
    //   newlines before sb = ord + 1 (the header's newline); before se = those before `last` (the bytes of [se, last) are a trimmed '\r')
    //   '\r' before sb = stcr[r] + those of the header line (counted here: header lines are short); before se = those before the
    //   next start (or all of the window) minus the trimmed one
    uint64_t nb = 0;
    if (se > sb) {
        const uint32_t nl_before_last = is_last ? n_nl - ((bytes[n - 1] == '\n') ? 1u : 0u) : stord[r + 1] - 1u;
        const uint32_t nl_in = nl_before_last - (ord + 1u);
        uint32_t cr_hdr = 0;
        for (uint32_t q = start; q < sb; q++) cr_hdr += bytes[q] == '\r';
        const uint32_t cr_before_se = (is_last ? n_cr : stcr[r + 1]) - (last - se);
        const uint32_t cr_in = cr_before_se - (stcr[r] + cr_hdr);
        nb = (uint64_t)(se - sb) - nl_in - cr_in;
    }
    o.num_bases = nb;
    o.line = 1 + (uint64_t)ord;
    recs[r] = o;
    // line_ending(): taken from the first record whose all() contains a newline (fasta.rs:358-360, utils.rs:106-117)
    if (first_nl < last) atomicMin(&first_le[0], r);
}
}  // namespace parse

struct RecordsPriv {                       // the table's pinned buffer goes back to the context's pool when the table is freed
    struct Recs {
        ntg_record* p = nullptr; size_t cap = 0;
        std::shared_ptr<PinPool> pool;
        bool alloc(size_t rows) {               // true: failed
            void* q = pool->take((rows ? rows : 1) * sizeof(ntg_record), &cap);
            p = static_cast<ntg_record*>(q);
            return q == nullptr;
        }
    } recs;
    ~RecordsPriv() { if (recs.p) recs.pool->give(recs.p, recs.cap); }
};

// Host-side access to a few input bytes: straight from the caller's host copy when there is one,
// else small device-to-host reads (device-resident inputs).
struct Peek {
    const uint8_t* host; const uint8_t* dev;
    bool get(uint64_t b, size_t len, uint8_t* dst) const {
        if (len == 0) return true;
        if (host) { std::memcpy(dst, host + b, len); return true; }
        return cudaMemcpy(dst, dev + b, len, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    uint8_t at(uint64_t p) const { uint8_t v = 0; get(p, 1, &v); return v; }
};
// parser/utils.rs:106-117 on all() of the first record: `first_nl` is the first newline inside it
static int host_line_ending(const Peek& pk, uint64_t b, uint64_t e, uint64_t first_nl) {
    if (first_nl < b || first_nl >= e) return NTG_LE_NONE;
    return (first_nl > b && pk.at(first_nl - 1) == '\r') ? NTG_LE_WINDOWS : NTG_LE_UNIX;
}
// ErrorPosition.id: first space-delimited token of the id (fastq.rs:287-303)
static void host_error_id(const Peek& pk, uint64_t start, uint64_t seq, ntg_parse_error* err) {
    err->has_id = 0; err->id[0] = 0;
    if (seq - start > 1) {
        uint64_t b = start + 1, e = seq - 1;
        size_t l = (size_t)(e - b);
        if (l > sizeof(err->id) - 1) l = sizeof(err->id) - 1;
        uint8_t tmp[sizeof(err->id)];
        pk.get(b, l, tmp);
        if (l == (size_t)(e - b) && l > 0 && tmp[l - 1] == '\r') l--;
        size_t sp = 0;
        while (sp < l && tmp[sp] != ' ') sp++;
        std::memcpy(err->id, tmp, sp); err->id[sp] = 0;
        err->has_id = 1;
    }
}

// One window of a stream.  `force_format` = 0: sniff (the window is the start of the stream), else the format of the stream
// this window continues.  `at_eof` = false: the stream continues behind the window - only records that are complete inside
// it are delivered (FASTQ: four newlines; FASTA: followed by another record start), no end-of-stream rule runs, and
// *consumed = offset of the first byte not covered by a delivered record (where the next window must start).  Offsets and
// line numbers are relative to the window.  keep_drecs / keep_dbytes: the caller takes over the device copies.
static int run_parse_device(ntg_ctx* ctx, const uint8_t* bytes, const uint8_t* dev_in, size_t n, ntg_records** out,
                            DevBuf<ntg_record>* keep_drecs, int force_format = 0, bool at_eof = true, uint64_t* consumed = nullptr,
                            DevBuf<uint8_t>* keep_dbytes = nullptr) {
    using namespace parse;
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output pointer");
    *out = nullptr;
    if (n && !bytes && !dev_in) return ntg_set_error(ctx, NTG_EINVAL, "null input");
    const Peek pk{bytes, dev_in};
    if (n >= 0xFFFFFFF0ull) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "ntg_parse_fastx: feed at most 4 GiB per call");
    auto* res = new ntg_records();
    auto* priv = new RecordsPriv();
    priv->recs.pool = ctx->pinpool;
    std::memset(res, 0, sizeof(*res));
    res->_priv = priv;
    auto done = [&](int st) { if (st != NTG_OK) { delete priv; delete res; } else *out = res; return st; };

    if (consumed) *consumed = 0;
    if (force_format) {
        res->format = force_format;
        if (n == 0) return done(NTG_OK);
    } else {
        // sniff: parse_fastx_reader / get_fastx_reader (parser/mod.rs:85-93,37-46)
        if (n < 2) { res->error.kind = NTG_EEMPTY_FILE; return done(NTG_OK); }
        const uint8_t b0 = pk.at(0);
        if (b0 == '>') res->format = NTG_FMT_FASTA;
        else if (b0 == '@') res->format = NTG_FMT_FASTQ;
        else { res->error.kind = NTG_EUNKNOWN_FORMAT; return done(NTG_OK); }
    }
    res->error.format = res->format;
    const bool fasta = res->format == NTG_FMT_FASTA;
    const uint32_t n32 = (uint32_t)n;

    // device buffers: the context's grow-only scratch, except the copies a caller takes over (keep_dbytes / keep_drecs)
    ScratchPool& pool = ctx->scratch;
    struct { const uint8_t* p; } dbytes{dev_in};
    if (!dev_in) {
        if (keep_dbytes) { if (keep_dbytes->alloc(n)) return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed")); dbytes.p = keep_dbytes->p; }
        else {
            dbytes.p = (const uint8_t*)pool.get(ScratchPool::BYTES, n);
            if (!dbytes.p) return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
        }
    }
    const uint32_t n_tiles = (uint32_t)((n + IDX_TILE - 1) / IDX_TILE);
    uint32_t* counts = (uint32_t*)pool.get(ScratchPool::COUNTS, 3 * ((size_t)n_tiles + 1) * sizeof(uint32_t));
    uint32_t* misc = (uint32_t*)pool.get(ScratchPool::MISC, 64);
    if (!counts || !misc) return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
    uint32_t *cnt_nl = counts, *cnt_st = counts + (n_tiles + 1), *cnt_cr = counts + 2 * ((size_t)n_tiles + 1);
#define PCUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return done(ntg_set_error(ctx, NTG_ECUDA, "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(_e))); } while (0)
#define PTRY(expr) do { int _s = (expr); if (_s != NTG_OK) return done(_s); } while (0)
    if (!dev_in) PCUDA(cudaMemcpyAsync(const_cast<uint8_t*>(dbytes.p), bytes, n, cudaMemcpyHostToDevice, ctx->stream));
    const int aligned = ((uintptr_t)dbytes.p & 15) == 0;
    k_index_count<<<n_tiles, BLOCK, 0, ctx->stream>>>(dbytes.p, n32, fasta ? 1 : 0, aligned, cnt_nl, cnt_st, cnt_cr);
    k_index_scan<<<1, 1024, 0, ctx->stream>>>(cnt_nl, cnt_st, cnt_cr, n_tiles, fasta ? 3 : 1);
    ctx->launches += 2;
    uint32_t n_nl = 0, n_st = 0, n_cr = 0;
    PCUDA(cudaMemcpyAsync(&n_nl, cnt_nl + n_tiles, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (fasta) {
        PCUDA(cudaMemcpyAsync(&n_st, cnt_st + n_tiles, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PCUDA(cudaMemcpyAsync(&n_cr, cnt_cr + n_tiles, 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PCUDA(cudaStreamSynchronize(ctx->stream));
    struct { uint32_t* p; } nlpos{nullptr}, stpos{nullptr}, stord{nullptr}, stcr{nullptr};
    nlpos.p = (uint32_t*)pool.get(ScratchPool::NLPOS, ((size_t)n_nl + 1) * 4);
    if (fasta) {
        stpos.p = (uint32_t*)pool.get(ScratchPool::STPOS, ((size_t)n_st + 1) * 4);
        stord.p = (uint32_t*)pool.get(ScratchPool::STORD, ((size_t)n_st + 1) * 4);
        stcr.p = (uint32_t*)pool.get(ScratchPool::STCR, ((size_t)n_st + 1) * 4);
    }
    if (!nlpos.p || (fasta && (!stpos.p || !stord.p || !stcr.p))) return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
    k_index_emit<<<n_tiles, BLOCK, 0, ctx->stream>>>(dbytes.p, n32, fasta ? 1 : 0, aligned, cnt_nl, cnt_st, cnt_cr, nlpos.p, stpos.p, stord.p, stcr.p);
    ctx->launches++;

    // record table on the device: n_recs_cap rows
    struct { ntg_record* p; } drecs{nullptr};
    auto alloc_recs = [&](size_t rows) -> bool {
        if (keep_drecs) { if (keep_drecs->alloc(rows)) return false; drecs.p = keep_drecs->p; return true; }
        drecs.p = (ntg_record*)pool.get(ScratchPool::RECS, (rows ? rows : 1) * sizeof(ntg_record));
        return drecs.p != nullptr;
    };
    uint32_t le_rec = 0;     // index of the first record whose all() contains a newline (FASTQ: always record 0)
    if (!fasta) {
        // -------------------------------------------------------------------------- FASTQ
        uint32_t n_complete = n_nl / 4, rem = n_nl % 4;
        struct { uint8_t* p; } errkind{(uint8_t*)pool.get(ScratchPool::ERRKIND, (size_t)n_complete + 1)};
        struct { uint32_t* p; } first_err{misc};
        if (!alloc_recs((size_t)n_complete + 1) || !errkind.p)      // (+1: a last record without newline)
            return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
        PCUDA(cudaMemsetAsync(first_err.p, 0xFF, 4, ctx->stream));
        if (n_complete) {
            k_fastq_records<<<grid_for(n_complete), BLOCK, 0, ctx->stream>>>(dbytes.p, nlpos.p, n_complete, drecs.p, errkind.p, first_err.p);
            ctx->launches++;
        }
        PCUDA(cudaGetLastError());
        uint32_t ferr = 0xFFFFFFFFu;
        uint32_t tailnl[4] = {0, 0, 0, 0};     // [0] = last newline of the last complete record, [1..rem] = trailing newlines
        PCUDA(cudaMemcpyAsync(&ferr, first_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (n_complete) PCUDA(cudaMemcpyAsync(&tailnl[0], nlpos.p + (4 * (size_t)n_complete - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (rem) PCUDA(cudaMemcpyAsync(&tailnl[1], nlpos.p + 4 * (size_t)n_complete, 4 * rem, cudaMemcpyDeviceToHost, ctx->stream));
        PCUDA(cudaStreamSynchronize(ctx->stream));

        uint64_t n_ok = n_complete;
        ntg_record eof_rec; bool have_eof_rec = false;
        if (ferr != 0xFFFFFFFFu) {
            // first invalid complete record ends the stream (fastq.rs:243,253,277)
            n_ok = ferr;
            uint8_t ek = 0; ntg_record bad;
            PCUDA(cudaMemcpy(&ek, errkind.p + ferr, 1, cudaMemcpyDeviceToHost));
            PCUDA(cudaMemcpy(&bad, drecs.p + ferr, sizeof(bad), cudaMemcpyDeviceToHost));
            res->error.kind = ek; res->error.record_index = ferr;
            res->error.line = 1 + 4ull * ferr + (ek == NTG_EINVALID_SEPARATOR ? 2 : 0);     // fastq.rs:246,256,281
            if (ek != NTG_EINVALID_START) host_error_id(pk, bad.start, bad.seq_b, &res->error);
        } else if (!at_eof) {
            if (consumed) *consumed = n_complete ? (uint64_t)tailnl[0] + 1 : 0;
        } else {
            // end of stream: check_end (fastq.rs:337-356)
            uint64_t start = n_complete ? (uint64_t)tailnl[0] + 1 : 0;
            uint64_t line = 1 + 4ull * n_complete;
            if (rem == 3) {
                uint64_t seq = (uint64_t)tailnl[1] + 1, sep = (uint64_t)tailnl[2] + 1, qual = (uint64_t)tailnl[3] + 1, end = n;
                auto trim = [&](uint64_t b, uint64_t e) { return (e > b && pk.at(e - 1) == '\r') ? e - 1 : e; };
                ntg_record o;
                o.start = start; o.id_b = start + 1; o.id_e = trim(start + 1, seq - 1);
                o.seq_b = seq; o.seq_e = trim(seq, sep - 1); o.qual_b = qual; o.qual_e = trim(qual, end);
                o.all_e = end; o.num_bases = o.seq_e - o.seq_b; o.line = line;
                int ek = 0;
                if (pk.at(start) != '@') ek = NTG_EINVALID_START;
                else if (pk.at(sep) != '+') ek = NTG_EINVALID_SEPARATOR;
                else if (o.seq_e - o.seq_b != o.qual_e - o.qual_b) ek = NTG_EUNEQUAL_LENGTHS;
                if (ek) {
                    res->error.kind = ek; res->error.record_index = n_complete;
                    res->error.line = line + (ek == NTG_EINVALID_SEPARATOR ? 2 : 0);
                    if (ek != NTG_EINVALID_START) host_error_id(pk, start, seq, &res->error);
                } else { eof_rec = o; have_eof_rec = true; }
            } else if (start < n || rem) {
                // leftover must consist solely of empty / "\r" lines (fastq.rs:346-350)
                bool blank = true;
                uint64_t ls = start;
                for (uint32_t i = 0; i <= rem && blank; i++) {
                    uint64_t le = (i < rem) ? (uint64_t)tailnl[1 + i] : n;
                    uint64_t len = le - ls;
                    if (len > 1 || (len == 1 && pk.at(ls) != '\r')) blank = false;
                    ls = le + 1;
                }
                if (!blank) {
                    res->error.kind = NTG_EUNEXPECTED_END; res->error.record_index = n_complete;
                    res->error.line = line + rem;                       // search_pos as line offset (fastq.rs:352-355)
                    if (rem > 0) host_error_id(pk, start, (uint64_t)tailnl[1] + 1, &res->error);
                }
            }
        }
        uint64_t total = n_ok + (have_eof_rec ? 1 : 0);
        if (priv->recs.alloc(total)) return done(ntg_set_error(ctx, NTG_ENOMEM, "pinned allocation failed"));
        if (n_ok) { PCUDA(cudaMemcpyAsync(priv->recs.p, drecs.p, n_ok * sizeof(ntg_record), cudaMemcpyDeviceToHost, ctx->stream)); PCUDA(cudaStreamSynchronize(ctx->stream)); }
        if (have_eof_rec) {
            priv->recs.p[n_ok] = eof_rec;
            PCUDA(cudaMemcpy(drecs.p + n_ok, &eof_rec, sizeof(eof_rec), cudaMemcpyHostToDevice));      // (callers that keep the device table)
        }
        res->n_records = total; res->records = priv->recs.p;
        // FastxReader::position() after the last next(): position of the last record attempted (fastq.rs:411-415)
        {
            const uint64_t idx = (ferr != 0xFFFFFFFFu) ? ferr : n_complete;
            res->final_line = 1 + 4ull * idx;
            res->final_byte = (idx == 0) ? 0 : (idx <= n_ok ? priv->recs.p[idx - 1].all_e + 1 : (uint64_t)tailnl[0] + 1);
        }
    } else {
        // -------------------------------------------------------------------------- FASTA
        struct { uint32_t* p; } last_bad{misc + 1}, first_le{misc + 2};
        if (!alloc_recs(n_st)) return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
        PCUDA(cudaMemsetAsync(last_bad.p, 0, 4, ctx->stream));
        PCUDA(cudaMemsetAsync(first_le.p, 0xFF, 4, ctx->stream));
        k_fasta_records<<<grid_for(n_st), BLOCK, 0, ctx->stream>>>(dbytes.p, n32, stpos.p, stord.p, stcr.p, n_st, nlpos.p, n_nl, n_cr, drecs.p, last_bad.p, first_le.p);
        ctx->launches++;
        PCUDA(cudaGetLastError());
        uint32_t bad = 0;
        PCUDA(cudaMemcpyAsync(&bad, last_bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PCUDA(cudaMemcpyAsync(&le_rec, first_le.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PCUDA(cudaStreamSynchronize(ctx->stream));
        uint64_t total = n_st - (bad ? 1 : 0);
        if (priv->recs.alloc(n_st)) return done(ntg_set_error(ctx, NTG_ENOMEM, "pinned allocation failed"));
        PCUDA(cudaMemcpyAsync(priv->recs.p, drecs.p, (size_t)n_st * sizeof(ntg_record), cudaMemcpyDeviceToHost, ctx->stream)); PCUDA(cudaStreamSynchronize(ctx->stream));
        res->n_records = total; res->records = priv->recs.p;
        if (!at_eof) {
            // the last record start of the window is not delivered: the record may continue behind the window
            const ntg_record& l = priv->recs.p[n_st - 1];
            res->n_records = n_st - 1;
            res->final_line = l.line; res->final_byte = l.start;
            if (consumed) *consumed = l.start;
        } else if (bad) {
            const ntg_record& b = priv->recs.p[n_st - 1];
            res->error.kind = NTG_EUNEXPECTED_END; res->error.record_index = n_st - 1;
            res->error.line = b.line;                                   // fasta.rs:348-356
            res->final_line = b.line; res->final_byte = b.start;
        } else {
            const ntg_record& l = priv->recs.p[n_st - 1];
            res->final_line = l.line; res->final_byte = l.start;
        }
    }
#undef PCUDA
#undef PTRY
    if (le_rec < res->n_records) {
        // the first newline inside that record is the one ending its header line
        const ntg_record& lr = res->records[le_rec];
        uint64_t q = lr.id_e;
        if (pk.at(q) == '\r') q++;                               // id() had a '\r' trimmed
        res->line_ending = host_line_ending(pk, lr.start, lr.all_e, q);
    }
    return done(NTG_OK);
}
