// fused_host.cuh — host orchestration of the fused tallies path.  Part of the unity build.
//
// One pass = a sequence of LAUNCHES of fused::k_fused over consecutive tile ranges of one byte stream.  The look-back state
// (slot ring, generations) carries across launches, so a stream can be fed segment by segment through two/three device
// buffers of SEG bytes: inputs of any size run in bounded device memory (the reference streams through a growable
// buffer: src/parser/fastq.rs:312-384, src/parser/utils.rs:34-49).  Each launch writes its own control block (tallies,
// flags, first-error key); the host adds the blocks up.
//
// Error semantics (iterator: records before the first error are delivered, fastq.rs:243,253,277): the kernel reports the
// start byte E of the first failing record; the host REPLAYS the stream truncated at E (a clean end of stream) and
// classifies the error with the record scanner on a window at E.  Inputs the single pass cannot take (newline-dense
// tiles, whitespace runs longer than the halo) go to the exact record-table path, window by window.
#pragma once
#include <functional>

#include "fused.cuh"
#include "fastq_warp.cuh"
#include "parse.cuh"

struct LaunchCtl {                    // device-resident control block of one launch (mirrored in pinned host memory)
    unsigned long long tallies[16];
    unsigned long long err_key;       // ~0 = no parse error
    unsigned long long fin[4];        // k_finalize: [0] kind of an end-of-stream error (FASTA), [1] its line, [2] its record index
    uint32_t flags, ticket;
    uint32_t pad[4];
};
static_assert(sizeof(LaunchCtl) == 192, "layout");
constexpr int NCTL = 8;                  // control blocks in flight (a streamed pass keeps at most three launches unchecked)
constexpr int CTL_REDO = NCTL - 1;       // block reserved for replays
constexpr uint32_t SLOT_SHIFT = 16;      // 65 536 slots (16 MiB) + 65 536 words: far more than the tiles in flight plus the look-back reach
constexpr size_t STREAM_BACK = size_t(1) << 20;      // bytes of history kept in front of every streamed segment (lines that cross into it)
constexpr size_t STREAM_SEG = size_t(64) << 20;      // segment size of streamed passes (rounded down to whole tiles)
constexpr int NSEG = 3;                              // device segments: a flagged launch and its neighbours stay intact until checked

// launch shape of a pass, chosen from the first bytes of the input: tile size of fused::k_fused; chunk size and fragments per
// sequence line of the record-owned fast path when the input is short-read FASTQ
struct PassShape { uint32_t tile_bytes = 0; bool fq_ok = false; uint32_t cb = 0, frags = 1; };

struct PassResult {                   // sum over the launches of one pass
    unsigned long long tallies[16] = {};
    unsigned long long err_key = ~0ull;
    unsigned long long fin[4] = {};
    uint32_t flags = 0;
    uint64_t fq_next = ~0ull;             // record-owned passes: where the first record that was not complete begins (the tail)
    void add(const LaunchCtl& c) {
        for (int i = 0; i < 16; i++) tallies[i] += c.tallies[i];
        if (c.err_key < err_key) err_key = c.err_key;
        if (c.fin[0]) for (int i = 0; i < 3; i++) fin[i] = c.fin[i];
        flags |= c.flags;
    }
};

struct FusedState {
    fused::TileSlot* slots = nullptr;        // ring of 1 << SLOT_SHIFT tile slots + look-back words (never cleared: generations)
    unsigned long long* cw = nullptr;
    fused::SState* final_state = nullptr;
    unsigned long long* fa_totals = nullptr;  // FASTA: record starts / newlines of the pass in flight
    LaunchCtl* ctl = nullptr;                // NCTL blocks, device
    LaunchCtl* h_ctl = nullptr;              // NCTL blocks, pinned
    cudaEvent_t ev_done[NCTL] = {};
    cudaEvent_t ev_copy[NSEG] = {};
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    uint32_t epoch = 0;
    uint64_t epoch_tiles = 0;                // tiles of the pass that owns `epoch`
    int max_ctas = 0;
    uint8_t* seg[NSEG] = {};                 // streamed passes: STREAM_BACK + seg_cap bytes each
    size_t seg_cap = 0;
    unsigned long long* reduce_buf = nullptr;   // 16 x u64: send/receive buffer of the in-stream tallies all-reduce
    unsigned long long* h_reduce = nullptr;     // pinned mirror
    // resident call pending between enqueue and collect
    bool pending = false, pending_reduce = false, pending_fq = false;
    fused::Params P{};
    ntg_tally_config cfg{};
    int format = 0;
    uint32_t general_tile_bytes = 0;
    // sniff cache of the resident entry point: (pointer, size) -> format + tile size, verified on the device (byte 0)
    uint64_t sniff_ptr = 0; size_t sniff_n = 0; int sniff_format = 0; uint32_t sniff_flags = 0; PassShape sniff_shape;
    uint64_t pend_dptr = 0; size_t pend_n = 0;
    // record-owned FASTQ fast path (fastq_warp.cuh)
    int fq_max_ctas = 0;
    uint8_t* fq_info = nullptr; size_t fq_info_cap = 0;        // per-chunk bytes of the launch in flight (ring of NCTL regions)
    uint32_t* fq_fix_list = nullptr; uint8_t* fq_fix_phase = nullptr;
    unsigned long long* fq_words = nullptr;                    // [0, NCTL): start word per launch slot; [NCTL, 2 NCTL): carry words
    uint32_t* fq_counters = nullptr;                           // per slot: ticket, fix ticket, fix count, pad
    bool fq_disable = false;                                   // NTG_TALLY_NO_FASTPATH
    // spectrum binding of the passes to come (spectrum.cuh sets and clears it)
    uint32_t* sp_dense = nullptr; unsigned long long* sp_keys = nullptr; uint32_t* sp_counts = nullptr; uint64_t sp_mask = 0; uint32_t* sp_overflow = nullptr;
};

typedef void (*fused_kernel_t)(const fused::Params, const uint64_t, const uint64_t, const uint32_t, uint32_t*);
// kernel variants: three constant-folded headline shapes + the generic ones
#define NTG_FUSED_KERNELS(X) \
    X((k_fused<1, true, 11, 31, 21>)) X((k_fused<1, true, 11, 21, 11>)) X((k_fused<1, false, 0, 31, 0>)) \
    X((k_fused<2, false, 0, 0, 0>)) X((k_fused<1, false, 0, 0, 0>)) X((k_fused<1, true, 11, 0, 0>)) X((k_fused<1, true, 0, 0, 0>)) \
    X((k_fused<2, false, 0, 51, 0>))
static fused_kernel_t pick_fused_kernel(uint32_t k, uint32_t m, bool has_query) {
    using namespace fused;
    if (has_query) {                           // (spectrum passes come here too: the generic walker is the one that counts)                           // the query count lives in the generic walkers only
        if (k > 32) return k_fused<2, false, 0, 0, 0>;
        if (m == 0) return k_fused<1, false, 0, 0, 0>;
        return (k - m + 1 == 11) ? k_fused<1, true, 11, 0, 0> : k_fused<1, true, 0, 0, 0>;
    }
    if (k == 51 && m == 0) return k_fused<2, false, 0, 51, 0>;
    if (k == 31 && m == 21) return k_fused<1, true, 11, 31, 21>;
    if (k == 21 && m == 11) return k_fused<1, true, 11, 21, 11>;
    if (k == 31 && m == 0) return k_fused<1, false, 0, 31, 0>;
    if (k > 32) return k_fused<2, false, 0, 0, 0>;
    if (m == 0) return k_fused<1, false, 0, 0, 0>;
    if (k - m + 1 == 11) return k_fused<1, true, 11, 0, 0>;
    return k_fused<1, true, 0, 0, 0>;
}

// ---- record-owned FASTQ fast path: kernel table, launch-shape choice --------------------------------------------------
typedef void (*fq_kernel_t)(const fqw::Params);
#define NTG_FQ_KERNELS(X) \
    X((k_records<1, true, 11, 31, 21>)) X((k_records<1, true, 11, 21, 11>)) X((k_records<1, false, 0, 31, 0>)) \
    X((k_records<2, false, 0, 0, 0>)) X((k_records<1, false, 0, 0, 0>)) X((k_records<1, true, 11, 0, 0>)) X((k_records<1, true, 0, 0, 0>)) \
    X((k_records<2, false, 0, 51, 0>))
static fq_kernel_t pick_fq_kernel(uint32_t k, uint32_t m, bool generic) {
    using namespace fqw;
    if (generic) {
        if (k > 32) return k_records<2, false, 0, 0, 0>;
        if (m == 0) return k_records<1, false, 0, 0, 0>;
        return (k - m + 1 == 11) ? k_records<1, true, 11, 0, 0> : k_records<1, true, 0, 0, 0>;
    }
    if (k == 51 && m == 0) return k_records<2, false, 0, 51, 0>;
    if (k == 31 && m == 21) return k_records<1, true, 11, 31, 21>;
    if (k == 21 && m == 11) return k_records<1, true, 11, 21, 11>;
    if (k == 31 && m == 0) return k_records<1, false, 0, 31, 0>;
    if (k > 32) return k_records<2, false, 0, 0, 0>;
    if (m == 0) return k_records<1, false, 0, 0, 0>;
    if (k - m + 1 == 11) return k_records<1, true, 11, 0, 0>;
    return k_records<1, true, 0, 0, 0>;
}
#ifndef NTG_FQ_MAX_FRAGS
#define NTG_FQ_MAX_FRAGS 1
#endif
constexpr uint32_t FQ_FIX_CAP = 1u << 16;
constexpr size_t FQ_INFO_SLOT = size_t(1) << 24;            // chunk bytes per launch slot (16 Mi chunks = 160 GB of text)
static int fq_init(ntg_ctx* ctx) {
    FusedState* st = ctx->fused;
    using namespace fqw;
#define NTG_X(k) (fq_kernel_t)k,
    fq_kernel_t ks[] = {NTG_FQ_KERNELS(NTG_X)};
#undef NTG_X
    int occ_min = 1 << 30;
    for (auto kf : ks) {
        NTG_CUDA(ctx, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(fqw::Smem)));
        int occ = 0;
        NTG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf, fqw::NT, sizeof(fqw::Smem)));
        if (occ < 1) return ntg_set_error(ctx, NTG_ECUDA, "record-owned FASTQ kernel does not fit on an SM");
        occ_min = occ < occ_min ? occ : occ_min;
    }
    st->fq_max_ctas = occ_min * ctx->sm_count;
    NTG_CUDA(ctx, cudaMalloc((void**)&st->fq_info, FQ_INFO_SLOT * 2));
    st->fq_info_cap = FQ_INFO_SLOT;
    NTG_CUDA(ctx, cudaMalloc((void**)&st->fq_fix_list, FQ_FIX_CAP * sizeof(uint32_t)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->fq_fix_phase, FQ_FIX_CAP));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->fq_words, 2 * NCTL * sizeof(unsigned long long)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->fq_counters, 4 * NCTL * sizeof(uint32_t)));
    return NTG_OK;
}
// Launch shape from the first bytes of a FASTQ stream: bytes per record -> fragments per sequence line and chunk size such
// that a chunk holds about 31 items (one per lane).  false: not a short-read shape (records above 1.3 KB), use fused::k_fused.
static bool fq_shape(const uint8_t* sample, size_t ns, uint32_t* cb, uint32_t* frags) {
    size_t nl = 0;
    for (size_t i = 0; i < ns; i++) nl += sample[i] == '\n';
    if (nl < 8) return false;
    const double rec = 4.0 * (double)ns / (double)nl;
    // (fragments > 1 are implemented — reads up to ~600 bp — but measured slower than fused::k_fused on the 250 bp shape:
    //  365 vs 440 Gbases/s, every fragment pays its own k-1 warm-up bases.  Until chunks can be larger, one fragment only.)
    uint32_t f = 1;
    while (f <= 4 && rec * 31.0 / f > fqw::CBMAX) f *= 2;
    if (f > NTG_FQ_MAX_FRAGS || rec * 2.5 > fqw::SLACK) return false;
    uint32_t c = ((uint32_t)(rec * 31.0 / f) + 15u) & ~15u;
    if (c < 1024u) c = 1024u;
    if (c > (uint32_t)fqw::CBMAX) c = fqw::CBMAX;
    *cb = c; *frags = f;
    return true;
}

static int fused_init(ntg_ctx* ctx) {
    if (ctx->fused) return NTG_OK;
    auto* st = new FusedState();
    ctx->fused = st;
    NTG_CUDA(ctx, cudaMalloc((void**)&st->ctl, NCTL * sizeof(LaunchCtl)));
    NTG_CUDA(ctx, cudaMallocHost((void**)&st->h_ctl, NCTL * sizeof(LaunchCtl)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->final_state, sizeof(fused::SState)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->fa_totals, 2 * sizeof(unsigned long long)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->slots, (size_t(1) << SLOT_SHIFT) * sizeof(fused::TileSlot)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->cw, (size_t(1) << SLOT_SHIFT) * sizeof(unsigned long long)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->reduce_buf, 16 * sizeof(unsigned long long)));
    NTG_CUDA(ctx, cudaMallocHost((void**)&st->h_reduce, 16 * sizeof(unsigned long long)));
    NTG_CUDA(ctx, cudaMemsetAsync(st->slots, 0, (size_t(1) << SLOT_SHIFT) * sizeof(fused::TileSlot), ctx->stream));
    NTG_CUDA(ctx, cudaMemsetAsync(st->cw, 0, (size_t(1) << SLOT_SHIFT) * sizeof(unsigned long long), ctx->stream));
    NTG_CUDA(ctx, cudaEventCreate(&st->ev_k0));
    NTG_CUDA(ctx, cudaEventCreate(&st->ev_k1));
    for (auto& e : st->ev_done) NTG_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : st->ev_copy) NTG_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    using namespace fused;
#define NTG_X(k) (fused_kernel_t)k,
    fused_kernel_t ks[] = {NTG_FUSED_KERNELS(NTG_X)};
#undef NTG_X
    int occ_min = 1 << 30;
    for (auto kf : ks) {
        NTG_CUDA(ctx, cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(fused::Smem)));
        int occ = 0;
        NTG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf, fused::NT, sizeof(fused::Smem)));
        if (occ < 1) return ntg_set_error(ctx, NTG_ECUDA, "fused kernel does not fit on an SM");
        occ_min = occ < occ_min ? occ : occ_min;
    }
    st->max_ctas = occ_min * ctx->sm_count;     // persistent grid: every CTA resident (look-back needs forward progress)
    NTG_TRY(fq_init(ctx));
    return NTG_OK;
}
static void fused_destroy(ntg_ctx* ctx) {
    FusedState* st = ctx->fused;
    if (!st) return;
    cudaFree(st->slots); cudaFree(st->cw); cudaFree(st->ctl); cudaFree(st->final_state); cudaFree(st->fa_totals); cudaFreeHost(st->h_ctl);
    cudaFree(st->reduce_buf); cudaFreeHost(st->h_reduce);
    cudaFree(st->fq_info); cudaFree(st->fq_fix_list); cudaFree(st->fq_fix_phase); cudaFree(st->fq_words); cudaFree(st->fq_counters);
    for (auto& p : st->seg) cudaFree(p);
    if (st->ev_k0) cudaEventDestroy(st->ev_k0);
    if (st->ev_k1) cudaEventDestroy(st->ev_k1);
    for (auto& e : st->ev_done) if (e) cudaEventDestroy(e);
    for (auto& e : st->ev_copy) if (e) cudaEventDestroy(e);
    delete st; ctx->fused = nullptr;
}

static int check_tally_cfg(ntg_ctx* ctx, const ntg_tally_config* cfg) {
    if (!cfg) return ntg_set_error(ctx, NTG_EINVAL, "null config");
    if (cfg->k == 0 || cfg->k > 64) return ntg_set_error(ctx, NTG_EINVAL, "k must be in 1..64");
    if (cfg->m && (cfg->k > 32 || cfg->m > cfg->k)) return ntg_set_error(ctx, NTG_EINVAL, "minimizers need 1 <= m <= k <= 32");
    if (cfg->qmask_score > 255) return ntg_set_error(ctx, NTG_EINVAL, "qmask_score is a quality byte (0 = off, 1..255)");
    if (cfg->has_query)
        for (uint32_t i = 0; i < cfg->k; i++)
            if (host_luts().code[cfg->query[i]] > 3 || (cfg->query[i] & 0x20)) return ntg_set_error(ctx, NTG_EINVAL, "query must be k bases of ACGT");
    return NTG_OK;
}

// Tile size: about NT sequence lines per tile, so that every walker thread gets exactly one line.
// `sample` = the first bytes of the input (host copy), used to estimate the line period.
static uint32_t pick_tile_bytes(const uint8_t* sample, size_t ns, int format) {
    size_t nl = 0;
    for (size_t i = 0; i < ns; i++) nl += sample[i] == '\n';
    double seqline_period;                       // bytes of input per sequence line
    if (nl < 8) return fused::TILE;              // long lines: they are cut into SEG-byte pieces anyway
    if (format == NTG_FMT_FASTQ) seqline_period = 4.0 * (double)ns / (double)nl;
    else {
        size_t hdr = 0;
        for (size_t i = 0; i + 1 < ns; i++) hdr += (sample[i] == '\n' && sample[i + 1] == '>');
        size_t seql = nl > hdr ? nl - hdr : 1;
        seqline_period = (double)ns / (double)seql;
        if ((double)ns / (double)nl > fused::SEG) return fused::TILE;
    }
    double want = 0.97 * (format == NTG_FMT_FASTQ ? fused::NTW : fused::NT) * seqline_period;   // FASTQ: the coordinator warp does not walk
    uint32_t tb = (uint32_t)(want / 256.0) * 256u;
    if (tb < 16384u) tb = 16384u;
    if (tb > (uint32_t)fused::TILE) tb = fused::TILE;
    return tb;
}

// Start a pass: a fresh range of slot generations and the launch-invariant kernel parameters.
// max_tiles: upper bound of the pass's tile count (wrap check); tiles_now: tiles known so far (streamed passes update epoch_tiles as they go)
static int fused_begin_pass(ntg_ctx* ctx, int format, const ntg_tally_config* cfg, uint32_t tile_bytes, bool allow_spec, uint64_t max_tiles,
                            uint64_t tiles_now) {
    NTG_TRY(fused_init(ctx));
    FusedState* st = ctx->fused;
    // the previous pass used generations epoch .. epoch + (its tiles >> SLOT_SHIFT); this one may use up to max_tiles >> SLOT_SHIFT
    // more.  Generations live in 30 bits: before they would wrap (and meet stale slots of old passes) the ring is cleared.
    uint64_t next = (uint64_t)st->epoch + 1 + (st->epoch_tiles >> SLOT_SHIFT);
    if (next + (max_tiles >> SLOT_SHIFT) + 2 >= 0x3FFFFFFFull) {
        NTG_CUDA(ctx, cudaMemsetAsync(st->slots, 0, (size_t(1) << SLOT_SHIFT) * sizeof(fused::TileSlot), ctx->stream));
        NTG_CUDA(ctx, cudaMemsetAsync(st->cw, 0, (size_t(1) << SLOT_SHIFT) * sizeof(unsigned long long), ctx->stream));
        next = 1;
    }
    st->epoch = (uint32_t)next;
    st->epoch_tiles = tiles_now;
    fused::Params& P = st->P;
    P = fused::Params{};
    P.slots = st->slots; P.cw = st->cw; P.slot_mask = (1u << SLOT_SHIFT) - 1; P.slot_shift = SLOT_SHIFT; P.final_state = st->final_state;
    P.fa_totals = st->fa_totals;
    NTG_CUDA(ctx, cudaMemsetAsync(st->fa_totals, 0, 2 * sizeof(unsigned long long), ctx->stream));
    P.k = cfg->k; P.m = cfg->m; P.w = cfg->m ? cfg->k - cfg->m + 1 : 1; P.format = format; P.has_query = cfg->has_query ? 1 : 0; P.one = 1; P.tile_bytes = tile_bytes;
    P.qmask = format == NTG_FMT_FASTQ ? (cfg->qmask_score & 0xFFu) : 0u;
    P.spec = (allow_spec && format == NTG_FMT_FASTQ && !(cfg->flags & NTG_TALLY_NO_SPECULATION)) ? 1 : 0;
    P.q_lo = P.q_hi = 0;
    if (cfg->has_query)
        for (uint32_t i = 0; i < cfg->k; i++) {
            P.q_hi = (P.q_hi << 2) | (P.q_lo >> 62);
            P.q_lo = (P.q_lo << 2) | host_luts().code[cfg->query[i]];
        }
    P.sp_dense = st->sp_dense; P.sp_keys = st->sp_keys; P.sp_counts = st->sp_counts; P.sp_mask = st->sp_mask; P.sp_overflow = st->sp_overflow;
    st->cfg = *cfg; st->format = format; st->general_tile_bytes = tile_bytes;
    return NTG_OK;
}

// Enqueue one launch of the pass over tiles [tb, te): `base` is the (virtual) address of stream byte 0, `gmin` the lowest
// stream position readable through it, `n_vis` the number of stream bytes that exist for this launch (the end of its last
// tile, or the stream length when `final`).  Results go to control block `ci` and its pinned mirror; ev_done[ci] fires
// when the mirror is valid.  `reduce` (final launches): k_finalize leaves the tallies in the all-reduce send buffer.
static int fused_enqueue_launch(ntg_ctx* ctx, const uint8_t* base, uint64_t gmin, uint64_t n_vis, uint64_t tb, uint64_t te, bool final, int ci,
                                bool reduce = false) {
    FusedState* st = ctx->fused;
    LaunchCtl* c = st->ctl + ci;
    NTG_CUDA(ctx, cudaMemsetAsync(c, 0, sizeof(LaunchCtl), ctx->stream));
    NTG_CUDA(ctx, cudaMemsetAsync(&c->err_key, 0xFF, sizeof(c->err_key), ctx->stream));
    fused::Params P = st->P;
    P.bytes = base; P.gmin = gmin; P.n = n_vis; P.num_tiles = final ? te : (uint64_t(1) << 62);
    P.tallies = c->tallies; P.flags = &c->flags; P.err_key = &c->err_key; P.fin = c->fin; P.ticket = nullptr;
    P.reduce_buf = reduce ? st->reduce_buf : nullptr;
    if (te > tb) {
        const uint64_t nt = te - tb;
        const unsigned grid = (unsigned)(nt < (uint64_t)st->max_ctas ? nt : (uint64_t)st->max_ctas);
        fused_kernel_t kf = pick_fused_kernel(P.k, P.m, P.has_query != 0 || P.sp_dense || P.sp_keys);
        kf<<<grid, fused::NT, sizeof(fused::Smem), ctx->stream>>>(P, tb, te, st->epoch, &c->ticket);
        ctx->launches++;
        NTG_CUDA(ctx, cudaGetLastError());
    }
    if (final) {
        NTG_CUDA(ctx, cudaEventRecord(st->ev_k1, ctx->stream));
        fused::k_finalize<<<1, 1, 0, ctx->stream>>>(P);
        ctx->launches++;
        NTG_CUDA(ctx, cudaGetLastError());
    }
    if (!reduce) {
        NTG_CUDA(ctx, cudaMemcpyAsync(st->h_ctl + ci, c, sizeof(LaunchCtl), cudaMemcpyDeviceToHost, ctx->stream));
        NTG_CUDA(ctx, cudaEventRecord(st->ev_done[ci], ctx->stream));
    }
    return NTG_OK;
}

// Enqueue one launch of the record-owned fast path: records starting at the device word fq_words[slot_in] (0 for the first
// launch of a pass) and complete inside the first n_vis stream bytes.  k_verify leaves the start of the first record that was
// not (the next launch's start) in fq_words[slot_out] and in ctl.fin[3].  `span` bounds the bytes the launch can cover.
static int fq_enqueue_launch(ntg_ctx* ctx, const uint8_t* base, uint64_t n_vis, uint64_t span, uint32_t cb, uint32_t frags, uint64_t seq, bool final, int ci,
                             bool reduce = false, bool keep_start = false) {
    FusedState* st = ctx->fused;
    LaunchCtl* c = st->ctl + ci;
    const int wi = (int)(seq % NCTL), wo = (int)((seq + 1) % NCTL);
    const bool first = seq == 0 && !keep_start;
    NTG_CUDA(ctx, cudaMemsetAsync(c, 0, sizeof(LaunchCtl), ctx->stream));
    NTG_CUDA(ctx, cudaMemsetAsync(&c->err_key, 0xFF, sizeof(c->err_key), ctx->stream));
    const uint64_t n_chunks_max = span / cb + 3;
    if (n_chunks_max + 64 > st->fq_info_cap) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "launch too large for the record-owned path");
    uint8_t* info = st->fq_info + (size_t)(seq & 1) * st->fq_info_cap;
    NTG_CUDA(ctx, cudaMemsetAsync(info, 0, (n_chunks_max + 63) & ~uint64_t(15), ctx->stream));      // (k_verify reads whole 16-byte vectors)
    NTG_CUDA(ctx, cudaMemsetAsync(st->fq_counters + 4 * wi, 0, 4 * sizeof(uint32_t), ctx->stream));
    NTG_CUDA(ctx, cudaMemsetAsync(st->fq_words + NCTL + wi, 0xFF, sizeof(unsigned long long), ctx->stream));
    if (first) NTG_CUDA(ctx, cudaMemsetAsync(st->fq_words + wi, 0, sizeof(unsigned long long), ctx->stream));
    fqw::Params P{};
    P.W = st->P;
    P.W.tallies = c->tallies; P.W.flags = &c->flags; P.W.err_key = &c->err_key; P.W.fin = c->fin;
    P.bytes = base; P.start = st->fq_words + wi; P.n_vis = n_vis; P.cb = cb; P.frags = frags; P.info = info;
    P.ticket = st->fq_counters + 4 * wi; P.carry = st->fq_words + NCTL + wi;
    P.fix_list = nullptr; P.fix_count = nullptr; P.fix_phase = nullptr; P.fix_cap = FQ_FIX_CAP;
    fq_kernel_t kf = pick_fq_kernel(P.W.k, P.W.m, P.W.has_query != 0 || P.W.sp_dense || P.W.sp_keys);
    uint64_t want = (n_chunks_max + fqw::WARPS - 1) / fqw::WARPS;
    const unsigned grid = (unsigned)(want < (uint64_t)st->fq_max_ctas ? want : (uint64_t)st->fq_max_ctas);
    kf<<<grid, fqw::NT, sizeof(fqw::Smem), ctx->stream>>>(P);
    fqw::k_verify<<<1, fqw::VT, 0, ctx->stream>>>(info, P.start, cb, st->fq_fix_list, st->fq_fix_phase, st->fq_counters + 4 * wi + 2, FQ_FIX_CAP, &c->flags,
                                                  P.carry, n_vis, st->fq_words + wo, &c->fin[3]);
    P.fix_list = st->fq_fix_list; P.fix_count = st->fq_counters + 4 * wi + 2; P.fix_phase = st->fq_fix_phase; P.ticket = st->fq_counters + 4 * wi + 1;
    kf<<<(unsigned)ctx->sm_count, fqw::NT, sizeof(fqw::Smem), ctx->stream>>>(P);      // fix-up: chunks without a local guess (usually none)
    ctx->launches += 3;
    NTG_CUDA(ctx, cudaGetLastError());
    if (final) {
        NTG_CUDA(ctx, cudaEventRecord(st->ev_k1, ctx->stream));
        if (reduce) { fused::k_reduce_copy<<<1, 32, 0, ctx->stream>>>(c->tallies, &c->flags, &c->fin[3], n_vis, st->reduce_buf); ctx->launches++; }
    }
    if (!reduce) {
        NTG_CUDA(ctx, cudaMemcpyAsync(st->h_ctl + ci, c, sizeof(LaunchCtl), cudaMemcpyDeviceToHost, ctx->stream));
        NTG_CUDA(ctx, cudaEventRecord(st->ev_done[ci], ctx->stream));
    }
    return NTG_OK;
}

// how a pass runs: fused::k_fused without / with speculated FASTQ line phases, or the record-owned fast path (short-read FASTQ)
enum PassMode { MODE_NOSPEC = 0, MODE_SPEC = 1, MODE_FQ = 2 };
static PassShape pass_shape(const uint8_t* sample, size_t ns, int format, const ntg_tally_config* cfg) {
    PassShape sh;
    sh.tile_bytes = pick_tile_bytes(sample, ns, format);
    sh.fq_ok = format == NTG_FMT_FASTQ && !(cfg->flags & (NTG_TALLY_NO_FASTPATH | NTG_TALLY_NO_SPECULATION)) && fq_shape(sample, ns, &sh.cb, &sh.frags);
    return sh;
}

// ---- a pass over bytes resident in device memory: one launch -----------------------------------------------------
static int pass_resident(ntg_ctx* ctx, const uint8_t* dbytes, uint64_t n, int format, const ntg_tally_config* cfg, const PassShape& sh, int mode,
                         PassResult* out) {
    FusedState* st = ctx->fused;
    const uint32_t tile_bytes = sh.tile_bytes;
    const bool spec = mode != MODE_NOSPEC;
    if (mode == MODE_FQ) {
        NTG_TRY(fused_begin_pass(ctx, format, cfg, tile_bytes, true, 0, 0));
        NTG_TRY(fq_enqueue_launch(ctx, dbytes, n, n, sh.cb, sh.frags, 0, true, 0));
        NTG_CUDA(ctx, cudaEventSynchronize(st->ev_done[0]));
        *out = PassResult{};
        out->add(st->h_ctl[0]);
        out->fq_next = st->h_ctl[0].fin[3];
        return NTG_OK;
    }
    const uint64_t num_tiles = (n + tile_bytes - 1) / tile_bytes;
    NTG_TRY(fused_begin_pass(ctx, format, cfg, tile_bytes, spec, num_tiles, num_tiles));
    NTG_TRY(fused_enqueue_launch(ctx, dbytes, 0, n, 0, num_tiles, true, 0));
    NTG_CUDA(ctx, cudaEventSynchronize(st->ev_done[0]));
    *out = PassResult{};
    out->add(st->h_ctl[0]);
    return NTG_OK;
}

// ---- a pass over a host byte stream fed in segments ------------------------------------------------------------------
// Device memory: NSEG buffers of STREAM_BACK + seg_cap bytes.  Launch L uses buffer L % NSEG: its data behind STREAM_BACK
// bytes of history copied from the tail of the previous buffer (lines and records that cross into the segment).
struct SegmentFeed {
    ntg_ctx* ctx = nullptr;
    FusedState* st = nullptr;
    uint32_t TB = 0;
    uint64_t L = 0;                  // launches submitted
    uint64_t next_tile = 0;
    size_t prev_len = 0;
    uint64_t checked = 0;            // launches whose control block has been read back
    struct Rec { uint64_t tb, te, n_vis; size_t len; bool final; } recs[NCTL] = {};
    size_t seg_target = STREAM_SEG;  // bytes of whole tiles per launch (device-side inflate uses larger segments)
    bool fq = false; uint32_t cb = 0, frags = 1;    // record-owned fast path: launches are byte ranges, chained by their carry words
    uint64_t bytes_submitted = 0;

    int open(ntg_ctx* c, int format, const ntg_tally_config* cfg, uint32_t tile_bytes, bool spec, size_t seg_bytes = STREAM_SEG) {
        ctx = c;
        NTG_TRY(fused_init(ctx));
        st = ctx->fused;
        TB = tile_bytes;
        seg_target = seg_bytes;
        if (st->seg_cap < seg_bytes + fused::TILE) {
            NTG_CUDA(ctx, cudaDeviceSynchronize());
            for (auto& p : st->seg) { cudaFree(p); p = nullptr; }
            st->seg_cap = seg_bytes + fused::TILE;
            for (auto& p : st->seg)
                if (cudaMalloc((void**)&p, STREAM_BACK + st->seg_cap) != cudaSuccess) { cudaGetLastError(); st->seg_cap = 0; return ntg_set_error(ctx, NTG_ENOMEM, "device segment allocation failed"); }
        }
        NTG_TRY(fused_begin_pass(ctx, format, cfg, tile_bytes, spec, uint64_t(1) << 40, 0));
        // the copy stream must not overwrite the segments while kernels of an earlier call still read them
        NTG_CUDA(ctx, cudaEventRecord(st->ev_copy[0], ctx->stream));
        NTG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, st->ev_copy[0], 0));
        return NTG_OK;
    }
    uint64_t seg_tiles() const { return seg_target / TB; }
    // Submit the next `len` stream bytes (whole tiles unless final) from host memory.  Before buffer L % NSEG is reused
    // the launch that used it (L - NSEG) must have been checked: `check(j, ctl)` is called for every launch in order.
    // `produce` (optional): instead of an H2D copy of `src`, the segment's bytes are produced on the device by work the
    // callback enqueues on the copy stream: produce(data region of this segment, data region of the previous one, prev_len).
    template <typename Check>
    int submit(const uint8_t* src, size_t len, bool final, uint64_t n_total, Check&& check,
               const std::function<int(uint8_t*, const uint8_t*, size_t)>* produce = nullptr) {
        if (len > st->seg_cap) return ntg_set_error(ctx, NTG_EINVAL, "segment larger than the device segment buffer");
        while (L >= (uint64_t)(NSEG - 1) && checked + (NSEG - 1) <= L) NTG_TRY(check_next(check));      // frees buffer L % NSEG
        const int b = (int)(L % NSEG), pb = (int)((L + NSEG - 1) % NSEG), ci = (int)(L % (NCTL - 1));
        uint8_t* buf = st->seg[b];
        if (L > 0) NTG_CUDA(ctx, cudaMemcpyAsync(buf, st->seg[pb] + prev_len, STREAM_BACK, cudaMemcpyDeviceToDevice, ctx->copy_stream));
        if (produce) NTG_TRY((*produce)(buf + STREAM_BACK, st->seg[pb] + STREAM_BACK, prev_len));
        else if (len) NTG_CUDA(ctx, cudaMemcpyAsync(buf + STREAM_BACK, src, len, cudaMemcpyHostToDevice, ctx->copy_stream));
        NTG_CUDA(ctx, cudaEventRecord(st->ev_copy[b], ctx->copy_stream));
        NTG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, st->ev_copy[b], 0));
        if (L == 0) NTG_CUDA(ctx, cudaEventRecord(st->ev_k0, ctx->stream));
        if (fq) {
            const uint64_t start = bytes_submitted, n_vis = final ? n_total : start + len;
            NTG_TRY(fq_enqueue_launch(ctx, buf + STREAM_BACK - start, n_vis, len + 2 * (uint64_t)fqw::WIN, cb, frags, L, final, ci));
            recs[ci] = Rec{0, 0, n_vis, len, final};
            prev_len = len; bytes_submitted += len; L++;
            return NTG_OK;
        }
        const uint64_t tb = next_tile, start = tb * (uint64_t)TB;
        const uint64_t te = final ? (n_total + TB - 1) / TB : tb + len / TB;
        const uint64_t n_vis = final ? n_total : te * (uint64_t)TB;
        const uint64_t gmin = start > STREAM_BACK ? start - STREAM_BACK : 0;
        NTG_TRY(fused_enqueue_launch(ctx, buf + STREAM_BACK - start, gmin, n_vis, tb, te, final, ci));
        recs[ci] = Rec{tb, te, n_vis, len, final};
        prev_len = len; next_tile = te; bytes_submitted += len; L++;
        st->epoch_tiles = te;                                     // generations this pass has used so far
        return NTG_OK;
    }
    template <typename Check>
    int check_next(Check&& check) {
        const int ci = (int)(checked % (NCTL - 1));
        NTG_CUDA(ctx, cudaEventSynchronize(st->ev_done[ci]));
        const uint64_t j = checked++;
        return check(j, st->h_ctl[ci]);
    }
    template <typename Check>
    int drain(Check&& check) {
        while (checked < L) NTG_TRY(check_next(check));
        return NTG_OK;
    }
    // device address of stream byte `g` while the launch that carried it is still resident (one of the last NSEG launches)
    const uint8_t* device_addr(uint64_t j, uint64_t g) const {
        const Rec& r = recs[j % (NCTL - 1)];
        return st->seg[j % NSEG] + STREAM_BACK + (g - r.tb * (uint64_t)TB);
    }
};

static int pass_host(ntg_ctx* ctx, const uint8_t* bytes, uint64_t n, int format, const ntg_tally_config* cfg, const PassShape& sh, int mode,
                     PassResult* out) {
    SegmentFeed f;
    f.fq = mode == MODE_FQ; f.cb = sh.cb; f.frags = sh.frags;
    NTG_TRY(f.open(ctx, format, cfg, sh.tile_bytes, mode != MODE_NOSPEC));
    *out = PassResult{};
    auto check = [&](uint64_t, const LaunchCtl& c) { out->add(c); if (f.fq) out->fq_next = c.fin[3]; return (int)NTG_OK; };
    const uint64_t seg_bytes = f.fq ? (uint64_t)STREAM_SEG : f.seg_tiles() * (uint64_t)sh.tile_bytes;
    uint64_t off = 0;
    for (;;) {
        const bool final = n - off <= seg_bytes;
        const size_t len = (size_t)(final ? n - off : seg_bytes);
        NTG_TRY(f.submit(bytes + off, len, final, n, check));
        off += len;
        if (final) break;
    }
    return f.drain(check);
}

// forward: exact record-table path (windowed)
struct ByteSource { const uint8_t* host; const uint8_t* dev; uint64_t n; bool more_behind = false; };   // more_behind: the stream continues behind byte n
static int exact_tally(ntg_ctx* ctx, const ByteSource& src, int format, const ntg_tally_config* cfg, PassResult* out, ntg_parse_error* err);

static void tallies_from_pass(const PassResult& r, ntg_tallies* out) {
    std::memset(out, 0, sizeof(*out));
    out->n_records = r.tallies[0]; out->n_bases = r.tallies[1]; out->n_kmers = r.tallies[2]; out->n_not_rc = r.tallies[3];
    out->kmer_sum_lo = r.tallies[4]; out->kmer_sum_hi = r.tallies[5]; out->n_query = r.tallies[6];
    out->n_minimizers = r.tallies[7]; out->minimizer_sum = r.tallies[8];
    for (int i = 0; i < 5; i++) out->reserved[2 + i] = r.tallies[9 + i];   // NTG_STATS builds: per-CTA cycle accounting
    out->reserved[5] = r.tallies[14]; out->reserved[6] = r.tallies[15];
}

// The error of the record that starts at stream byte E (the kernel's first-error key), classified by the record scanner on a
// window at E: kind, line, id exactly as the reference reports them.  `recs_before` = records delivered before it.
static int classify_error_at(ntg_ctx* ctx, const ByteSource& src, uint64_t E, uint64_t recs_before, ntg_parse_error* err, bool* confirmed) {
    *confirmed = false;
    for (uint64_t W = uint64_t(1) << 20;; W <<= 4) {
        const uint64_t len = src.n - E < W ? src.n - E : W;
        const bool whole = E + len == src.n, at_eof = whole && !src.more_behind;
        if (len >= 0xFFFFFFF0ull) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "a record larger than 4 GiB");
        ntg_records* r = nullptr;
        NTG_TRY(run_parse_device(ctx, src.host ? src.host + E : nullptr, src.dev ? src.dev + E : nullptr, (size_t)len, &r, nullptr, NTG_FMT_FASTQ, at_eof));
        const bool found = r->error.kind != 0 && r->error.record_index == 0;
        const bool cut = !at_eof && r->n_records == 0 && r->error.kind == 0;       // the window does not hold the whole record yet
        if (found) {
            if (err) { *err = r->error; err->line += 4 * recs_before; err->record_index += recs_before; }
            *confirmed = true;
        }
        ntg_records_free(r);
        if (!cut || whole) return NTG_OK;                            // (whole && cut: the source does not hold the whole record)
    }
}

// The tail of a record-owned pass: the bytes from the first record that was not complete to the end of the stream (at most one
// record and blank lines) go through the exact record-table path; its tallies join, its error (if any) is the stream's.
static int fq_tail(ntg_ctx* ctx, const ByteSource& src, int format, const ntg_tally_config* cfg, PassResult* r, ntg_parse_error* err) {
    if (r->fq_next >= src.n) return NTG_OK;
    ByteSource tail{src.host ? src.host + r->fq_next : nullptr, src.dev ? src.dev + r->fq_next : nullptr, src.n - r->fq_next, src.more_behind};
    PassResult x; ntg_parse_error e2;
    NTG_TRY(exact_tally(ctx, tail, format, cfg, &x, &e2));
    const uint64_t before = r->tallies[0];
    for (int i = 0; i < 9; i++) r->tallies[i] += x.tallies[i];
    if (e2.kind && err) { *err = e2; err->record_index += before; err->line += 4 * before; }
    return NTG_OK;
}

static int tally_masked_copy(ntg_ctx* ctx, const ByteSource& src, int format, const ntg_tally_config* cfg, ntg_tallies* out, ntg_parse_error* err);

// Whole-input tallies with the reference's iterator semantics.  `run(n_eff, mode, &result)` makes one pass over the first
// n_eff bytes of the source (resident or streamed); every pass can be repeated because the caller still holds the bytes.
template <typename Run>
static int tally_whole(ntg_ctx* ctx, const ByteSource& src, int format, const ntg_tally_config* cfg, bool fq_ok, Run&& run, ntg_tallies* out,
                       ntg_parse_error* err) {
    if (err) { std::memset(err, 0, sizeof(*err)); err->format = format; }
    PassResult r;
    uint32_t spec_missed = 0;
    bool fq = false;
    const bool qmask = cfg->qmask_score != 0 && format == NTG_FMT_FASTQ;
    if (qmask && !fq_ok) return tally_masked_copy(ctx, src, format, cfg, out, err);       // (only the record-owned kernel masks in place)
    if (fq_ok) {
        // the record-owned fast path takes clean streams and streams whose only problem is a failing record; anything else
        // (a wrong phase guess, a record longer than its slack, a newline-dense chunk) is fused::k_fused's business
        NTG_TRY(run(src.n, MODE_FQ, &r));
        fq = r.flags == 0 || (r.flags == fused::FLAG_PARSE_ERROR && r.err_key != ~0ull);
    }
    if (!fq && qmask) return tally_masked_copy(ctx, src, format, cfg, out, err);
    if (!fq) {
        NTG_TRY(run(src.n, MODE_SPEC, &r));
        if (r.flags & fused::FLAG_SPEC_MISS) { spec_missed = r.flags; NTG_TRY(run(src.n, MODE_NOSPEC, &r)); }   // a speculated FASTQ line phase was wrong
    }
    if (r.flags == 0) {
        if (fq) NTG_TRY(fq_tail(ctx, src, format, cfg, &r, err));
        tallies_from_pass(r, out); out->reserved[1] = spec_missed | (fq ? NTG_RESERVED_FAST_PATH : 0);
        return NTG_OK;
    }
    const uint32_t why = r.flags;
    if (r.flags == fused::FLAG_PARSE_ERROR && format == NTG_FMT_FASTA) {
        // only the end-of-stream rule can fail (fasta.rs:348-356): the last record is not delivered, nothing of it was tallied
        tallies_from_pass(r, out);
        if (err) { err->kind = (int32_t)r.fin[0]; err->line = r.fin[1]; err->record_index = r.fin[2]; }
        out->reserved[1] = spec_missed;
        return NTG_OK;
    }
    if (r.flags == fused::FLAG_PARSE_ERROR && r.err_key != ~0ull) {
        // replay the stream truncated at the first failing record, then classify the failure there
        const uint64_t E = r.err_key >> 2;
        PassResult t;
        if (E >= 2) {
            bool done = false;
            if (fq) { NTG_TRY(run(E, MODE_FQ, &t)); done = t.flags == 0 && t.fq_next >= E; }
            if (!done && qmask) return tally_masked_copy(ctx, src, format, cfg, out, err);
            if (!done) {
                NTG_TRY(run(E, MODE_SPEC, &t));
                if (t.flags & fused::FLAG_SPEC_MISS) NTG_TRY(run(E, MODE_NOSPEC, &t));
            }
        }
        if (t.flags == 0) {
            bool confirmed = false;
            ntg_parse_error e2; std::memset(&e2, 0, sizeof(e2)); e2.format = format;
            NTG_TRY(classify_error_at(ctx, src, E, t.tallies[0], &e2, &confirmed));
            if (confirmed) {
                tallies_from_pass(t, out);
                if (err) *err = e2;
                out->reserved[1] = spec_missed | (fq ? NTG_RESERVED_FAST_PATH : 0);
                return NTG_OK;
            }
        }
    }
    // ---- exact path: record table on the device window by window, one thread per delivered record
    PassResult x;
    NTG_TRY(exact_tally(ctx, src, format, cfg, &x, err));
    tallies_from_pass(x, out);
    out->reserved[1] = spec_missed;
    out->reserved[0] = why;                 // why the exact path ran: 1 parse error, 2 newline-dense tile, 4 whitespace run > halo
    return NTG_OK;
}

static int exact_tally(ntg_ctx* ctx, const ByteSource& src, int format, const ntg_tally_config* cfg, PassResult* out, ntg_parse_error* err) {
    FusedState* st = ctx->fused;
    *out = PassResult{};
    if (err) { std::memset(err, 0, sizeof(*err)); err->format = format; }
    LaunchCtl* c = st->ctl + CTL_REDO;
    NTG_CUDA(ctx, cudaMemsetAsync(c, 0, sizeof(LaunchCtl), ctx->stream));
    fused::Params P{};
    P.k = cfg->k; P.m = cfg->m; P.w = cfg->m ? cfg->k - cfg->m + 1 : 1; P.format = format; P.has_query = cfg->has_query ? 1 : 0;
    if (cfg->has_query)
        for (uint32_t i = 0; i < cfg->k; i++) { P.q_hi = (P.q_hi << 2) | (P.q_lo >> 62); P.q_lo = (P.q_lo << 2) | host_luts().code[cfg->query[i]]; }
    P.tallies = c->tallies; P.flags = &c->flags;
    uint64_t pos = 0, rec_base = 0, line_base = 0;
    uint64_t W = uint64_t(1) << 30;                                   // window (the record scanner indexes with 32 bits)
    bool stop = false;
    while (pos < src.n && !stop) {
        const uint64_t len = src.n - pos < W ? src.n - pos : W;
        const bool at_eof = pos + len == src.n;
        ntg_records* recs = nullptr;
        DevBuf<ntg_record> drecs; DevBuf<uint8_t> dwin;
        uint64_t consumed = 0;
        NTG_TRY(run_parse_device(ctx, src.host ? src.host + pos : nullptr, src.dev ? src.dev + pos : nullptr, (size_t)len, &recs, &drecs, format, at_eof,
                                 &consumed, &dwin));
        if (recs->n_records) {
            P.bytes = src.dev ? src.dev + pos : dwin.p;
            unsigned grid = (unsigned)((recs->n_records + 127) / 128);
            if (grid > (unsigned)ctx->sm_count * 16) grid = ctx->sm_count * 16;
            if (P.k > 32) fused::k_tally_records<2, false><<<grid, 128, 0, ctx->stream>>>(P, drecs.p, recs->n_records);
            else if (P.m == 0) fused::k_tally_records<1, false><<<grid, 128, 0, ctx->stream>>>(P, drecs.p, recs->n_records);
            else fused::k_tally_records<1, true><<<grid, 128, 0, ctx->stream>>>(P, drecs.p, recs->n_records);
            ctx->launches++;
            cudaError_t e = cudaGetLastError();
            if (!e) e = cudaStreamSynchronize(ctx->stream);           // (the window buffers are freed at the end of this iteration)
            if (e) { ntg_records_free(recs); return ntg_set_error(ctx, NTG_ECUDA, "exact path: %s", cudaGetErrorString(e)); }
        }
        if (recs->error.kind) {
            if (err) { *err = recs->error; err->record_index += rec_base; err->line += line_base; }
            stop = true;
        } else if (!at_eof) {
            if (consumed == 0) {                                       // not even one complete record in the window: widen it
                ntg_records_free(recs);
                if (W >= 0xF0000000ull) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "a record larger than 3.75 GiB");
                W = W * 2 < 0xF0000000ull ? W * 2 : 0xF0000000ull;
                continue;
            }
            line_base += recs->final_line - 1;
        }
        rec_base += recs->n_records;
        pos = at_eof ? src.n : pos + consumed;
        ntg_records_free(recs);
    }
    NTG_CUDA(ctx, cudaMemcpyAsync(st->h_ctl + CTL_REDO, c, sizeof(LaunchCtl), cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->add(st->h_ctl[CTL_REDO]);
    out->flags = 0;
    return NTG_OK;
}

// quality_mask for inputs the record-owned kernel does not take (long reads, newline-dense files, a wrong phase guess): the text
// is copied on the device, every delivered record's low-quality bases are overwritten with 'N' through the record table
// (window by window), and the copy is tallied without the mask.
namespace fused {
__global__ void __launch_bounds__(128) k_mask_records(uint8_t* __restrict__ text, const ntg_record* __restrict__ recs, uint64_t n_recs, uint32_t score) {
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= n_recs) return;
    const ntg_record r = recs[w];
    const uint64_t n = r.seq_e - r.seq_b < r.qual_e - r.qual_b ? r.seq_e - r.seq_b : r.qual_e - r.qual_b;
    for (uint64_t j = lane; j < n; j += 32) if (text[r.qual_b + j] < score) text[r.seq_b + j] = 'N';
}
}  // namespace fused
static int tally_masked_copy(ntg_ctx* ctx, const ByteSource& src, int format, const ntg_tally_config* cfg, ntg_tallies* out, ntg_parse_error* err) {
    DevBuf<uint8_t> copy;
    if (copy.alloc(src.n + 16)) { cudaGetLastError(); return ntg_set_error(ctx, NTG_ENOMEM, "quality mask: no room for a device copy of %llu bytes", (unsigned long long)src.n); }
    NTG_CUDA(ctx, cudaMemcpyAsync(copy.p, src.host ? (const void*)src.host : (const void*)src.dev, src.n, src.host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    uint64_t pos = 0, W = uint64_t(1) << 30;
    while (pos < src.n) {
        const uint64_t len = src.n - pos < W ? src.n - pos : W;
        const bool at_eof = pos + len == src.n;
        ntg_records* recs = nullptr; DevBuf<ntg_record> drecs; uint64_t consumed = 0;
        NTG_TRY(run_parse_device(ctx, nullptr, copy.p + pos, (size_t)len, &recs, &drecs, format, at_eof, &consumed, nullptr));
        const uint64_t nrec = recs->n_records; const bool stop = recs->error.kind != 0;
        ntg_records_free(recs);
        if (nrec) {
            fused::k_mask_records<<<(unsigned)((nrec * 32 + 127) / 128), 128, 0, ctx->stream>>>(copy.p + pos, drecs.p, nrec, cfg->qmask_score);
            ctx->launches++;
            NTG_CUDA(ctx, cudaGetLastError());
            NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // (the record table is freed at the end of this iteration)
        }
        if (stop || at_eof) break;                                   // (records behind the first error are never delivered: leave them alone)
        if (consumed == 0) { if (W >= 0xF0000000ull) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "a record larger than 3.75 GiB"); W = W * 2 < 0xF0000000ull ? W * 2 : 0xF0000000ull; continue; }
        pos += consumed;
    }
    ntg_tally_config plain = *cfg; plain.qmask_score = 0;
    const uint8_t* d = copy.p;
    std::vector<uint8_t> sample(src.n < 65536 ? src.n : 65536);
    NTG_CUDA(ctx, cudaMemcpy(sample.data(), d, sample.size(), cudaMemcpyDeviceToHost));
    const PassShape sh = pass_shape(sample.data(), sample.size(), format, &plain);
    const ByteSource masked{nullptr, d, src.n, src.more_behind};
    auto run = [&](uint64_t n_eff, int mode, PassResult* r) { return pass_resident(ctx, d, n_eff, format, &plain, sh, mode, r); };
    return tally_whole(ctx, masked, format, &plain, sh.fq_ok, run, out, err);
}

static int sniff_format(ntg_ctx* ctx, uint8_t b0, size_t n, ntg_tallies* out, ntg_parse_error* err, int* format) {
    // parse_fastx_reader / get_fastx_reader (parser/mod.rs:85-93,37-46)
    *format = NTG_FMT_NONE;
    (void)ctx;
    if (n < 2) { if (out) std::memset(out, 0, sizeof(*out)); if (err) { std::memset(err, 0, sizeof(*err)); err->kind = NTG_EEMPTY_FILE; } return 1; }
    if (b0 == '>') *format = NTG_FMT_FASTA;
    else if (b0 == '@') *format = NTG_FMT_FASTQ;
    else { if (out) std::memset(out, 0, sizeof(*out)); if (err) { std::memset(err, 0, sizeof(*err)); err->kind = NTG_EUNKNOWN_FORMAT; } return 1; }
    return 0;
}
