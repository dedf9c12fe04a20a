// fused_host.cuh — host orchestration of the fused tallies path.  Part of the unity build.
#pragma once
#include "fused.cuh"
#include "parse.cuh"

struct FusedControl {           // device-resident control block, zeroed per call
    unsigned long long tallies[16];
    uint32_t flags;
    uint32_t tickets[59];       // one ticket per kernel launch of a call (chunked feeds use several)
};
static_assert(sizeof(FusedControl) == 128 + 4 + 59 * 4, "layout");
constexpr int FUSED_MAX_LAUNCHES = 59;
constexpr uint32_t SLOT_SHIFT = 16;      // 65 536 slots (16 MiB) + 65 536 words: far more than the tiles in flight plus the look-back reach

struct FusedState {
    fused::TileSlot* slots = nullptr;        // ring of 1 << SLOT_SHIFT tile slots + look-back words (never cleared: generations)
    unsigned long long* cw = nullptr;
    FusedControl* ctrl = nullptr;
    fused::SState* final_state = nullptr;
    FusedControl* h_ctrl = nullptr;          // pinned
    uint32_t epoch = 0;
    int max_ctas = 0;
    // pending call
    bool pending = false;
    fused::Params P{};
    ntg_tally_config cfg{};
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_ready = nullptr;
    cudaEvent_t ev_chunk[FUSED_MAX_LAUNCHES] = {};
    uint8_t* feed_buf = nullptr; size_t feed_cap = 0;   // device staging for host feeds
    const uint8_t* host_bytes = nullptr;                // when the call was fed from host memory
    uint32_t general_tile_bytes = 0;                    // tile size for a re-run without speculation
};

typedef void (*fused_kernel_t)(const fused::Params, const uint64_t, const uint64_t, const uint32_t, uint32_t*);
// kernel variants: three constant-folded headline shapes + the generic ones
#define NTG_FUSED_KERNELS(X) \
    X((k_fused<1, true, 11, 31, 21>)) X((k_fused<1, true, 11, 21, 11>)) X((k_fused<1, false, 0, 31, 0>)) \
    X((k_fused<2, false, 0, 0, 0>)) X((k_fused<1, false, 0, 0, 0>)) X((k_fused<1, true, 11, 0, 0>)) X((k_fused<1, true, 0, 0, 0>)) \
    X((k_fused<2, false, 0, 51, 0>))
static fused_kernel_t pick_fused_kernel(uint32_t k, uint32_t m, bool has_query) {
    using namespace fused;
    if (has_query) {                           // the query count lives in the generic walkers only
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

static int fused_init(ntg_ctx* ctx) {
    if (ctx->fused) return NTG_OK;
    auto* st = new FusedState();
    ctx->fused = st;
    NTG_CUDA(ctx, cudaMalloc((void**)&st->ctrl, sizeof(FusedControl)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->final_state, sizeof(fused::SState)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->slots, (size_t(1) << SLOT_SHIFT) * sizeof(fused::TileSlot)));
    NTG_CUDA(ctx, cudaMalloc((void**)&st->cw, (size_t(1) << SLOT_SHIFT) * sizeof(unsigned long long)));
    NTG_CUDA(ctx, cudaMemsetAsync(st->slots, 0, (size_t(1) << SLOT_SHIFT) * sizeof(fused::TileSlot), ctx->stream));
    NTG_CUDA(ctx, cudaMemsetAsync(st->cw, 0, (size_t(1) << SLOT_SHIFT) * sizeof(unsigned long long), ctx->stream));
    NTG_CUDA(ctx, cudaMallocHost((void**)&st->h_ctrl, sizeof(FusedControl)));
    NTG_CUDA(ctx, cudaEventCreate(&st->ev_k0));
    NTG_CUDA(ctx, cudaEventCreate(&st->ev_k1));
    NTG_CUDA(ctx, cudaEventCreateWithFlags(&st->ev_ready, cudaEventDisableTiming));
    for (auto& e : st->ev_chunk) NTG_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
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
    return NTG_OK;
}
static void fused_destroy(ntg_ctx* ctx) {
    FusedState* st = ctx->fused;
    if (!st) return;
    cudaFree(st->slots); cudaFree(st->cw); cudaFree(st->ctrl); cudaFree(st->final_state); cudaFreeHost(st->h_ctrl); cudaFree(st->feed_buf);
    if (st->ev_k0) cudaEventDestroy(st->ev_k0);
    if (st->ev_k1) cudaEventDestroy(st->ev_k1);
    if (st->ev_ready) cudaEventDestroy(st->ev_ready);
    for (auto& e : st->ev_chunk) if (e) cudaEventDestroy(e);
    delete st; ctx->fused = nullptr;
}

static int check_tally_cfg(ntg_ctx* ctx, const ntg_tally_config* cfg) {
    if (!cfg) return ntg_set_error(ctx, NTG_EINVAL, "null config");
    if (cfg->k == 0 || cfg->k > 64) return ntg_set_error(ctx, NTG_EINVAL, "k must be in 1..64");
    if (cfg->m && (cfg->k > 32 || cfg->m > cfg->k)) return ntg_set_error(ctx, NTG_EINVAL, "minimizers need 1 <= m <= k <= 32");
    if (cfg->has_query)
        for (uint32_t i = 0; i < cfg->k; i++)
            if (host_luts().code[cfg->query[i]] > 3 || (cfg->query[i] & 0x20)) return ntg_set_error(ctx, NTG_EINVAL, "query must be k bases of ACGT");
    return NTG_OK;
}

// Prepare a call over n device-resident bytes; launches nothing yet.
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

static int fused_begin(ntg_ctx* ctx, const uint8_t* dbytes, size_t n, int format, const ntg_tally_config* cfg, uint32_t tile_bytes,
                       bool allow_spec = true) {
    NTG_TRY(fused_init(ctx));
    FusedState* st = ctx->fused;
    if (st->pending) return ntg_set_error(ctx, NTG_EINVAL, "a tally call is already pending: collect it first");
    st->general_tile_bytes = tile_bytes;
    if ((reinterpret_cast<uintptr_t>(dbytes) & 15) != 0) return ntg_set_error(ctx, NTG_EINVAL, "device pointer must be 16-byte aligned");
    const uint64_t num_tiles = (n + tile_bytes - 1) / tile_bytes;
    // a fresh range of slot generations for this call: the previous call used epoch .. epoch + (its tiles >> SLOT_SHIFT)
    st->epoch = (st->epoch + 1 + (uint32_t)(st->P.num_tiles >> SLOT_SHIFT)) & 0x3FFFFFFFu;
    if (st->epoch + (num_tiles >> SLOT_SHIFT) + 2 >= 0x3FFFFFFFull) {       // (once per 2^30 calls) start the generations over
        NTG_CUDA(ctx, cudaMemsetAsync(st->slots, 0, (size_t(1) << SLOT_SHIFT) * sizeof(fused::TileSlot), ctx->stream));
        NTG_CUDA(ctx, cudaMemsetAsync(st->cw, 0, (size_t(1) << SLOT_SHIFT) * sizeof(unsigned long long), ctx->stream));
        st->epoch = 1;
    }
    NTG_CUDA(ctx, cudaMemsetAsync(st->ctrl, 0, sizeof(FusedControl), ctx->stream));
    fused::Params& P = st->P;
    P.bytes = dbytes; P.n = n; P.num_tiles = num_tiles; P.slots = st->slots; P.cw = st->cw; P.slot_mask = (1u << SLOT_SHIFT) - 1; P.slot_shift = SLOT_SHIFT; P.gmin = 0; P.ticket = nullptr;
    P.tallies = st->ctrl->tallies; P.flags = &st->ctrl->flags; P.final_state = st->final_state;
    P.k = cfg->k; P.m = cfg->m; P.w = cfg->m ? cfg->k - cfg->m + 1 : 1; P.format = format; P.has_query = cfg->has_query ? 1 : 0; P.one = 1; P.tile_bytes = tile_bytes;
    P.spec = (allow_spec && format == NTG_FMT_FASTQ && !(cfg->flags & NTG_TALLY_NO_SPECULATION)) ? 1 : 0;
    P.q_lo = P.q_hi = 0;
    if (cfg->has_query)
        for (uint32_t i = 0; i < cfg->k; i++) {
            P.q_hi = (P.q_hi << 2) | (P.q_lo >> 62);
            P.q_lo = (P.q_lo << 2) | host_luts().code[cfg->query[i]];
        }
    st->cfg = *cfg;
    st->pending = true;
    return NTG_OK;
}
// Launch the fused kernel over tiles [tb, te) as launch number `li` of this call.
static int fused_launch(ntg_ctx* ctx, uint64_t tb, uint64_t te, int li) {
    FusedState* st = ctx->fused;
    if (te <= tb) return NTG_OK;
    uint64_t nt = te - tb;
    unsigned grid = (unsigned)(nt < (uint64_t)st->max_ctas ? nt : (uint64_t)st->max_ctas);
    fused_kernel_t kf = pick_fused_kernel(st->P.k, st->P.m, st->P.has_query != 0);
    kf<<<grid, fused::NT, sizeof(fused::Smem), ctx->stream>>>(st->P, tb, te, st->epoch, &st->ctrl->tickets[li]);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    return NTG_OK;
}
static int fused_finish_enqueue(ntg_ctx* ctx) {
    FusedState* st = ctx->fused;
    fused::k_finalize<<<1, 1, 0, ctx->stream>>>(st->P);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    NTG_CUDA(ctx, cudaMemcpyAsync(st->h_ctrl, st->ctrl, sizeof(FusedControl), cudaMemcpyDeviceToHost, ctx->stream));
    return NTG_OK;
}

// forward: device-side parse used by the exact fallback (parse.cuh)
static int run_parse_device(ntg_ctx* ctx, const uint8_t* host_bytes, const uint8_t* dbytes, size_t n, ntg_records** out,
                            DevBuf<ntg_record>* keep_drecs);

static void tallies_from_ctrl(const FusedControl* c, ntg_tallies* out) {
    std::memset(out, 0, sizeof(*out));
    out->n_records = c->tallies[0]; out->n_bases = c->tallies[1]; out->n_kmers = c->tallies[2]; out->n_not_rc = c->tallies[3];
    out->kmer_sum_lo = c->tallies[4]; out->kmer_sum_hi = c->tallies[5]; out->n_query = c->tallies[6];
    out->n_minimizers = c->tallies[7]; out->minimizer_sum = c->tallies[8];
    for (int i = 0; i < 5; i++) out->reserved[2 + i] = c->tallies[9 + i];   // producer cycle accounting of the warp-specialised kernel
    out->reserved[5] = c->tallies[14]; out->reserved[6] = c->tallies[15];   // (walker wait / work cycles replace two of them)
}

static int fused_collect(ntg_ctx* ctx, ntg_tallies* out, ntg_parse_error* err, float* fused_kernel_ms) {
    FusedState* st = ctx->fused;
    if (!st || !st->pending) return ntg_set_error(ctx, NTG_EINVAL, "no pending tally call");
    st->pending = false;
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (fused_kernel_ms) NTG_CUDA(ctx, cudaEventElapsedTime(fused_kernel_ms, st->ev_k0, st->ev_k1));
    if (err) std::memset(err, 0, sizeof(*err));
    if (err) err->format = st->P.format;
    uint32_t fast_flags = st->h_ctrl->flags;
    if (fast_flags == 0) { tallies_from_ctrl(st->h_ctrl, out); return NTG_OK; }
    const uint32_t ws_flags = (fast_flags & fused::FLAG_SPEC_MISS) ? fast_flags : 0;
    if (fast_flags & fused::FLAG_SPEC_MISS) {
        // a speculated FASTQ line phase was wrong: one plain pass (no speculation) over the same bytes
        const ntg_tally_config cfg = st->cfg;
        NTG_TRY(fused_begin(ctx, st->P.bytes, st->P.n, st->P.format, &cfg, st->general_tile_bytes, /*allow_spec=*/false));
        int s2 = fused_launch(ctx, 0, st->P.num_tiles, 0);
        if (s2 == NTG_OK) s2 = fused_finish_enqueue(ctx);
        st->pending = false;
        NTG_TRY(s2);
        NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        fast_flags = st->h_ctrl->flags;
        if (fast_flags == 0) { tallies_from_ctrl(st->h_ctrl, out); out->reserved[1] = ws_flags; return NTG_OK; }
    }

    // ---- exact fallback: record table on the device, then one thread per delivered record
    ntg_records* recs = nullptr;
    DevBuf<ntg_record> drecs;
    NTG_TRY(run_parse_device(ctx, st->host_bytes, st->P.bytes, st->P.n, &recs, &drecs));
    if (err) *err = recs->error;
    NTG_CUDA(ctx, cudaMemsetAsync(st->ctrl, 0, sizeof(FusedControl), ctx->stream));
    if (recs->n_records) {
        unsigned grid = (unsigned)((recs->n_records + 127) / 128);
        if (grid > (unsigned)ctx->sm_count * 16) grid = ctx->sm_count * 16;
        if (st->P.k > 32) fused::k_tally_records<2, false><<<grid, 128, 0, ctx->stream>>>(st->P, drecs.p, recs->n_records);
        else if (st->P.m == 0) fused::k_tally_records<1, false><<<grid, 128, 0, ctx->stream>>>(st->P, drecs.p, recs->n_records);
        else fused::k_tally_records<1, true><<<grid, 128, 0, ctx->stream>>>(st->P, drecs.p, recs->n_records);
        ctx->launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (!e) e = cudaMemcpyAsync(st->h_ctrl, st->ctrl, sizeof(FusedControl), cudaMemcpyDeviceToHost, ctx->stream);
    if (!e) e = cudaStreamSynchronize(ctx->stream);
    ntg_records_free(recs);
    if (e) return ntg_set_error(ctx, NTG_ECUDA, "fallback: %s", cudaGetErrorString(e));
    tallies_from_ctrl(st->h_ctrl, out);
    out->reserved[1] = ws_flags;
    out->reserved[0] = fast_flags;          // why the exact path ran: 1 parse error, 2 newline-dense tile, 4 whitespace run > halo
    return NTG_OK;
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
