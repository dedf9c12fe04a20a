// fused_ws.cuh — warp-specialised variant of the fused tallies kernel for the headline shape
// (FASTQ, short reads, constant-folded K / M).
//
// ncu on fused::k_fused (profiles/r1c): ~40 % of warp time sits at CTA barriers — every phase of a tile
// (load, newline scan, look-back, walk) is CTA-synchronous.  Here the phases are decoupled:
//
//   warp NW-1 (producer; highest issue priority): claims tiles, keeps one TMA bulk copy in flight, scans newlines, publishes the
//           aggregate, does the decoupled look-back, validates lines (fastq.rs:240-285), counts
//           n_records / n_bases, and hands the tile over as a "stage" (ring of NS shared-memory buffers);
//   warps 0..NW-2 (walkers): claim batches of 32 sequence lines from the CTA-wide item stream (a batch may
//           span two stages, so lanes stay full), run walk_fast<K,M>, reduce, and release the stages.
//
// The kernel only handles what the headline needs: anything else — a sequence line longer than SEG, a
// deleted byte inside a sequence line, more newlines than NLW in a tile, a stage with < 32 sequence
// lines, a parse error — raises FLAG_WS_BAIL / FLAG_PARSE_ERROR and the host re-runs the general kernel
// (fused::k_fused), which in turn owns the exact fallback.  Results are therefore always exact.
// Part of the unity build (ntgpu.cu).
#pragma once
#include "fused.cuh"

namespace fused_ws {
using namespace fused;

constexpr int NW = 6;                    // warps per CTA: 1 producer + 5 walkers; 3 CTAs per SM
constexpr int WNT = NW * 32;
constexpr int NS = 3;                    // stages in the ring
constexpr int TB = 20 * 1024;            // tile bytes (80 rows of 256 B)
constexpr int ROUNDS = (TB / ROWB + 31) / 32;   // 3
constexpr int NLW = 1024;                // newline capacity per tile
constexpr int CUMN = 64;                 // ring of cumulative item counts per stage sequence number
constexpr uint32_t NOT_ENDED = 0xFFFFFFFFu;

struct __align__(16) Stage {
    uint8_t halo[HALO];
    uint8_t tile[TB];
    uint16_t nl[NLW + 8];
    uint64_t t;                          // tile index held by this buffer
    uint32_t n_nl, n_nl_raw, i_first, n_items, avail;
    int32_t line0_lo;
    uint32_t line0_exact, line0_starts_here;
};
struct __align__(16) SmemWS {
    Stage st[NS];
    uint8_t lut[256];
    uint32_t rins[256];
    uint64_t tma_bar[NS];
    uint64_t tallies[9];                 // CTA-wide, flushed once at the end
    uint32_t cum_end[CUMN];              // items of this CTA's stream before the end of stage seq (mod CUMN)
    volatile uint32_t remaining[NS];     // items of the stage in buffer b not yet finished by the walkers
    volatile uint32_t ready_seq;         // number of stages published so far
    volatile uint32_t end_seq;           // sequence number of the end marker (NOT_ENDED until known)
    uint32_t claim;                      // next unclaimed item of the CTA's stream
    uint32_t flags;
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// fast-walker result of one item
struct ItemSums { uint64_t nk, nrc, ksum, msum; };

template <int K, int M>
__device__ __forceinline__ bool walk_item(const uint8_t* __restrict__ sb, const FastLuts& L, int ws, int b, ItemSums& o) {
    Acc a;                                 // walk_fast adds into an Acc; keep it local so nothing stays live across items
    if (!walk_fast<K, M>(sb, L, ws, b, a)) return false;
    o.nk = a.n_kmers; o.nrc = a.n_not_rc; o.ksum = a.ksum_lo; o.msum = a.msum;
    return true;
}

template <int K, int M>
__global__ void __launch_bounds__(WNT, 3) k_fused_ws(const Params P, const uint64_t tile_begin, const uint64_t tile_end,
                                                     const uint32_t epoch, uint32_t* __restrict__ ticket) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemWS& S = *reinterpret_cast<SmemWS*>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += WNT) {
        const uint8_t c = class_of(i);
        S.lut[i] = c;
        S.rins[i] = (3u - (c & 3u)) << (2 * (K - 1) - 32);
    }
    if (tid == 0) {
        for (int b = 0; b < NS; b++) { mbar_init(&S.tma_bar[b], 1); S.remaining[b] = 0; }
        fence_mbar_init();
        for (int i = 0; i < 9; i++) S.tallies[i] = 0;
        S.ready_seq = 0; S.end_seq = NOT_ENDED; S.claim = 0; S.flags = 0;
    }
    __syncthreads();

    if (warp == NW - 1) {     // the highest warp id wins issue arbitration: the producer must never starve
        // ================================================================== producer
        // Non-blocking state machine over the ring: claim+load -> scan+publish aggregate -> look-back+handover.
        // Between claiming a tile and publishing its aggregate the warp never waits on another tile's
        // look-back (otherwise delayed aggregates cascade from CTA to CTA and the walkers starve).
        uint32_t seq_claim = 0, seq_scan = 0, seq_done = 0, cum = 0, par[NS] = {0, 0, 0};
        uint32_t my_flags = 0;
        bool no_more = false;
        uint64_t acc_records = 0, acc_bases = 0;
        long long c_claim = 0, c_scan = 0, c_lbfail = 0, c_resolve = 0, c_idle = 0;   // cycle accounting (profiles/)
        auto claim_tile = [&]() -> uint64_t {
            uint32_t v = 0;
            if (lane == 0) v = (ld_acquire_u32(P.flags) & FLAG_WS_BAIL) ? 0xFFFFFFFFu : atomicAdd(ticket, 1u);
            v = __shfl_sync(0xffffffffu, v, 0);
            if (v == 0xFFFFFFFFu) return ~0ull;
            const uint64_t t = tile_begin + v;
            return t < tile_end ? t : ~0ull;
        };
        for (;;) {
            bool progressed = false;
            long long c0 = clock64();
            // ---- (1) claim a tile and start its bulk copy (at most two tiles ahead of the scanner)
            if (!no_more && seq_claim < seq_done + NS && seq_claim < seq_scan + 2) {
                const int b = seq_claim % NS;
                if (S.remaining[b] == 0) {
                    const uint64_t t = claim_tile();
                    if (t == ~0ull) no_more = true;
                    else {
                        __threadfence_block();
                        Stage& G = S.st[b];
                        const uint64_t tile_start = t * (uint64_t)TB;
                        const uint32_t avail = (uint32_t)min((uint64_t)TB, P.n - tile_start);
                        const uint32_t halo = t > 0 ? HALO : 0, bulk = avail & ~15u;
                        fence_proxy_async();
                        if (lane == 0) {
                            G.t = t; G.avail = avail;
                            if (halo + bulk) {
                                mbar_expect_tx(&S.tma_bar[b], halo + bulk);
                                bulk_g2s(G.halo + (HALO - halo), P.bytes + tile_start - halo, halo + bulk, &S.tma_bar[b]);
                            }
                        }
                        if (avail < TB) for (uint32_t i = bulk + lane; i < TB; i += 32) G.tile[i] = i < avail ? P.bytes[tile_start + i] : 0;
                        if (t == 0) for (int i = lane; i < HALO; i += 32) G.halo[i] = 0;
                        __syncwarp();
                        seq_claim++;
                    }
                    progressed = true;
                }
            }
            { const long long c1 = clock64(); c_claim += c1 - c0; c0 = c1; }
            // ---- (2) scan the oldest loaded tile and publish its aggregate
            if (seq_scan < seq_claim) {
                const int b = seq_scan % NS;
                Stage& G = S.st[b];
                const uint64_t t = G.t;
                const uint32_t avail = G.avail;
                const uint32_t halo = t > 0 ? HALO : 0;
                const bool has_tma = (halo + (avail & ~15u)) != 0;
                if (!has_tma || mbar_test(&S.tma_bar[b], par[b])) {
                    if (has_tma) par[b] ^= 1;
                    __syncwarp();
                    const uint64_t tile_start = t * (uint64_t)TB;
                    // newline scan: 256 B rows, row = round * 32 + lane, rotated word order
                    const uint32_t nrows = (avail + ROWB - 1) / ROWB;
                    uint32_t cnt[ROUNDS]; uint64_t wmask[ROUNDS];
#pragma unroll
                    for (int rd = 0; rd < ROUNDS; rd++) {
                        cnt[rd] = 0; wmask[rd] = 0;
                        const uint32_t rowi = rd * 32 + lane;
                        if (rowi < nrows) {
                            const uint32_t* row = reinterpret_cast<const uint32_t*>(G.tile) + rowi * ROWW;
#pragma unroll 8
                            for (int j = 0; j < ROWW; j++) {
                                const uint32_t jj = (j + lane) & (ROWW - 1);
                                const uint32_t x = row[jj] ^ 0x0A0A0A0Au;
                                if ((x - 0x01010101u) & ~x & 0x80808080u) {
                                    wmask[rd] |= 1ull << jj;
                                    uint32_t z = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
                                    z = ~(z | x | 0x7F7F7F7Fu);
                                    cnt[rd] += __popc(z);
                                }
                            }
                        }
                    }
                    uint32_t C = 0, off[ROUNDS];
#pragma unroll
                    for (int rd = 0; rd < ROUNDS; rd++) {
                        uint32_t inc = cnt[rd];
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += v; }
                        off[rd] = C + inc - cnt[rd];
                        C += __shfl_sync(0xffffffffu, inc, 31);
                    }
                    const bool overflow = C > NLW;
                    if (!overflow) {
#pragma unroll
                        for (int rd = 0; rd < ROUNDS; rd++) {
                            if (!cnt[rd]) continue;
                            const uint32_t rowi = rd * 32 + lane;
                            const uint32_t* row = reinterpret_cast<const uint32_t*>(G.tile) + rowi * ROWW;
                            uint32_t o = off[rd];
                            uint64_t mk = wmask[rd];
                            while (mk) {
                                const int jj = __ffsll((long long)mk) - 1;
                                mk &= mk - 1;
                                const uint32_t wv = row[jj];
#pragma unroll
                                for (int bsel = 0; bsel < 4; bsel++)
                                    if (((wv >> (8 * bsel)) & 0xFF) == '\n') G.nl[o++] = (uint16_t)(rowi * ROWB + jj * 4 + bsel);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        G.n_nl = overflow ? 0 : C; G.n_nl_raw = C;
                        if (t > 0) {                               // publish the aggregate (tile 0 goes straight to inclusive)
                            SState agg = identity_state();
                            agg.count = C;
                            if (!overflow) for (uint32_t j = 0; j < 4 && j < C; j++) agg.last[j] = tile_start + G.nl[C - 1 - j];
                            TileSlot* slot = &P.slots[t];
                            slot->agg = agg;
                            __threadfence();
                            st_release_u32(&slot->flag, epoch * 4 + 1);
                        }
                    }
                    __syncwarp();
                    seq_scan++;
                    progressed = true;
                }
            }
            { const long long c1 = clock64(); c_scan += c1 - c0; c0 = c1; }
            // ---- (3) resolve the oldest scanned tile: look-back (non-blocking), events, handover to the walkers
            bool resolved = false;
            if (seq_done < seq_scan) {
                const int b = seq_done % NS;
                Stage& G = S.st[b];
                const uint64_t t = G.t;
                SState pre = identity_state();
                if (t == 0 || warp_lookback_try<false>(P, t, epoch, lane, pre)) {
                    const uint8_t* sb = G.tile;
                    const uint32_t avail = G.avail, Cs = G.n_nl, C = G.n_nl_raw;
                    const bool overflow = C > NLW;
                    if (overflow) my_flags |= FLAG_WS_BAIL;
                    const uint32_t halo = t > 0 ? HALO : 0;
                    const uint64_t tile_start = t * (uint64_t)TB;
                    if (lane == 0) {
                        SState agg = identity_state();
                        agg.count = C;
                        if (!overflow) for (uint32_t j = 0; j < 4 && j < C; j++) agg.last[j] = tile_start + G.nl[C - 1 - j];
                        const SState inc = combine(pre, agg);
                        TileSlot* slot = &P.slots[t];
                        slot->inc = inc;
                        __threadfence();
                        st_release_u32(&slot->flag, epoch * 4 + 2);
                        if (t + 1 == P.num_tiles) *P.final_state = inc;
                    }
                    // per-line events (validate, n_bases, n_records) — fastq.rs:240-285
                    const bool line0_starts_here = (t == 0) || (sb[-1] == '\n');
                    const uint32_t ord0 = (uint32_t)(pre.count & 3);
                    auto line_start_rel = [&](uint32_t i) -> int { return i ? (int)G.nl[i - 1] + 1 : 0; };
                    auto prev_nl = [&](uint32_t i, uint32_t back) -> uint64_t {
                        if (i >= back) return tile_start + G.nl[i - back];
                        const uint32_t r = back - i - 1;
                        return r < 4 ? pre.last[r] : NONE;
                    };
                    auto byte_g = [&](uint64_t gpos) -> uint8_t {
                        if (gpos + halo >= tile_start && gpos < tile_start + TB) return sb[(int64_t)gpos - (int64_t)tile_start];
                        return gpos < P.n ? P.bytes[gpos] : 0;
                    };
                    auto cr_before = [&](uint64_t q, uint64_t prevq) -> uint32_t {
                        const uint64_t ls = prevq == NONE ? 0 : prevq + 1;
                        return (q > ls && byte_g(q - 1) == '\r') ? 1u : 0u;
                    };
                    for (uint32_t i = lane; i <= Cs; i += 32) {
                        const uint32_t role = (ord0 + i) & 3;
                        const int s = line_start_rel(i);
                        const bool starts = (i > 0 || line0_starts_here) && (uint32_t)s < avail;
                        if (starts) {
                            if (role == 0 && sb[s] != '@') my_flags |= FLAG_PARSE_ERROR;
                            if (role == 2 && sb[s] != '+') my_flags |= FLAG_PARSE_ERROR;
                        }
                        if (role == 1) {
                            const int e = i < Cs ? (int)G.nl[i] : (int)avail;
                            if (e - s > SEG) my_flags |= FLAG_WS_BAIL;             // long reads: the general kernel cuts pieces
                        }
                        if (i < Cs) {
                            const uint64_t q = tile_start + G.nl[i];
                            if (role == 1) {
                                const uint64_t p1 = prev_nl(i, 1);
                                const uint64_t ls = p1 == NONE ? 0 : p1 + 1;
                                acc_bases += (q - ls) - cr_before(q, p1);
                            } else if (role == 3) {
                                const uint64_t q2 = prev_nl(i, 1), q1 = prev_nl(i, 2), q0 = prev_nl(i, 3);
                                if (q2 == NONE || q1 == NONE || q0 == NONE) my_flags |= FLAG_PARSE_ERROR;
                                else {
                                    if ((q1 - q0 - 1) - cr_before(q1, q0) != (q - q2 - 1) - cr_before(q, q2)) my_flags |= FLAG_PARSE_ERROR;
                                    acc_records++;
                                }
                            }
                        }
                    }
                    // stage metadata: the sequence lines of this tile are items cum .. cum + n_items - 1
                    const uint32_t i_first = (1u - ord0) & 3u;
                    uint32_t n_items = (Cs >= i_first) ? (Cs - i_first) / 4 + 1 : 0;
                    if (n_items) {                                           // an empty last fragment is not an item
                        const uint32_t il = i_first + 4 * (n_items - 1);
                        const int a = line_start_rel(il), e = il < Cs ? (int)G.nl[il] : (int)avail;
                        if (e <= a) n_items--;
                    }
                    const bool last_tile = (t + 1 == P.num_tiles);
                    if (n_items < 32 && !last_tile) my_flags |= FLAG_WS_BAIL;    // batches may span at most two stages
                    my_flags = __reduce_or_sync(0xffffffffu, my_flags);
                    if (my_flags & FLAG_WS_BAIL) n_items = 0;                    // results are discarded: let the walkers drain
                    if (lane == 0) {
                        G.i_first = i_first; G.n_items = n_items;
                        G.line0_starts_here = line0_starts_here ? 1 : 0;
                        if (line0_starts_here) { G.line0_lo = 0; G.line0_exact = 1; }
                        else {
                            const uint64_t p1 = pre.last[0];
                            const int64_t ls = (p1 == NONE ? 0 : (int64_t)p1 + 1) - (int64_t)tile_start;
                            G.line0_exact = ls >= -(int64_t)halo ? 1 : 0;
                            G.line0_lo = G.line0_exact ? (int)ls : -(int)halo;
                        }
                        cum += n_items;
                        S.cum_end[seq_done % CUMN] = cum;
                        S.remaining[b] = n_items;
                        if (my_flags) { atomicOr(&S.flags, my_flags); atomicOr(P.flags, my_flags); }
                        __threadfence_block();
                        S.ready_seq = seq_done + 1;
                    }
                    cum = __shfl_sync(0xffffffffu, cum, 0);
                    seq_done++;
                    progressed = true; resolved = true;
                }
            }
            { const long long c1 = clock64(); if (resolved) c_resolve += c1 - c0; else c_lbfail += c1 - c0; c0 = c1; }
            if (no_more && seq_done == seq_claim) break;
            if (!progressed) { __nanosleep(40); c_idle += clock64() - c0; }
        }
        // end marker
        if (lane == 0) { __threadfence_block(); S.end_seq = seq_done; }
        // producer tallies
#pragma unroll
        for (int d = 16; d; d >>= 1) { acc_records += __shfl_xor_sync(0xffffffffu, acc_records, d); acc_bases += __shfl_xor_sync(0xffffffffu, acc_bases, d); }
        if (lane == 0) {
            atomicAdd((unsigned long long*)&S.tallies[0], (unsigned long long)acc_records); atomicAdd((unsigned long long*)&S.tallies[1], (unsigned long long)acc_bases);
            atomicAdd(&P.tallies[9], (unsigned long long)c_claim); atomicAdd(&P.tallies[10], (unsigned long long)c_scan);
            atomicAdd(&P.tallies[11], (unsigned long long)c_lbfail); atomicAdd(&P.tallies[12], (unsigned long long)c_resolve);
            atomicAdd(&P.tallies[13], (unsigned long long)c_idle);
        }
    } else {
        // ================================================================== walkers
        const FastLuts L{S.lut, S.rins, nullptr};
        uint32_t q = 0;                       // stage sequence number holding the first item of my next batch (monotone)
        uint32_t walker_flags = 0;
        long long c_wait = 0, c_work = 0;
        for (;;) {
            const long long w0 = clock64();
            uint32_t g0 = 0;
            if (lane == 0) g0 = atomicAdd(&S.claim, 32u);
            g0 = __shfl_sync(0xffffffffu, g0, 0);
            // advance q to the stage that contains item g0
            bool ended = false;
            for (;;) {
                while (S.ready_seq <= q && S.end_seq > q) __nanosleep(64);
                if (S.ready_seq <= q) { ended = true; break; }           // end marker reached: no such item
                __threadfence_block();
                if (g0 < S.cum_end[q % CUMN]) break;
                q++;
            }
            if (ended) break;
            // lanes: item g = g0 + lane lives in stage q, q+1 or q+2 (every non-final stage has >= 32 items)
            const uint32_t g = g0 + lane;
            uint32_t my_q = 0xFFFFFFFFu, my_j = 0;
            uint32_t q_hi = q;
            for (uint32_t qq = q; qq < q + NS; qq++) {
                while (S.ready_seq <= qq && S.end_seq > qq) __nanosleep(64);
                if (S.ready_seq <= qq) break;                            // stream ended before qq
                __threadfence_block();
                const uint32_t ce = S.cum_end[qq % CUMN], cb = qq ? S.cum_end[(qq - 1) % CUMN] : 0;
                q_hi = qq;
                if (g >= cb && g < ce && my_q == 0xFFFFFFFFu) { my_q = qq; my_j = g - cb; }
                if (g0 + 31 < ce) break;
            }
            const long long w1 = clock64();
            c_wait += w1 - w0;
            ItemSums r{0, 0, 0, 0};
            if (my_q != 0xFFFFFFFFu) {
                const Stage& G = S.st[my_q % NS];
                const uint8_t* sb = G.tile;
                const uint32_t i = G.i_first + 4u * my_j;
                int a = i ? (int)G.nl[i - 1] + 1 : 0;
                int b = i < G.n_nl ? (int)G.nl[i] : (int)G.avail;
                if (b > a && sb[b - 1] == '\r') b--;
                if (b > a) {
                    int lo = a; bool lo_exact = true;
                    if (i == 0 && !G.line0_starts_here) { lo = G.line0_lo; lo_exact = G.line0_exact != 0; }
                    uint32_t slow = 0;
                    const int ws = find_ws(sb, S.lut, a, lo, lo_exact, K, slow);
                    if (slow || !walk_item<K, M>(sb, L, ws, b, r)) walker_flags |= FLAG_WS_BAIL;
                }
            }
            // reduce the batch and release the stages it touched
            uint64_t v[4] = {r.nk, r.nrc, r.ksum, r.msum};
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++)
#pragma unroll
                for (int d = 16; d; d >>= 1) v[k4] += __shfl_xor_sync(0xffffffffu, v[k4], d);
            if (lane == 0) {
                atomicAdd((unsigned long long*)&S.tallies[2], (unsigned long long)v[0]);
                atomicAdd((unsigned long long*)&S.tallies[3], (unsigned long long)v[1]);
                atomicAdd((unsigned long long*)&S.tallies[4], (unsigned long long)v[2]);
                if (M > 0) { atomicAdd((unsigned long long*)&S.tallies[7], (unsigned long long)v[0]); atomicAdd((unsigned long long*)&S.tallies[8], (unsigned long long)v[3]); }
            }
            __threadfence_block();
            for (uint32_t qq = q; qq <= q_hi; qq++) {
                const uint32_t c = __popc(__ballot_sync(0xffffffffu, my_q == qq));
                if (c && lane == 0) atomicSub((uint32_t*)&S.remaining[qq % NS], c);
            }
            c_work += clock64() - w1;
        }
        if (lane == 0) { atomicAdd(&P.tallies[14], (unsigned long long)c_wait); atomicAdd(&P.tallies[15], (unsigned long long)c_work); }
        walker_flags = __reduce_or_sync(0xffffffffu, walker_flags);
        if (lane == 0 && walker_flags) { atomicOr(&S.flags, walker_flags); atomicOr(P.flags, walker_flags); }
    }
    __syncthreads();
    if (tid < 9 && S.tallies[tid]) atomicAdd(&P.tallies[tid], (unsigned long long)S.tallies[tid]);
}
}  // namespace fused_ws
