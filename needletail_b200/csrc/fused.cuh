// fused.cuh — the hot path: ONE pass over FASTX bytes -> tallies.
//
//   delimiter scan (fastq.rs:155-187,306-308 / fasta.rs:220-243)  +  validate (fastq.rs:240-285)
//   + normalize (sequence.rs:19-62) + reverse_complement (sequence.rs:67-105,202-208)
//   + CanonicalKmers (kmer.rs:84-129) + BitNuclKmer/minimizer (bitkmer.rs:26-162)
//
// Persistent CTAs (9 walker warps + 1 coordinator warp, 2 per SM) take tiles of <= 84 KiB round-robin.  Per tile:
//   P0  the tile (+128 B back halo) is brought into shared memory by one cp.async.bulk (TMA 1-D
//       bulk copy, SASS UBLKCP) completing on an mbarrier; the coordinator prefetches the next tile into L2;
//   P1  every walker thread scans 256 B rows for '\n' (scan_row: 16 x LDS.128 per row in a rotated, bank-conflict
//       free order, branch-free SIMD-in-register byte test);
//   P2  block scan -> sorted newline list; the coordinator warp publishes the tile's aggregate (newline
//       count, last four newline positions, FASTA header state) and obtains the global prefix by
//       decoupled look-back, 32 predecessors per step (single pass: the input is read from HBM once).  For
//       speculative FASTQ tiles the look-back is deferred by one tile (resolve_pending);
//   P3  line roles (FASTQ: newline ordinal mod 4 — walkers start on a locally inferred phase that the
//       coordinator verifies; FASTA: '>' at line start) -> validation events, n_records / n_bases, and one
//       sequential "walker" per sequence-line fragment: 2-bit rolling forward / reverse-complement words,
//       canonical select, sliding-window (van Herk) minimizer.  walk_clean takes all-ACGT items, walk_fast items
//       with other bases, walk anything else;
//   P4  per-thread tallies stay in registers across tiles; one block reduction + 9 atomics per CTA.
// Anything the fast path cannot prove clean (parse error, > NLMAX newlines in a tile, whitespace
// runs longer than the halo) raises a flag and the host re-runs the exact materialising path.
// Part of the unity build (ntgpu.cu).
#pragma once
#include <type_traits>
#include "common.cuh"

namespace fused {
// CTA shape (compile-time; tools/ab_shapes.sh measures the alternatives): walker threads, CTAs per SM, tile capacity, newline capacity
#ifndef NTG_NTW
#define NTG_NTW 288
#endif
#ifndef NTG_CTAS
#define NTG_CTAS 2
#endif
#ifndef NTG_TILE_KB
#define NTG_TILE_KB 84
#endif
#ifndef NTG_NLMAX
#define NTG_NLMAX 2048
#endif
constexpr int NTW = NTG_NTW;             // walker warps x 32 ...
constexpr int NT = NTW + 32;             // ... + 1 coordinator warp (highest warp id); NTG_CTAS CTAs per SM, 96 registers per thread
constexpr int CTAS_PER_SM = NTG_CTAS;
constexpr int ROWB = 256;                // P1 scans the tile in 256 B rows, one row per thread per round
constexpr int ROWW = ROWB / 4;
constexpr int TILE = NTG_TILE_KB * 1024;          // capacity of the shared-memory tile; the tile size in use is Params::tile_bytes
constexpr int MAXROUNDS = (TILE / ROWB + NTW - 1) / NTW;   // 2  (rows are scanned by the walker threads)
constexpr int HALO = 128;                // back halo (>= k-1 bases for k <= 64, plus slack)
constexpr int NLMAX = NTG_NLMAX;              // newline capacity per tile (mean line >= 42 B at full tile size)
constexpr int SEG = 512;                 // long lines are cut into SEG-byte pieces
constexpr int LONGMAX = TILE / SEG + 2;
constexpr int NWK = NT;                  // threads that run the tile loop's cooperative phases
constexpr uint64_t NONE = ~0ull;
constexpr uint64_t INHDR = ~0ull - 1;

enum : uint32_t { FLAG_PARSE_ERROR = 1, FLAG_NL_OVERFLOW = 2, FLAG_HALO_OVERFLOW = 4, FLAG_WS_BAIL = 8, FLAG_SPEC_MISS = 16, FLAG_FORMAT = 32 };

// carried scan state (prefix over tiles)
struct SState {
    uint64_t count;        // newlines so far
    uint64_t last[4];      // global positions of the most recent newlines, last[0] newest; NONE if absent
    uint64_t n_starts;     // FASTA record starts
    uint64_t hdr;          // FASTA: NONE = no record start in this span; INHDR = span ends inside a header
                           //        line; else position of the newline that ended the latest header
    uint64_t first_nl;     // FASTA: first newline of the span (NONE if none)
};
struct __align__(128) TileSlot {
    SState agg;
    SState inc;
    uint32_t flag;         // 0 = empty, 1 = aggregate published, 2 = inclusive prefix published
    uint32_t pad[15];
};

__host__ __device__ inline SState identity_state() {
    SState s; s.count = 0; s.n_starts = 0; s.hdr = NONE; s.first_nl = NONE;
    for (int i = 0; i < 4; i++) s.last[i] = NONE;
    return s;
}
// a = earlier span, b = later span
__host__ __device__ inline SState combine(const SState& a, const SState& b) {
    SState r;
    r.count = a.count + b.count;
    int j = 0;
    for (int i = 0; i < 4 && j < 4; i++) if (b.last[i] != NONE) r.last[j++] = b.last[i];
    for (int i = 0; i < 4 && j < 4; i++) if (a.last[i] != NONE) r.last[j++] = a.last[i];
    for (; j < 4; j++) r.last[j] = NONE;
    r.n_starts = a.n_starts + b.n_starts;
    r.first_nl = a.first_nl != NONE ? a.first_nl : b.first_nl;
    if (b.hdr != NONE) r.hdr = b.hdr;
    else if (a.hdr == INHDR) r.hdr = (b.first_nl != NONE) ? b.first_nl : INHDR;
    else r.hdr = a.hdr;
    return r;
}

struct Params {
    const uint8_t* bytes;
    uint64_t n;
    uint64_t num_tiles;
    TileSlot* slots;                   // ring of 1 << slot_shift tile slots (tile t uses slot t & slot_mask, generation t >> slot_shift)
    unsigned long long* cw;            // FASTQ: one 64-bit look-back word per tile, same ring (see fastq_lookback)
    uint32_t slot_mask, slot_shift;
    uint64_t gmin;                     // lowest global byte position that is addressable through `bytes` in this launch (streamed rings)
    uint32_t* ticket;
    unsigned long long* tallies;      // 16 x u64
    uint32_t* flags;
    unsigned long long* err_key;       // first parse error of the stream: min over (start byte of the failing record << 2 | check order); ~0 = none
    unsigned long long* fin;           // k_finalize: [0] error kind of an end-of-stream error, [1] its line (FASTA)
    unsigned long long* reduce_buf;    // non-null: k_finalize copies the tallies (+ a "needs the host" count) into the all-reduce send buffer
    SState* final_state;
    unsigned long long* fa_totals;     // FASTA: [0] record starts, [1] newlines of the whole pass (sums; the look-back carries the header state only)
    // k-mer spectrum passes (spectrum.cuh): every canonical k-mer the generic walker tallies is also counted, in a dense
    // histogram (k <= 14: 4^k counters) or in an open-addressing hash table (keys ~0 = empty)
    uint32_t* sp_dense;
    unsigned long long* sp_keys;
    uint32_t* sp_counts;
    uint64_t sp_mask;
    uint32_t* sp_overflow;
    uint32_t k, m, w;
    uint32_t qmask;                    // != 0: bases whose quality byte is below it count as 'N' (record-owned FASTQ kernel only)
    uint32_t tile_bytes;               // multiple of 256, <= TILE: sized so one tile holds about NT sequence lines
    int format;                        // NTG_FMT_FASTA / NTG_FMT_FASTQ
    int has_query;
    uint32_t one;                      // == 1 (run-time constant for mad.wide)
    uint32_t spec;                     // FASTQ: walkers start on a locally inferred line phase while the coordinator warp
                                       // does the look-back; a wrong guess raises FLAG_SPEC_MISS and the host re-runs with spec = 0
    uint64_t q_lo, q_hi;
};

struct Acc {
    uint64_t n_records = 0, n_bases = 0, n_kmers = 0, n_not_rc = 0, ksum_lo = 0, ksum_hi = 0, n_query = 0, n_mini = 0, msum = 0;
};

// ---- PTX helpers: mbarrier + 1-D bulk async copy (TMA) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// First parse error: the host replays the stream truncated at the start of the failing record (records before it are
// delivered, fastq.rs:243,253,277) and classifies the error there with the record scanner (parse.cuh).  `check` = order of
// the reference's checks inside one record (0 start byte, 1 separator, 2 lengths, 3 end of stream).
__host__ __device__ __forceinline__ void note_parse_error(unsigned long long* err_key, uint32_t& slow, uint64_t rec_start, uint32_t check) {
    slow |= 1u;                                       // FLAG_PARSE_ERROR
    const unsigned long long key = ((unsigned long long)rec_start << 2) | check;
#ifdef __CUDA_ARCH__
    atomicMin(err_key, key);
#else
    if (key < *err_key) *err_key = key;
#endif
}

// ---- shared memory layout ------------------------------------------------------------------
struct __align__(16) Smem {
    uint8_t halo[HALO];
    uint8_t tile[TILE];
    uint32_t nl[NLMAX + 8];            // sorted tile-relative newline offsets
    uint32_t rstart[NLMAX + 8];        // FASTA: per line, tile-relative (+HALO) start of its sequence region
    uint8_t lut[256];                  // 0..3 ACGT, 4 kept non-ACGT, 0x85 deleted (space/tab), 0x86 deleted (\r \n)
    uint32_t rins[256];                // fast walker: complement base pre-shifted into the high word of R
    uint32_t comb[256];                // clean walker: lut | rins in one word
    uint64_t bar;
    uint32_t warp_tmp[4][NT / 32 + 2];  // one array per block-scan call site (the scans use a single barrier)
    uint64_t red[NT / 32][9];
    SState prefix;                     // exclusive prefix of this tile
    volatile uint32_t prefix_seq;      // number of tiles of this CTA whose prefix has been resolved by the coordinator
    // the tile whose look-back the coordinator has deferred by one tile (speculative FASTQ, see P2c); coordinator warp only
    SState pend_agg;
    uint64_t pend_t;
    uint32_t pend_guess, pend_cs, pend_avail, pend_line0;
    uint32_t pend_nl4[4];              // the tile's first four newline offsets
    volatile uint32_t pend_valid;
    uint32_t tile_idx_next;            // ticket of this CTA's next tile, claimed by walker thread 0 at the end of its walk
    uint32_t n_long;
    uint32_t arrived;                  // warps that have finished this tile's walk (the last one takes the CTA's next ticket)
    uint32_t long_line[LONGMAX];       // line indices of long sequence lines
    uint32_t long_pref[LONGMAX + 1];   // exclusive prefix of piece counts
    int32_t bcast[4];
};

// end of a tile's walk: the last warp of the CTA to get here takes the CTA's next ticket (read by all after the next barrier)
__device__ __forceinline__ void claim_when_last(Smem& S, uint32_t* ticket, uint32_t lane) {
    __syncwarp();
    if (lane == 0 && atomicAdd(&S.arrived, 1u) == NT / 32 - 1) { S.arrived = 0; S.tile_idx_next = atomicAdd(ticket, 1u); }
}
// barrier of the threads that run the tile loop (all of the CTA, or the walkers only when the coordinator is decoupled)
__device__ __forceinline__ void tile_sync() { __syncthreads(); }
// Block-wide scans with ONE barrier (round 2: the serial fold by thread 0 between two more barriers held 4.7 % of the stall
// samples): warp scans by shuffle, the warp totals go through shared memory, every warp folds the totals of the warps
// before it itself.  `tmp` must not be written again before another CTA barrier: every call site has its own array.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* tmp) {
    constexpr int NW = NWK / 32;
    static_assert(NW <= 16, "the warp totals are folded by one 16-lane shuffle scan");
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) tmp[w] = inc;
    tile_sync();
    const uint32_t x = lane < NW ? tmp[lane] : 0u;
    uint32_t xi = x;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi += t; }
    *total = __shfl_sync(0xffffffffu, xi, NW - 1);
    return inc - v + __shfl_sync(0xffffffffu, xi - x, w);
}
__device__ __forceinline__ uint32_t block_incl_max(uint32_t v, uint32_t* tmp) {   // inclusive max-scan across threads
    constexpr int NW = NWK / 32;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc = max(inc, t); }
    if (lane == 31) tmp[w] = inc;
    tile_sync();
    const uint32_t x = lane < NW ? tmp[lane] : 0u;
    uint32_t xi = x;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi = max(xi, t); }
    const uint32_t before = __shfl_sync(0xffffffffu, xi, w ? w - 1 : 0);           // max over the warps before w
    return w ? max(inc, before) : inc;
}

// ---- P1 helper: newlines of one 256 B row ---------------------------------------------------
// cnt = number of '\n' bytes, mask bit w = word w of the row holds at least one.  Sixteen 128-bit loads; at step j a
// lane reads 16 B column (j + lane) & 15, so the eight lanes of a quarter-warp (rows are 256 B apart) hit eight
// different bank groups.  Branch-free: per word an exact SIMD-in-register byte test, a byte-lane counter and one
// predicated OR with a compile-time bit; the lane's rotation is undone once at the end.
__host__ __device__ __forceinline__ void scan_row(const uint32_t* __restrict__ row, uint32_t lane, uint32_t& cnt, uint64_t& mask) {
    static_assert(ROWB == 256, "scan_row: 16 x 16 B");
    const uint4* __restrict__ row4 = reinterpret_cast<const uint4*>(row);
    uint32_t acc = 0, mlo = 0, mhi = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint4 v = row4[(j + lane) & 15u];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t x = w[q] ^ 0x0A0A0A0Au;
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;    // 0x80 in every byte that is exactly '\n'
            acc += z >> 7;                                                              // <= 64 per byte lane
            const int bit = j * 4 + q;
            if (z) { if (bit < 32) mlo |= 1u << (bit & 31); else mhi |= 1u << (bit & 31); }
        }
    }
    const uint32_t t = (acc & 0x00FF00FFu) + ((acc >> 8) & 0x00FF00FFu);
    cnt = (t & 0xFFFFu) + (t >> 16);
    const uint64_t m = ((uint64_t)mhi << 32) | mlo;
    const uint32_t rot = (4u * lane) & 63u;                                            // step j saw column (j + lane) & 15
    mask = rot ? ((m << rot) | (m >> (64u - rot))) : m;
}

// ---- k-mer spectrum: count one canonical k-mer (k <= 32) ----------------------------------------------------
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t h) {          // MurmurHash3 finaliser
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
    return h;
}
constexpr uint32_t SPECTRUM_MAX_PROBES = 8192;
__device__ __forceinline__ void spectrum_add(unsigned long long* keys, uint32_t* counts, uint64_t mask, uint32_t* overflow, uint64_t key, uint32_t inc) {
    uint64_t h = fmix64(key) & mask;
    for (uint32_t probe = 0; probe < SPECTRUM_MAX_PROBES; probe++) {
        unsigned long long old = keys[h];
        if (old == ~0ull) old = atomicCAS(&keys[h], ~0ull, (unsigned long long)key);
        if (old == ~0ull || old == key) { atomicAdd(&counts[h], inc); return; }
        h = (h + 1) & mask;
    }
    atomicAdd(overflow, 1u);                                               // table full: the host reports NTG_ENOMEM
}

// =============================================================================== the walker
// Processes the sequence bytes sb[a..b) (tile-relative; sb[-HALO..-1] is the back halo) of one
// sequence-line fragment.  `lo` = lowest index the warm-up may read; lo_exact tells whether lo is the
// true start of the sequence region (nothing before it belongs to this sequence).
// K-mers are owned by their LAST base: every k-mer whose last base lies in [a,b) is tallied here.
//   KW   : 1 -> k <= 32 (u64 words), 2 -> k <= 64 (2 x u64)
//   MINI : also bit_kmers(k,false) -> bitkmer::minimizer(m)      (KW == 1 only)
//   W    : compile-time window k-m+1 (0 = run-time window, arrays indexed dynamically)
// warm-up: step back from a over at most k-1 kept good bases, not below lo; returns where to start walking
__host__ __device__ __forceinline__ int find_ws(const uint8_t* __restrict__ sb, const uint8_t* __restrict__ lut, int a, int lo, bool lo_exact,
                                       int k, uint32_t& slow) {
    int ws = a, got = 0, p = a - 1;
    bool stopped = false;
    while (p >= lo && got < k - 1) {
        const uint8_t c = lut[sb[p]];
        if (c <= 3) { got++; ws = p; }
        else if (c == 4) { stopped = true; break; }         // a non-ACGT base resets everything before it
        p--;
    }
    if (!stopped && got < k - 1 && p < lo && !lo_exact) slow |= FLAG_HALO_OVERFLOW;
    return ws;
}

// find_ws that also returns the warm-up bases themselves, packed 2 bits each with the base next to `a` in bits 0..1, and their
// number.  For walk_clean_w: a warm-up that crosses deleted bytes (wrapped FASTA: the newline before every line) is
// handed over as codes, so that the item proper stays free of deleted bytes.
__host__ __device__ __forceinline__ int find_ws_codes(const uint8_t* __restrict__ sb, const uint8_t* __restrict__ lut, int a, int lo, bool lo_exact,
                                             int k, uint32_t& slow, int& got_out, uint64_t& codes_out) {
    int ws = a, got = 0, p = a - 1;
    uint64_t codes = 0;
    bool stopped = false;
    while (p >= lo && got < k - 1) {
        const uint8_t c = lut[sb[p]];
        if (c <= 3) { codes |= (uint64_t)c << (2 * got); got++; ws = p; }
        else if (c == 4) { stopped = true; break; }         // a non-ACGT base resets everything before it
        p--;
    }
    if (!stopped && got < k - 1 && p < lo && !lo_exact) slow |= FLAG_HALO_OVERFLOW;
    got_out = got; codes_out = codes;
    return ws;
}

template <int KW, bool MINI, int W>
__host__ __device__ __forceinline__ void walk(const uint8_t* __restrict__ sb, const uint8_t* __restrict__ lut, int ws, int a, int b,
                                     const Params& P, Acc& acc, bool count_bases) {
    const int k = (int)P.k;
    uint64_t f0 = 0, f1 = 0, r0 = 0, r1 = 0;              // forward / reverse-complement words (x1 = high word, KW == 2)
    int run = 0;
    const uint64_t kmask = (KW == 1) ? mask2k(P.k) : mask2k(P.k - 32);        // mask of the top word
    const int rshift = (KW == 1) ? 2 * (k - 1) : 2 * (k - 33);               // where a new rc base enters the top word
    // minimizer: sliding-window minimum by van Herk block decomposition over the stream of kept bases
    const uint64_t mmask = MINI ? mask2k(P.m) : 0, lmask = MINI ? mask2k(P.k - P.m) : 0;
    constexpr int WA = MINI ? ((W > 0) ? W : 32) : 1;
    uint64_t cur[WA + 1], suf[WA + 1], pre = 0;
#pragma unroll
    for (int i = 0; i <= WA; i++) { cur[i] = 0; suf[i] = 0; }
    const int w = MINI ? ((W > 0) ? W : (int)P.w) : 1;

    int p = ws;
    for (;;) {
#pragma unroll
        for (int i = 0; i < ((W > 0) ? W : w); i++) {
            uint8_t c;
            do {                                             // fetch the next kept base
                if (p >= b) return;
                c = lut[sb[p]];
                if (count_bases && p >= a && c != 0x86) acc.n_bases++;   // FASTA num_bases: everything but \r \n
                p++;
            } while (c >= 0x80);
            const uint64_t code = c & 3;
            run = (c <= 3) ? run + 1 : 0;                    // a bad base (c == 4) still occupies a slot
            if (KW == 1) {
                f0 = ((f0 << 2) | code) & kmask;
                r0 = (r0 >> 2) | ((3 - code) << rshift);
            } else {
                f1 = ((f1 << 2) | (f0 >> 62)) & kmask;
                f0 = (f0 << 2) | code;
                r0 = (r0 >> 2) | (r1 << 62);
                r1 = (r1 >> 2) | ((3 - code) << rshift);
            }
            uint64_t win = 0;
            if (MINI) {
                // score of the m-mer ending here: min(x, RC_k(x)), x = f & mmask, RC_k(x) = r | lmask  (bitkmer.rs:146-162)
                const uint64_t x = f0 & mmask, y = r0 | lmask;
                const uint64_t s = x < y ? x : y;
                pre = (i == 0) ? s : (s < pre ? s : pre);
                cur[i] = s;
                const uint64_t sv = suf[i + 1];
                win = (i == w - 1) ? pre : (sv < pre ? sv : pre);
            }
            if (p > a && run >= k) {                         // last base index p-1 >= a
                const bool lt = (KW == 1) ? (f0 < r0) : (f1 < r1 || (f1 == r1 && f0 < r0));
                acc.n_kmers++;
                acc.n_not_rc += lt ? 1 : 0;                  // ties => was_rc = true (kmer.rs:124-128)
                const uint64_t c0 = lt ? f0 : r0, c1 = lt ? f1 : r1;
                acc.ksum_lo += c0;
                if (KW == 2) acc.ksum_hi += c1;
                if (P.has_query && c0 == P.q_lo && (KW == 1 || c1 == P.q_hi)) acc.n_query++;
#ifdef __CUDA_ARCH__
                if (KW == 1 && !MINI) {
                    if (P.sp_dense) atomicAdd(&P.sp_dense[c0], 1u);
                    else if (P.sp_keys) spectrum_add(P.sp_keys, P.sp_counts, P.sp_mask, P.sp_overflow, c0, 1u);
                }
#endif
                if (MINI) { acc.n_mini++; acc.msum += win; }
            }
        }
        if (MINI) {                                          // suffix minima of the block just finished
            suf[w - 1] = cur[w - 1];
#pragma unroll
            for (int j = WA - 2; j >= 0; j--) if (j < w - 1) { const uint64_t t = suf[j + 1]; suf[j] = cur[j] < t ? cur[j] : t; }
        }
    }
}

// =============================================================================== the fast walker
// Constant-folded walker for the headline shapes (K, M compile-time, 17 <= K <= 31): the item bytes
// sb[ws..b) must contain no deleted bytes (whitespace); if one shows up the function returns false
// and the caller redoes the item with the generic walker above (acc is untouched in that case).
//  - F / R are kept as explicit 32-bit word pairs (low-aligned, < 2^62);
//  - every 64-bit "a < b" is ONE DSETP on the otherwise idle FP64 pipe: two non-negative integers
//    below 2^62 order exactly like the doubles with the same bit patterns (finite, positive);
//  - checksums are carried per item as a 64-bit sum of the low words and a 32-bit sum of the high words
//    (3 integer adds per value) and folded when the item ends;
//  - non-ACGT bases never branch: they only push `next_ok`, the first index where a k-mer may end.
struct FalseT { static constexpr bool value = false; };
struct TrueT { static constexpr bool value = true; };
struct FastLuts {
    const uint8_t* cls;      // 0..3 code, 4 = kept non-ACGT, >= 0x80 = deleted byte
    const uint32_t* rins;    // ((3 - code) << (2(K-1) - 32)) : the complement base entering the high word of R
    const uint32_t* comb;    // cls | rins in one word (walk_clean: one table look-up per base)
};
__host__ __device__ __forceinline__ double u64_bits_as_double(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)v);
#else
    double d; std::memcpy(&d, &v, sizeof d); return d;
#endif
}
__host__ __device__ __forceinline__ bool lt62(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)a) < __longlong_as_double((long long)b);
#else
    return a < b;                                    // (host build of the walkers: tests/cpp/test_walkers.cu)
#endif
}

template <int K, int M>
__host__ __device__ __forceinline__ bool walk_fast(const uint8_t* __restrict__ sb, const FastLuts& L, int ws, int b, Acc& acc, uint32_t* seen_out = nullptr) {
    static_assert(K >= 17 && K <= 31 && M >= 0 && M <= K, "fast walker shape");
    constexpr bool MINI = M > 0;
    constexpr int W = MINI ? K - M + 1 : 1;
    constexpr uint64_t KMASK = (1ull << (2 * K)) - 1;
    constexpr uint64_t MMASK = MINI ? ((1ull << (2 * M)) - 1) : 0, LMASK = MINI ? ((1ull << (2 * (K - M))) - 1) : 0;
    // (Measured, same box: giving this walker walk_clean's score shortcuts and whole blocks without the per-element bounds test made
    //  the 1 % N shape 10 % faster — 380 -> 417..425 Gbases/s — but cost the clean shapes 1.6 - 5 %: the callee's register needs shift
    //  the register allocation and the placement of the callers' hot loops.  The headline shape wins; kept as it was.)
    using ScoreT = uint64_t;
    auto smin = [](ScoreT a, ScoreT c) -> ScoreT { return lt62(a, c) ? a : c; };
    uint64_t f = 0, r = 0;
    uint32_t seen = 0;
    int next_ok = ws + K - 1;
    // tallies of this item: integer adds measured 11 % faster than exact-double accumulation on the FP64 pipe
    // (the kernel is issue-slot bound: the doubles cost extra register-pair moves)
    uint64_t s_kl = 0, s_ml = 0;                     // sums of the low words (carries kept)
    uint32_t s_kh = 0, s_mh = 0, n_k = 0, n_nrc = 0; // sums of the high words are needed mod 2^32 only
    ScoreT cur[W + 1], suf[W + 1], pre = 0;
#pragma unroll
    for (int i = 0; i <= W; i++) { cur[i] = 0; suf[i] = 0; }
    int p = ws;

    auto roll = [&](int pp) {                       // consume byte pp: class, F, R, next_ok
        const uint32_t byte = sb[pp];
        const uint32_t c = L.cls[byte];
        const uint32_t ri = L.rins[byte];
        seen |= c;
        if (c > 3) next_ok = pp + K;                  // a non-ACGT base: no k-mer may end before pp + K
        f = ((f << 2) | (uint64_t)(c & 3)) & KMASK;
        r = (r >> 2) | ((uint64_t)ri << 32);
    };
    // score of the m-mer ending here: min(x, RC_k(x)); x = F & MMASK, RC_k(x) = R | LMASK   (bitkmer.rs:146-162)
    auto score = [&]() -> ScoreT {
        const uint64_t x = f & MMASK, y = r | LMASK;
        return lt62(x, y) ? x : y;
    };
    auto tally = [&](int pp, ScoreT win) {           // k-mer ending at pp (if allowed) + its minimizer
        const bool emit = pp >= next_ok;
        const bool lt = lt62(f, r);                     // ties => was_rc = true (kmer.rs:124-128)
        const uint64_t c = lt ? f : r;
        if (emit) {
            s_kl += (uint32_t)c; s_kh += (uint32_t)(c >> 32);
            n_k++;
            n_nrc += lt ? 1u : 0u;
            if (MINI) { s_ml += (uint32_t)win; s_mh += (uint32_t)(win >> 32); }
        }
    };
    auto element = [&](int i, int pp) {               // element i of the running van Herk block
        roll(pp);
        ScoreT win = 0;
        if (MINI) {
            const ScoreT sc = score();
            pre = (i == 0) ? sc : smin(sc, pre);
            cur[i] = sc;
            win = (i == W - 1) ? pre : smin(suf[i + 1], pre);
        }
        tally(pp, win);
    };

    // phase 1: the first M-1 bases only feed F / R (no m-mer is complete, no k-mer can end)
    {
        const int e0 = ws + (MINI ? M - 1 : K - 1), e1 = b < e0 ? b : e0;
        for (; p < e1; p++) roll(p);
    }
    // phase 2: van Herk blocks of W m-mer scores
    while (p < b) {
        if (seen & 0x80u) return false;
        const bool full = p + W <= b;
#pragma unroll
        for (int i = 0; i < W; i++) {
            if (!full && p + i >= b) break;
            element(i, p + i);
        }
        p += W;
        if (MINI && full) {
            suf[W - 1] = cur[W - 1];
#pragma unroll
            for (int j = W - 2; j >= 1; j--) suf[j] = smin(cur[j], suf[j + 1]);     // suf[0] is never read
        }
    }
    if (seen & 0x80u) return false;                  // a deleted byte inside the item: not this walker's business
    if (seen_out) *seen_out = seen;
    const uint64_t nk = n_k;
    acc.n_kmers += nk; acc.n_not_rc += n_nrc;
    acc.ksum_lo += s_kl + ((uint64_t)s_kh << 32);
    if (MINI) { acc.n_mini += nk; acc.msum += s_ml + ((uint64_t)s_mh << 32); }
    return true;
}


// (Round 2, measured on B200 with tools/ubench/pipes.cu, profiles/r2a_*: SEL, FSEL, LOP3, SHF, IADD3, PRMT, VIMNMX and ISETP all
// issue on the ALU pipe at 2 cycles per warp instruction, IMAD / IMAD.WIDE / IMAD.MOV on the FMA pipe at 2 cycles, DSETP at
// ~4.5 cycles on the FP64 pipe.  Moving the 64-bit selects of the window minima to the FMA pipe as predicated multiply-adds
// by a run-time 1 was tried (NTG_WV builds, profiles/r2b_*): ptxas folds most of them back into SEL or adds register moves;
// 462 -> 450..462 Gbases/s.  Dropped.)

// =============================================================================== the clean walker
// The common case made cheap: the item bytes sb[ws..b) are all ACGT/acgt.  Anything else (a kept non-ACGT base, a
// deleted byte) only sets a bit in `seen`; the function then returns false with acc untouched and the caller redoes
// the item with walk_fast / walk.  Knowing that every base is good removes the per-base "may a k-mer end here" logic:
// exactly the positions >= ws + K - 1 emit, so the walk is  head (K-1 bases, nothing emitted)  +  unconditional blocks.
//  - one combined table word per base (class bits 0..7, complement base pre-shifted for the high word of R);
//  - F is kept unmasked (old bases fall off the top of the 64-bit word), masked copies feed the compares;
//  - the van Herk buffers share ONE array: buf[i] holds the current block's scores below the running position and the
//    previous block's suffix minima above it (the body is rotated so that the suffix pass follows element W-1);
//  - checksums are plain wrapping 64-bit adds (2 instructions per value), the k-mer count is b - (ws + K - 1).
//  - WARM: the K-1 warm-up bases are handed over as 2-bit codes in `wcodes` (find_ws_codes: code h = bits 2(K-2-h).., h = 0 the
//    farthest) instead of bytes: lines of wrapped FASTA, whose warm-up crosses the previous line break.  `ws` is then the
//    first byte of the item proper and every base of [ws,b) ends a k-mer.  (Round 2, measured: wrapped FASTA 75 -> 117 Gbases/s.)
template <int K, int M, bool WARM>
__host__ __device__ __forceinline__ bool walk_clean(const uint8_t* __restrict__ sb, const uint32_t* __restrict__ comb, int ws, int b, uint64_t wcodes, Acc& acc) {
    static_assert(K >= 21 && K <= 31 && M >= 0 && M <= K, "clean walker shape (class bits 0..7 must not overlap the R insert)");
    constexpr bool MINI = M > 0;
    constexpr int W = MINI ? K - M + 1 : 1;
    constexpr int B = MINI ? W : 8;                  // bases per unrolled block
    constexpr uint64_t KMASK = (1ull << (2 * K)) - 1;
    constexpr uint64_t MMASK = MINI ? ((1ull << (2 * M)) - 1) : 0, LMASK = MINI ? ((1ull << (2 * (K - M))) - 1) : 0;
    constexpr uint32_t RMASK = 3u << (2 * (K - 1) - 32);
    constexpr bool RARE = MINI && (K - M) >= 8 && (K - M) <= 16;                 // see score()
    constexpr uint32_t RTOP_LIMIT = (2 * M >= 32) ? (1u << (MINI ? 2 * M - 32 : 0)) : 1u;   // R < 4^M  <=>  top < RTOP_LIMIT
    // scores of at most 32 bits (RARE shapes with M <= 16, e.g. k=21 m=11): the window minima are one VIMNMX each instead of
    // DSETP + 2 SEL — 30 of the 41 64-bit minima per 11 bases
    constexpr bool S32 = RARE && 2 * M <= 32;
    // (FP64-pipe minima for the 42-bit scores of m = 21 — doubles 2^52 + x, min(a, c) = a - ((a - c) + |a - c|) / 2, three DADD / DFMA and no
    //  INT-pipe instruction — measured 539 vs 551 Gbases/s on the headline shape: dropped)
    using ScoreT = typename std::conditional<S32, uint32_t, uint64_t>::type;
    uint64_t f = 0, r = 0, s_k = 0, s_m = 0;
    ScoreT pre = 0;
    uint32_t seen = 0, n_nrc = 0, rtop_min = 0xFFFFFFFFu;
    ScoreT buf[W + 1];
    auto smin = [](ScoreT a, ScoreT c) -> ScoreT {
        if constexpr (S32) return a < c ? a : c;
        else return lt62(a, c) ? a : c;
    };
#pragma unroll
    for (int i = 0; i <= W; i++) buf[i] = 0;
    int p = ws;

    // Leave as soon as a lane of the (converged part of the) warp has met a byte this walker cannot handle: the warp then
    // redoes its items with walk_fast together instead of finishing a walk whose result is thrown away.  A hint only:
    // exactness rests on each lane's own `seen` test at the end.
    auto bail = [&]() -> bool {
#if defined(__CUDA_ARCH__)
        return __any_sync(__activemask(), (seen & 0x84u) != 0u);
#else
        return (seen & 0x84u) != 0u;
#endif
    };
    auto roll = [&](int pp) {
        const uint32_t u = comb[sb[pp]];
        seen |= u;
        f = (f << 2) | (uint64_t)(u & 3u);
        r = (r >> 2) | ((uint64_t)(u & RMASK) << 32);
    };
    // score of the m-mer x ending here: min(x, RC_k(x)), RC_k(x) = R | LMASK (bitkmer.rs:146-162).  RC_k(x) can only be
    // the smaller one when R < 4^M, i.e. when the last K-M bases are all T (4^-(K-M) per position in random sequence):
    // RARE shapes take x and only remember the smallest top part of R seen; an item where that ever reached zero
    // is handed to walk_fast like one with a non-ACGT base.
    auto score = [&]() -> ScoreT {
        if constexpr (RARE) {
            const uint32_t top = (2 * M >= 32) ? (uint32_t)(r >> 32) : (uint32_t)(r >> (2 * M));
            rtop_min = top < rtop_min ? top : rtop_min;
            if constexpr (S32) return (uint32_t)f & (uint32_t)MMASK;
            else return f & MMASK;
        } else {
            const uint64_t x = f & MMASK, y = r | LMASK;
            return lt62(x, y) ? x : y;
        }
    };
    auto tally = [&](ScoreT win) {
        const uint64_t fm = f & KMASK;
        const bool lt = lt62(fm, r);                 // ties => was_rc = true (kmer.rs:124-128)
        s_k += lt ? fm : r;
        n_nrc += lt ? 1u : 0u;
        if (MINI) s_m += (uint64_t)win;
    };
    // one rotated block: element W-1 of the running van Herk block, the suffix pass, elements 0..W-2 of the next block
    auto block = [&](auto check) {
        constexpr bool CHECK = decltype(check)::value;
#pragma unroll
        for (int j = 0; j < B; j++) {
            if (CHECK && p + j >= b) return;
            roll(p + j);
            ScoreT win = 0;
            if (MINI) {
                const ScoreT sc = score();
                if (j == 0) {
                    pre = W == 1 ? sc : smin(sc, pre);
                    win = pre;
                    buf[W - 1] = sc;
#pragma unroll
                    for (int q = W - 2; q >= 1; q--) buf[q] = smin(buf[q], buf[q + 1]);
                } else {
                    const int i = j - 1;
                    pre = i == 0 ? sc : smin(sc, pre);
                    win = smin(buf[i + 1], pre);
                    buf[i] = sc;
                }
            }
            tally(win);
        }
    };

    // head: K-1 bases that cannot end a k-mer — M-1 of them only feed F / R, the other W-1 are the first scores
    if (WARM) {
        (void)wcodes;
#pragma unroll
        for (int h = 0; h < K - 1; h++) {
            const uint32_t code = (uint32_t)(wcodes >> (2 * (K - 2 - h))) & 3u;
            f = (f << 2) | (uint64_t)code;
            r = (r >> 2) | ((uint64_t)((3u - code) << (2 * (K - 1) - 32)) << 32);
            if (MINI && h >= M - 1) {
                const int i = h - (M - 1);
                const ScoreT sc = score();
                pre = i == 0 ? sc : smin(sc, pre);
                buf[i] = sc;
            }
        }
    } else {
        const int e0 = ws + (MINI ? M - 1 : K - 1), e1 = b < e0 ? b : e0;
#pragma unroll 4
        for (; p < e1; p++) roll(p);
        if (MINI) {
#pragma unroll
            for (int i = 0; i < W - 1; i++) {
                if (p + i < b) {
                    roll(p + i);
                    const ScoreT sc = score();
                    pre = i == 0 ? sc : smin(sc, pre);
                    buf[i] = sc;
                }
            }
            p += W - 1;
        }
    }
    if (bail()) return false;
    while (p + B <= b) { block(FalseT{}); p += B; }
    if (p < b) block(TrueT{});
    if ((seen & 0x84u) || (RARE && rtop_min < RTOP_LIMIT)) return false;
    const int nk_i = WARM ? b - ws : b - (ws + K - 1);
    const uint64_t nk = nk_i > 0 ? (uint64_t)nk_i : 0;
    acc.n_kmers += nk; acc.n_not_rc += n_nrc;
    acc.ksum_lo += s_k;
    if (MINI) { acc.n_mini += nk; acc.msum += s_m; }
    return true;
}

// The clean walk for 33 <= K <= 63 (round 2, measured: C5 shape 195 -> 400 Gbases/s; no minimizers: the reference's BitKmer is a u64): F and R are
// 128-bit, the canonical k-mer is the smaller of the two as a 128-bit number (== the reference's lexicographic byte
// compare, A<C<G<T), checksums are the wrapping sums of its low and high 64 bits.  Same contract as walk_clean: any byte
// that is not ACGT/acgt makes it return false with acc untouched.  `comb2` = class bits | complement base pre-shifted for
// the 32-bit word of R's high half that receives it.
template <int K>
struct Clean2Shape {
    static constexpr int P1 = 2 * (K - 1) - 64;          // bit of R's high 64-bit word where the complement base enters
    static constexpr int SH = P1 >= 32 ? P1 - 32 : P1;   // ... inside its 32-bit half
    static constexpr bool ok = K >= 33 && K <= 63 && SH != 0 && SH != 2 && SH != 6;   // must not touch the class bits 0..2 and 7
};
template <int K>
__host__ __device__ __forceinline__ bool walk_clean2(const uint8_t* __restrict__ sb, const uint32_t* __restrict__ comb2, int ws, int b, Acc& acc) {
    static_assert(Clean2Shape<K>::ok, "two-word clean walker shape");
    constexpr int P1 = Clean2Shape<K>::P1, SH = Clean2Shape<K>::SH;
    constexpr uint32_t RMASK = 3u << SH;
    constexpr uint64_t M1 = (1ull << (2 * K - 64)) - 1;  // mask of F's high word
    constexpr int B = 8;
    uint64_t f0 = 0, f1 = 0, r0 = 0, r1 = 0, s0 = 0, s1 = 0;
    uint32_t seen = 0, n_nrc = 0;
    int p = ws;
    auto bail = [&]() -> bool {
#if defined(__CUDA_ARCH__)
        return __any_sync(__activemask(), (seen & 0x84u) != 0u);
#else
        return (seen & 0x84u) != 0u;
#endif
    };
    auto roll = [&](int pp) {
        const uint32_t u = comb2[sb[pp]];
        seen |= u;
        f1 = (f1 << 2) | (f0 >> 62);                 // (unmasked: old bases fall off the top)
        f0 = (f0 << 2) | (uint64_t)(u & 3u);
        r0 = (r0 >> 2) | (r1 << 62);
        r1 = (r1 >> 2) | ((uint64_t)(u & RMASK) << (P1 >= 32 ? 32 : 0));
    };
    auto tally = [&]() {
        const uint64_t fm1 = f1 & M1;
        const bool lt = lt62(fm1, r1) || (fm1 == r1 && f0 < r0);      // ties => was_rc = true (kmer.rs:124-128)
        s0 += lt ? f0 : r0;
        s1 += lt ? fm1 : r1;
        n_nrc += lt ? 1u : 0u;
    };
    {
        const int e0 = ws + K - 1, e1 = b < e0 ? b : e0;
#pragma unroll 4
        for (; p < e1; p++) roll(p);
    }
    if (bail()) return false;
    while (p + B <= b) {
#pragma unroll
        for (int j = 0; j < B; j++) { roll(p + j); tally(); }
        p += B;
    }
    for (; p < b; p++) { roll(p); tally(); }
    if (seen & 0x84u) return false;
    const int nk_i = b - (ws + K - 1);
    const uint64_t nk = nk_i > 0 ? (uint64_t)nk_i : 0;
    acc.n_kmers += nk; acc.n_not_rc += n_nrc;
    acc.ksum_lo += s0; acc.ksum_hi += s1;
    return true;
}

// =============================================================================== look-back
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// ---- the ring of tile slots ---------------------------------------------------------------
// Tile t uses slot t & slot_mask; its flag / look-back word carries the generation epoch + (t >> slot_shift), so a slot that
// still holds an older tile reads as "not published" and the ring never has to be cleared (streams of any length).
__device__ __forceinline__ TileSlot* slot_of(const Params& P, uint64_t t) { return &P.slots[t & P.slot_mask]; }
__device__ __forceinline__ uint32_t epoch_of(const Params& P, uint32_t epoch, uint64_t t) { return (epoch + (uint32_t)(t >> P.slot_shift)) & 0x3FFFFFFFu; }

// ---- FASTA look-back on one 64-bit word per tile (round 2) ---------------------------------------------------------------
// What a FASTA tile needs from everything before it is the header state only: does it start inside a header line, and where did
// the newest header line end (the start of the current sequence region).  Record starts and newlines are plain sums and go to
// Params::fa_totals.  combine(a, b) on that state: b's header event if it has one; else, if a ends inside a header, that header
// ends at b's first newline (or still has not ended); else a's.  So a tile publishes ONE word — generation << 34 | state << 32
// | kind << 30 | offset — first its aggregate (state 1: kind 0 nothing, 1 first newline at `offset`, 2 newest header ended at
// `offset`, 3 ends inside a header; offsets are tile-relative), later the inclusive prefix (state 2: kind 0 no header yet,
// 2 newest header ended `offset` bytes before the END of the tile (saturating: only distances up to the halo matter), 3 inside a
// header).  A look-back step is one relaxed load per predecessor, 128 predecessors per step, and a shuffle tree over
// (kind, position) pairs — instead of 256-byte slots and a tree over eight 64-bit fields per predecessor, which cost the
// 10 kbp FASTA shape half of every CTA's time (NTG_STATS, profiles/r2a_*).
constexpr int LBQ_G = 4;                            // predecessors per lane and look-back step
constexpr uint32_t FA_OFF_BITS = 30;
constexpr uint64_t FA_FAR = (1ull << FA_OFF_BITS) - 1;
struct FaState { uint32_t kind; int64_t pos; };                    // pos: stream position (kind 1: first newline, kind 2: header-ending newline)
__host__ __device__ __forceinline__ FaState fa_combine(const FaState& a, const FaState& b) {          // a = earlier span, b = later span
    // aggregate semantics: kind 0 = no event, no newline; 1 = no header event, first newline at pos; 2 = header ended at pos; 3 = ends in a header
    // (an inclusive prefix never has kind 1: folded against what precedes it a bare newline is no event)
    if (b.kind >= 2) return b;
    if (a.kind == 3) return b.kind == 1 ? FaState{2u, b.pos} : FaState{3u, 0};
    if (a.kind == 2) return a;
    return b.kind == 1 && a.kind == 1 ? a : (a.kind ? a : b);      // no header event on either side: keep the EARLIEST first newline
}
// the header state of a span as the look-back word carries it (SState::hdr / first_nl -> kind, position)
__host__ __device__ __forceinline__ FaState fa_of(const SState& s) {
    if (s.hdr == INHDR) return FaState{3u, 0};
    if (s.hdr != NONE) return FaState{2u, (int64_t)s.hdr};
    if (s.first_nl != NONE) return FaState{1u, (int64_t)s.first_nl};
    return FaState{0u, 0};
}
__device__ __forceinline__ unsigned long long fa_word(uint32_t gen, uint32_t state, uint32_t kind, uint64_t off) {
    return ((unsigned long long)gen << 34) | ((unsigned long long)state << 32) | ((unsigned long long)kind << FA_OFF_BITS) | (off < FA_FAR ? off : FA_FAR);
}
// header state in front of tile t (t > 0) as an SState whose hdr field is NONE / INHDR / the stream position of the newest
// header-ending newline (a position farther back than FA_FAR bytes comes back as "FA_FAR before the tile": beyond any halo)
__device__ __forceinline__ SState fasta_lookback(const Params& P, uint64_t t, uint32_t epoch, uint32_t lane, uint32_t* dbg = nullptr) {
    const int64_t TB = (int64_t)P.tile_bytes;
    FaState suffix{0u, 0};
    int64_t base = (int64_t)t - 1;
    uint32_t backoff = 32;
    for (;;) {
        const int64_t j0 = base - (int64_t)lane * LBQ_G;
        unsigned long long w[LBQ_G];
#pragma unroll
        for (int g = 0; g < LBQ_G; g++) { const int64_t j = j0 - g; w[g] = j >= 0 ? ld_relaxed_u64(&P.cw[(uint64_t)j & P.slot_mask]) : 0ull; }
        FaState mine{0u, 0};
        uint32_t lane_state = 0;                                    // 0: G aggregates folded, 1: reached an inclusive prefix, 2: blocked
#pragma unroll
        for (int g = 0; g < LBQ_G; g++) {
            const int64_t j = j0 - g;
            uint32_t st = 2, kind = 0; int64_t pos = 0;             // before the first tile: inclusive(nothing)
            if (j >= 0) {
                st = ((uint32_t)(w[g] >> 34) == epoch_of(P, epoch, (uint64_t)j)) ? (uint32_t)(w[g] >> 32) & 3u : 0u;
                kind = (uint32_t)(w[g] >> FA_OFF_BITS) & 3u;
                const int64_t off = (int64_t)(w[g] & FA_FAR);
                pos = st == 2 ? (j + 1) * TB - off : j * TB + off;   // inclusive: distance back from the tile's end; aggregate: tile-relative
            }
            if (lane_state == 0) {
                if (st == 0) lane_state = 2;
                // (an aggregate with a header event of its own absorbs everything before it — fa_combine(a, b) = b for
                //  b.kind >= 2 — so it ends the look-back like an inclusive prefix does: no wait for farther predecessors)
                else { mine = fa_combine(FaState{kind, pos}, mine); if (st == 2 || kind >= 2u) lane_state = 1; }
            }
        }
        const uint32_t inc_mask = __ballot_sync(0xffffffffu, lane_state == 1), blk_mask = __ballot_sync(0xffffffffu, lane_state == 2);
        const int first_inc = inc_mask ? __ffs((int)inc_mask) - 1 : 32;
        const uint32_t need = first_inc >= 31 ? 0xffffffffu : ((2u << first_inc) - 1u);
        if (blk_mask & need) { if (dbg) dbg[0]++; __nanosleep(backoff); backoff = backoff < 256 ? backoff * 2 : 256; continue; }
        if (dbg) dbg[1]++;
        const int top = first_inc < 32 ? first_inc : 31;
        FaState acc = (int)lane <= top ? mine : FaState{0u, 0};      // lanes above `top` contribute the identity
        // ordered tree reduction: lane l holds the tiles base-4l ..; higher lanes are EARLIER tiles
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int src = (int)lane + d < 32 ? (int)lane + d : (int)lane;
            FaState e; e.kind = __shfl_sync(0xffffffffu, acc.kind, src); e.pos = __shfl_sync(0xffffffffu, acc.pos, src);
            if ((int)lane + d < 32) acc = fa_combine(e, acc);
        }
        acc.kind = __shfl_sync(0xffffffffu, acc.kind, 0); acc.pos = __shfl_sync(0xffffffffu, acc.pos, 0);
        suffix = fa_combine(acc, suffix);
        if (first_inc < 32) break;
        base -= 32 * LBQ_G;
    }
    SState pre = identity_state();
    pre.hdr = suffix.kind == 3 ? INHDR : (suffix.kind == 2 ? (uint64_t)suffix.pos : NONE);
    return pre;
}

// ---- FASTQ look-back on one 64-bit word per tile (round 2) ------------------------------------------------------
// FASTQ line roles need the newline ordinal mod 4 only, and the end-of-stream rules (fastq.rs:337-356) only whether at least
// four newlines exist: the prefix is the 3-bit monoid  enc(c) = (c & 3) | (c >= 4 ? 4 : 0),  enc(a + b) = sum of the low parts
// mod 4, saturation bit = either saturated or the low parts reach 4.  A tile publishes ONE word, first with its own count
// (state 1), later with the inclusive prefix (state 2):  generation (30 bits) << 34 | state << 32 | enc.  The word is its own
// payload: a single relaxed load per predecessor, 128 predecessors per step (4 per lane, all loads in flight together), and
// the fold is two warp reductions - instead of a 64-byte state per predecessor behind a flag and a shuffle tree of combines.
// The r1 look-back of the headline run took 33 000 cycles per tile (NTG_STATS) and the tile loop ran 12 % faster without it.
// The positions of the last four newlines before a tile (line lengths of the lines that cross into it) come from the
// aggregates of its nearest predecessors (collect_last).
__host__ __device__ __forceinline__ uint32_t enc_count(uint64_t c) { return (uint32_t)(c & 3u) | (c >= 4 ? 4u : 0u); }
__host__ __device__ __forceinline__ uint32_t enc_combine(uint32_t a, uint32_t b) {
    const uint32_t s = (a & 3u) + (b & 3u);
    return (s & 3u) | ((a | b) & 4u) | (s >= 4 ? 4u : 0u);
}
__device__ __forceinline__ unsigned long long cw_make(uint32_t gen, uint32_t state, uint32_t enc) {
    return ((unsigned long long)gen << 34) | ((unsigned long long)state << 32) | enc;
}
// encoded exclusive prefix of tile t (t > 0); whole warp.  Ends with a fence: the aggregates of every predecessor are visible.
__device__ __forceinline__ uint32_t fastq_lookback(const Params& P, uint64_t t, uint32_t epoch, uint32_t lane) {
    uint32_t suffix = 0;
    int64_t base = (int64_t)t - 1;                                  // nearest predecessor not folded yet
    uint32_t backoff = 32;
    for (;;) {
        const int64_t j0 = base - (int64_t)lane * LBQ_G;            // this lane: tiles j0, j0-1, ... (nearest first)
        unsigned long long w[LBQ_G];
#pragma unroll
        for (int g = 0; g < LBQ_G; g++) { const int64_t j = j0 - g; w[g] = j >= 0 ? ld_relaxed_u64(&P.cw[(uint64_t)j & P.slot_mask]) : 0ull; }
        uint32_t lane_enc = 0, lane_state = 0;                      // 0: G aggregates folded, 1: reached an inclusive prefix, 2: blocked
#pragma unroll
        for (int g = 0; g < LBQ_G; g++) {
            const int64_t j = j0 - g;
            uint32_t st = 2, en = 0;                                 // before the first tile: inclusive(identity)
            if (j >= 0) { st = ((uint32_t)(w[g] >> 34) == epoch_of(P, epoch, (uint64_t)j)) ? (uint32_t)(w[g] >> 32) & 3u : 0u; en = (uint32_t)w[g] & 7u; }
            if (lane_state == 0) {
                if (st == 0) lane_state = 2;
                else { lane_enc = enc_combine(lane_enc, en); if (st == 2) lane_state = 1; }
            }
        }
        const uint32_t inc_mask = __ballot_sync(0xffffffffu, lane_state == 1), blk_mask = __ballot_sync(0xffffffffu, lane_state == 2);
        const int first_inc = inc_mask ? __ffs((int)inc_mask) - 1 : 32;
        const uint32_t need = first_inc >= 31 ? 0xffffffffu : ((2u << first_inc) - 1u);
        if (blk_mask & need) { __nanosleep(backoff); backoff = backoff < 256 ? backoff * 2 : 256; continue; }
        const uint32_t mine = ((int)lane <= first_inc) ? lane_enc : 0u;
        const uint32_t low = __reduce_add_sync(0xffffffffu, mine & 3u), sat = __reduce_or_sync(0xffffffffu, mine & 4u);
        suffix = enc_combine((low & 3u) | sat | (low >= 4 ? 4u : 0u), suffix);
        if (first_inc < 32) break;
        base -= 32 * LBQ_G;
    }
    __threadfence();                                                // words observed -> aggregates (slot->agg) readable
    return suffix;
}
// the four most recent newline positions before tile `t_excl` (newest first, NONE-padded) from the aggregates of the tiles
// before it; every one of them has been published (fastq_lookback has seen their words).  One thread.
constexpr int LAST_WALK_MAX = 4096;
__device__ __forceinline__ void collect_last(const Params& P, uint64_t t_excl, uint64_t* last, uint32_t filled, uint32_t& slow) {
    int64_t j = (int64_t)t_excl - 1;
    for (int steps = 0; filled < 4 && j >= 0; j--, steps++) {
        if (steps >= LAST_WALK_MAX) { slow |= FLAG_HALO_OVERFLOW; break; }     // a line spanning thousands of tiles: exact path
        const SState* a = &slot_of(P, (uint64_t)j)->agg;
        const uint64_t c = a->count;
        for (uint32_t i = 0; i < 4 && i < c && filled < 4; i++) last[filled++] = a->last[i];
    }
    for (; filled < 4; filled++) last[filled] = NONE;
}

// Exclusive prefix of tile t in the form the line events need, by one whole warp (result in every lane).  FASTQ: `count` holds
// enc(newlines before t) - its low two bits are the line phase - and last[] the last four newline positions; FASTA: the
// general state.
__device__ __forceinline__ SState tile_prefix(const Params& P, uint64_t t, uint32_t epoch, uint32_t lane, bool fasta, uint32_t& slow, uint32_t* dbg = nullptr) {
    SState pre = identity_state();
    if (t == 0) return pre;
    if (fasta) return fasta_lookback(P, t, epoch, lane, dbg);
    pre.count = fastq_lookback(P, t, epoch, lane);
    if (lane == 0) collect_last(P, t, pre.last, 0, slow);
#pragma unroll
    for (int i = 0; i < 4; i++) pre.last[i] = __shfl_sync(0xffffffffu, pre.last[i], 0);
    return pre;
}
// one thread: make the tile's aggregate visible to the look-backs of its successors
__device__ __forceinline__ void publish_aggregate(const Params& P, uint64_t t, uint32_t epoch, const SState& agg, bool fasta) {
    if (fasta) {
        // header state of the tile alone (see fasta_lookback); the sums go to the pass totals
        const uint64_t ts = t * (uint64_t)P.tile_bytes;
        uint32_t kind = 0; uint64_t off = 0;
        if (agg.hdr == INHDR) kind = 3;
        else if (agg.hdr != NONE) { kind = 2; off = agg.hdr - ts; }
        else if (agg.first_nl != NONE) { kind = 1; off = agg.first_nl - ts; }
        st_release_u64(&P.cw[t & P.slot_mask], fa_word(epoch_of(P, epoch, t), 1, kind, off));
        if (agg.n_starts) atomicAdd(&P.fa_totals[0], (unsigned long long)agg.n_starts);
        if (agg.count) atomicAdd(&P.fa_totals[1], (unsigned long long)agg.count);
        return;
    }
    TileSlot* slot = slot_of(P, t);
    slot->agg = agg;
    st_release_u64(&P.cw[t & P.slot_mask], cw_make(epoch_of(P, epoch, t), 1, enc_count(agg.count)));
}
// one thread: publish the inclusive prefix of tile t (and hand the stream's final state to k_finalize)
__device__ __forceinline__ void publish_inclusive(const Params& P, uint64_t t, uint32_t epoch, const SState& pre, const SState& agg, bool fasta) {
    SState inc;
    if (fasta) {
        inc = combine(pre, agg);                                   // (only hdr is meaningful: pre carries nothing else)
        const uint64_t te = (t + 1) * (uint64_t)P.tile_bytes;
        const uint32_t kind = inc.hdr == INHDR ? 3u : (inc.hdr != NONE ? 2u : 0u);
        st_release_u64(&P.cw[t & P.slot_mask], fa_word(epoch_of(P, epoch, t), 2, kind, kind == 2 ? te - inc.hdr : 0));
    } else {
        inc = identity_state();
        inc.count = enc_combine((uint32_t)pre.count, enc_count(agg.count));
        st_release_u64(&P.cw[t & P.slot_mask], cw_make(epoch_of(P, epoch, t), 2, (uint32_t)inc.count));
        if (t + 1 == P.num_tiles) {                                // newest four newline positions of the whole stream
            int f = 0;
            for (int i = 0; i < 4 && (uint64_t)i < agg.count; i++) inc.last[f++] = agg.last[i];
            for (int i = 0; i < 4 && f < 4; i++) inc.last[f++] = pre.last[i];
        }
    }
    if (t + 1 == P.num_tiles) *P.final_state = inc;
}

// =============================================================================== the kernel
__device__ __forceinline__ uint8_t byte_at(const Params& P, const uint8_t* sb, uint64_t tile_start, uint32_t halo, uint64_t gpos) {
    // global position -> byte, from shared memory when resident, else from global memory
    if (gpos + halo >= tile_start && gpos < tile_start + P.tile_bytes) return sb[(int64_t)gpos - (int64_t)tile_start];
    return (gpos < P.n && gpos >= P.gmin) ? P.bytes[gpos] : 0;
}
__device__ __forceinline__ uint8_t class_of(int i) {
    uint8_t c = c_ncls[i];
    if (c == 5) c = (i == '\r' || i == '\n') ? 0x86 : 0x85;
    return c;
}

template <int KW, bool MINI, int W>
__device__ __noinline__ void walk_slow(const uint8_t* sb, const uint8_t* lut, int ws, int a, int b, const Params& P, Acc& acc, bool count_bases) {
    walk<KW, MINI, W>(sb, lut, ws, a, b, P, acc, count_bases);
}
template <int K, int M>
__device__ __noinline__ bool walk_fast_cold(const uint8_t* sb, const uint8_t* lut, const uint32_t* rins, int ws, int b, Acc& acc, uint32_t* seen_out) {
    FastLuts L{lut, rins, nullptr};
    return walk_fast<K, M>(sb, L, ws, b, acc, seen_out);
}

// One sequence-line fragment sb[a..b): warm-up, then (kernels with a constant-folded shape, FK > 0) the clean walker;
// items with a non-ACGT base fall to the constant-folded walker that tracks them, items with deleted bytes
// (whitespace inside the item) and all other shapes to the generic walker.
// `mode` (per thread, kept across items): non-zero after an item that the clean walker could not take.  While any lane of
// the warp is in that state the warp skips the clean attempt - on data full of N (BASELINE C4: 78 % of the reads)
// every warp would otherwise walk each item twice; a clean item puts the lane back.
template <int KW, bool MINI, int W, int FK, int FM>
__device__ __forceinline__ void run_item(const uint8_t* sb, const uint8_t* lut, const uint32_t* rins, const uint32_t* comb, int a, int b, int lo,
                                         bool lo_exact, const Params& P, Acc& acc, bool fasta, uint32_t& slow, uint32_t& mode) {
    if (b > a && sb[b - 1] == '\r') b--;                   // a trailing '\r' is deleted by normalize: nothing to walk
    if (b <= a) return;
    constexpr bool ONE = FK >= 21 && FK <= 31;              // constant-folded one-word shapes (clean / fast walkers)
    int got = 0; uint64_t wcodes = 0;
    const int ws = ONE ? find_ws_codes(sb, lut, a, lo, lo_exact, (int)P.k, slow, got, wcodes) : find_ws(sb, lut, a, lo, lo_exact, (int)P.k, slow);
    if (FK > 32) {
        if (!__any_sync(__activemask(), mode != 0u) && walk_clean2<(FK > 32 ? FK : 51)>(sb, comb, ws, b, acc)) {
            if (fasta) acc.n_bases += (uint64_t)(b - a);   // no deleted bytes in [ws,b): every byte of the item is a base
            return;
        }
        mode = 1u;                                             // an item the clean walker refused: the warp's next item goes straight
        uint32_t any_bad = 0;                                  // to the generic walker, and comes back once an item was all ACGT
        for (int q = ws; q < b; q++) any_bad |= lut[sb[q]];
        if (any_bad <= 3u) mode = 0u;
        walk<KW, MINI, W>(sb, lut, ws, a, b, P, acc, fasta);
    } else if (ONE) {
        constexpr int CK = ONE ? FK : 21, CM = ONE ? FM : 0;
        bool done = false;
        if (!__any_sync(__activemask(), mode != 0u)) {
            // a warm-up that crosses deleted bytes (wrapped FASTA: the previous line break) is fed from the register
            if (got == CK - 1 && a - ws != got) done = walk_clean<CK, CM, true>(sb, comb, a, b, wcodes, acc);
            else done = walk_clean<CK, CM, false>(sb, comb, ws, b, 0, acc);
        }
        if (!done) {
            uint32_t seen = 0x80u;
            done = walk_fast_cold<CK, CM>(sb, lut, rins, ws, b, acc, &seen);
            mode = seen > 3u ? 1u : 0u;                     // (stays set when walk_fast gave up on a deleted byte)
        }
        if (done) {
            if (fasta) acc.n_bases += (uint64_t)(b - a);   // no deleted bytes in [a,b): every byte is a base
            return;
        }
        walk_slow<KW, MINI, W>(sb, lut, ws, a, b, P, acc, fasta);
    } else {
        walk<KW, MINI, W>(sb, lut, ws, a, b, P, acc, fasta);
    }
}

// FASTQ line phase inferred from the tile alone: which p makes role(i) = (p + i) & 3 consistent with the first bytes
// of the first lines that start in this tile ('@' at role 0, '+' at role 2)?  Returns 0..3, or 4 when fewer than four
// line starts are visible or the evidence is not unique.  Used only to START early; the true phase (newline ordinal
// from the look-back) is checked afterwards and a mismatch raises FLAG_SPEC_MISS.
// Must be called by whole (converged) warps: lane j < 8 looks at the j-th visible line start, the evidence is AND-reduced.
template <typename NLT>
__device__ __forceinline__ uint32_t guess_phase(const NLT* __restrict__ nl, const uint8_t* __restrict__ sb, uint32_t Cs,
                                                uint32_t avail, bool line0_starts_here) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i = (line0_starts_here ? 0u : 1u) + lane;
    uint32_t ok = 0xF;
    bool valid = false;
    if (lane < 8 && i <= Cs) {
        const uint32_t s = i ? (uint32_t)nl[i - 1] + 1u : 0u;
        if (s < avail) {                                     // (line starts increase with i: the valid lanes form a prefix)
            valid = true;
            const uint8_t c = sb[s];
            ok &= ~((c != '@' ? 1u : 0u) << ((0u - i) & 3u));
            ok &= ~((c != '+' ? 1u : 0u) << ((2u - i) & 3u));
        }
    }
    const uint32_t seen = (uint32_t)__popc(__ballot_sync(0xffffffffu, valid));
    ok = __reduce_and_sync(0xffffffffu, ok);
    if (seen < 4 || __popc(ok) != 1) return 4;
    return (uint32_t)__ffs((int)ok) - 1u;
}
// start of the line that continues into the tile, found in the back halo (tile-relative, < 0); exact when a newline is
// visible, else the halo limit
__device__ __forceinline__ void halo_line_start(const uint8_t* __restrict__ sb, uint32_t halo, int& lo, bool& lo_exact) {
    for (int p = -1; p >= -(int)halo; p--)
        if (sb[p] == '\n') { lo = p + 1; lo_exact = true; return; }
    lo = -(int)halo; lo_exact = false;
}


#ifndef NTG_EARLY_TICKET
#define NTG_EARLY_TICKET 0                           // 1: the coordinator claims the CTA's next tile right after publishing this one's aggregate and
#endif                                               //    prefetches it into L2; 0: walker thread 0 claims it at the end of its walk (no prefetch)
#ifndef NTG_STATS
#define NTG_STATS 0                                  // 1: per-CTA cycle accounting into tallies[9..15] (ntg_tallies.reserved[2..6]):
#endif                                               //    [9] sum of CTA lifetimes, [10] coordinator cycles inside look-backs, [11] look-backs,
                                                     //    [12] longest CTA lifetime, [14] thread 0 at the end-of-walk barrier, [15] thread 0 walking
// Line events (fastq.rs:240-285) of line i < 4 of a tile, from global memory: start byte ('@' at role 0, '+' at role 2),
// n_bases of a sequence line and the length check / n_records of a quality line that END in the tile.  `nl4` = the tile's
// first four newline offsets (tile-relative), Cs = number of newlines in the tile, `pre` = the tile's exclusive prefix
// (line roles come from its newline count, lines that began in earlier tiles from its last[] positions).  i <= Cs.
// __host__ __device__: tests/cpp/test_walkers.cu checks it against a direct evaluation of the definition.
__host__ __device__ __forceinline__ void first_lines_event(const uint8_t* __restrict__ bytes, uint64_t gmin, uint64_t tile_start, const uint32_t* nl4, uint32_t Cs,
                                                           uint32_t avail, bool line0_starts_here, const SState& pre, uint32_t i, Acc& acc,
                                                           uint32_t& slow, unsigned long long* err_key) {
    const uint32_t ord0 = (uint32_t)(pre.count & 3);
    auto nl = [&](uint32_t j) -> uint64_t { return tile_start + nl4[j]; };                        // j < min(Cs, 4)
    auto prev_nl = [&](uint32_t back) -> uint64_t {                                               // `back` newlines before newline i
        if (i >= back) return nl(i - back);
        const uint32_t r = back - i - 1;
        return r < 4 ? pre.last[r] : NONE;
    };
    auto cr_before = [&](uint64_t q, uint64_t prevq) -> uint32_t {                                // trim_cr on the line (prevq, q)
        const uint64_t ls = prevq == NONE ? 0 : prevq + 1;
        if (q <= ls) return 0u;
        if (q - 1 < gmin) { slow |= FLAG_HALO_OVERFLOW; return 0u; }                              // (a streamed window that no longer holds the byte)
        return bytes[q - 1] == '\r' ? 1u : 0u;
    };
    const uint32_t role = (ord0 + i) & 3;
    auto rec_start = [&]() -> uint64_t { const uint64_t p = prev_nl(role + 1); return p == NONE ? 0 : p + 1; };   // line i - role starts the record
    const uint64_t s = i ? nl(i - 1) + 1 : tile_start;
    const bool starts = (i > 0 || line0_starts_here) && (s - tile_start) < avail;
    if (starts) {
        const uint8_t c = bytes[s];
        if (role == 0 && c != '@') note_parse_error(err_key, slow, s, 0);
        if (role == 2 && c != '+') note_parse_error(err_key, slow, rec_start(), 1);
    }
    if (i < Cs) {
        const uint64_t q = nl(i);
        if (role == 1) {
            const uint64_t p1 = prev_nl(1);
            const uint64_t ls = p1 == NONE ? 0 : p1 + 1;
            acc.n_bases += (q - ls) - cr_before(q, p1);
        } else if (role == 3) {
            const uint64_t q2 = prev_nl(1), q1 = prev_nl(2), q0 = prev_nl(3);
            if (q2 == NONE || q1 == NONE || q0 == NONE) note_parse_error(err_key, slow, 0, 0);    // inconsistent state: replay from the start
            else {
                const uint64_t seq_len = (q1 - q0 - 1) - cr_before(q1, q0);
                const uint64_t qual_len = (q - q2 - 1) - cr_before(q, q2);
                if (seq_len != qual_len) note_parse_error(err_key, slow, rec_start(), 2);
                acc.n_records++;
            }
        }
    }
}

// Coordinator warp: resolve the tile deferred at the previous P2c — look-back, inclusive prefix, check of the speculated
// line phase, and the line events (fastq.rs:240-285) of the tile's first four lines, the ones that may need the prefix.
// The tile's bytes have left shared memory by now: the few bytes involved are read from global memory (L2).
__device__ __noinline__ void resolve_pending(const Params& P, Smem& S, uint32_t epoch, uint32_t lane, Acc& acc, uint32_t& slow) {
    const uint64_t t = S.pend_t;
    const SState agg = S.pend_agg;
    const SState pre = tile_prefix(P, t, epoch, lane, false, slow);      // (only FASTQ tiles are deferred)
    if (lane == 0) publish_inclusive(P, t, epoch, pre, agg, false);
    const uint32_t ord0 = (uint32_t)(pre.count & 3);
    if (S.pend_guess != ord0) slow |= FLAG_SPEC_MISS;
    const uint32_t Cs = S.pend_cs;
    if (lane < 4 && lane <= Cs)
        first_lines_event(P.bytes, P.gmin, t * (uint64_t)P.tile_bytes, S.pend_nl4, Cs, S.pend_avail, S.pend_line0 != 0, pre, lane, acc, slow, P.err_key);
    __syncwarp();
    if (lane == 0) S.pend_valid = 0;
    __syncwarp();
}



template <int KW, bool MINI, int W, int FK, int FM>
__global__ void __launch_bounds__(NT, CTAS_PER_SM) k_fused(const Params P, const uint64_t tile_begin, const uint64_t tile_end,
                                                 const uint32_t epoch, uint32_t* __restrict__ ticket) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    for (int i = tid; i < 256; i += NT) {
        const uint8_t c = class_of(i);
        S.lut[i] = c;
        const uint32_t ri = (FK > 32) ? ((3u - (c & 3u)) << Clean2Shape<(FK > 32 ? FK : 51)>::SH)
                                      : (FK >= 17) ? ((3u - (c & 3u)) << (2 * ((FK >= 17 && FK <= 32 ? FK : 17) - 1) - 32)) : 0u;
        S.rins[i] = ri;
        S.comb[i] = ri | c;
    }
    if (tid == 0) {
        mbar_init(&S.bar, 1); fence_mbar_init(); S.pend_valid = 0; S.arrived = 0;
    }
    __syncthreads();
    uint32_t parity = 0, slow = 0, my_seq = 0, mode = 0;
#if NTG_STATS
    const long long st_t0 = clock64();
    long long st_lb = 0, st_wait = 0, st_walk = 0, st_mark = 0; uint32_t st_nlb = 0;
    uint32_t st_dbg[2] = {0, 0}; long long st_d = 0;            // NTG_STATS == 3: look-back polls / steps, claim -> aggregate cycles
    long long st_p0 = 0, st_p1 = 0, st_p2 = 0, st_ph = 0;      // NTG_STATS == 2: thread 0's cycles in P0 / P1 / P2 (replace the look-back counters)
#endif
    Acc acc;
    const uint8_t* sb = S.tile;
    const bool fasta = P.format == NTG_FMT_FASTA;
    const bool is_coord = tid >= NTW;                 // the coordinator warp: look-back, line events (FASTQ)
    const bool spec = P.spec != 0 && !fasta;          // walkers start on an inferred phase, the coordinator verifies it
    if (tid == 0) S.prefix_seq = 0;

    // Tiles are handed out by an atomic ticket (round 2, measured: C2 +2 %, C3 +57 % over static round-robin: with aggregates
    // published early and the look-back deferred a late claimer no longer stalls its successors, and SMs that run ahead take
    // more tiles instead of polling for slower predecessors).  The coordinator claims the CTA's NEXT ticket as soon as it has
    // published this tile's aggregate and prefetches that tile into L2 (NTG_EARLY_TICKET).
    if (tid == 0) S.tile_idx_next = atomicAdd(ticket, 1u);
    __syncthreads();
    uint32_t next_ticket = S.tile_idx_next;
    for (;;) {
        const uint64_t t = tile_begin + (uint64_t)next_ticket;
        if (t >= tile_end) break;
        const uint32_t TB = P.tile_bytes;
        const uint64_t tile_start = t * (uint64_t)TB;
        const uint32_t avail = (uint32_t)min((uint64_t)TB, P.n - tile_start);
        const uint32_t halo = t > 0 ? HALO : 0;
        const uint32_t bulk = avail & ~15u;

        if (tid == 0) S.n_long = 0;
#if NTG_STATS
        st_ph = clock64();
        const long long st_ph0 = st_ph;
#endif
        // ---- P0: stage the tile (+ back halo) with one bulk async copy
        if (tid == 0 && halo + bulk) {
            mbar_expect_tx(&S.bar, halo + bulk);
            bulk_g2s(S.halo + (HALO - halo), P.bytes + tile_start - halo, halo + bulk, &S.bar);
        }
        if (avail < TB) for (uint32_t i = bulk + tid; i < TB; i += NWK) S.tile[i] = i < avail ? P.bytes[tile_start + i] : 0;
        if (t == 0) for (int i = tid; i < HALO; i += NWK) S.halo[i] = 0;
        if (halo + bulk) { mbar_wait(&S.bar, parity); parity ^= 1; }
        tile_sync();
#if NTG_STATS
        { const long long c = clock64(); st_p0 += c - st_ph; st_ph = c; }
#endif

        // ---- P1: newline scan, 256 B rows, row = round * NT + tid (rotated word order: conflict-free LDS.32)
        const uint32_t nrows = TB / ROWB;
        uint32_t cnt[MAXROUNDS];
        uint64_t wmask[MAXROUNDS];
#pragma unroll
        for (int rd = 0; rd < MAXROUNDS; rd++) {
            cnt[rd] = 0; wmask[rd] = 0;
            const uint32_t rowi = rd * NTW + tid;
            if (tid < NTW && rowi < nrows) scan_row(reinterpret_cast<const uint32_t*>(S.tile) + rowi * ROWW, lane, cnt[rd], wmask[rd]);
        }
#if NTG_STATS
        { const long long c = clock64(); st_p1 += c - st_ph; st_ph = c; }
#endif
        // ---- P2: ordered newline list (rows are ordered round-major: one scan of the packed per-round counts)
        // The two per-round counts share one u32 for a single block scan: round 0 covers NTW rows (at most NTW * 256 = 73 728
        // newlines: 17 bits), round 1 the remaining TILE / 256 - NTW rows (at most 12 288: 14 bits) -> 17 + 15 bits, no wrap
        // even for a tile made of newlines (a run of blank lines is valid FASTA / FASTQ tail).
        static_assert(MAXROUNDS == 2 && NTW * ROWB < (1 << 17) && (TILE / ROWB - NTW) * ROWB < (1 << 15) && TILE / ROWB >= NTW, "packed scan: 17 + 15 bits");
        uint32_t Cpacked;
        const uint32_t offp = block_excl_scan(cnt[0] | (cnt[1] << 17), &Cpacked, S.warp_tmp[0]);
        const uint32_t C0 = Cpacked & 0x1FFFFu, C = C0 + (Cpacked >> 17);
        const bool overflow = C > NLMAX;
        if (overflow) slow |= FLAG_NL_OVERFLOW;
        if (!overflow) {
#pragma unroll
            for (int rd = 0; rd < MAXROUNDS; rd++) {
                if (!cnt[rd]) continue;
                const uint32_t rowi = rd * NTW + tid;
                const uint32_t* row = reinterpret_cast<const uint32_t*>(S.tile) + rowi * ROWW;
                uint32_t o = rd == 0 ? (offp & 0x1FFFFu) : C0 + (offp >> 17);
                uint64_t mk = wmask[rd];
                while (mk) {
                    const int jj = __ffsll((long long)mk) - 1;
                    mk &= mk - 1;
                    const uint32_t wv = row[jj];
#pragma unroll
                    for (int bsel = 0; bsel < 4; bsel++)
                        if (((wv >> (8 * bsel)) & 0xFF) == '\n') S.nl[o++] = (uint32_t)(rowi * ROWB + jj * 4 + bsel);
                }
            }
        }
        tile_sync();
#if NTG_STATS
        { const long long c = clock64(); st_p2 += c - st_ph; st_ph = c; }
#endif
        const uint32_t Cs = overflow ? 0 : C;                          // lines are only interpreted when the list is complete
        // line i (0..Cs) spans (nl[i-1], nl[i]) ; helpers on tile-relative coordinates
        auto line_start_rel = [&](uint32_t i) -> int { return i ? (int)S.nl[i - 1] + 1 : 0; };   // for i == 0: start of the in-tile fragment
        auto line_end_rel = [&](uint32_t i) -> int { return i < Cs ? (int)S.nl[i] : (int)avail; };
        const bool line0_starts_here = (t == 0) || (sb[-1] == '\n');

        // ---- P2b: FASTA start events that do not need the prefix
        uint32_t my_last_start = 0, my_nstarts = 0;                    // (line index + 1) of the last start seen by this thread
        if (fasta) {
            for (uint32_t i = tid; i <= Cs; i += NWK) {
                const int s = line_start_rel(i);
                const bool starts = (i > 0 || line0_starts_here) && (uint32_t)s < avail;
                if (starts && sb[s] == '>') { my_last_start = i + 1; my_nstarts++; }
            }
        }
        uint32_t last_start1 = 0, nstarts_tot = 0;
        if (fasta) {
            last_start1 = block_incl_max(my_last_start, S.warp_tmp[1]);
            uint32_t tot; block_excl_scan(my_nstarts, &tot, S.warp_tmp[2]); nstarts_tot = tot;
            if (tid == NWK - 1) S.bcast[0] = (int32_t)last_start1;
            tile_sync();
            last_start1 = (uint32_t)S.bcast[0];
        }

        // ---- P2c: publish aggregate, decoupled look-back (coordinator warp, 32 predecessors per step), publish inclusive prefix.
        // Speculative FASTQ: when the tile offers a unique local line phase nobody in this CTA needs the tile's prefix
        // now, so its look-back is DEFERRED by one tile: the coordinator publishes the aggregate at once, then resolves the
        // tile it deferred last time (whose predecessors published their aggregates a whole tile ago: no waiting on the
        // skew between CTAs, and the end-of-tile barrier no longer waits for a look-back that has just begun), verifies
        // that tile's guess and does the events of its first four lines from global memory.
        const uint32_t guess = spec ? guess_phase(S.nl, sb, Cs, avail, line0_starts_here) : 4u;
        const bool defer = spec && guess != 4u && t > 0;      // (tile 0 has nothing to look back at)
        SState pre = identity_state();
        if (is_coord) {
            SState agg = identity_state();
            agg.count = C;
            if (!overflow) for (uint32_t j = 0; j < 4 && j < C; j++) agg.last[j] = tile_start + S.nl[C - 1 - j];
            if (fasta) {
                agg.n_starts = nstarts_tot;
                agg.first_nl = Cs ? tile_start + S.nl[0] : NONE;
                if (last_start1) { const uint32_t Lh = last_start1 - 1; agg.hdr = (Lh < Cs) ? tile_start + S.nl[Lh] : INHDR; }
            }
            if (lane == 0) {
                publish_aggregate(P, t, epoch, agg, fasta);
#if NTG_EARLY_TICKET
                // the next tile of this CTA: claimed now so that it can be pulled into L2 while the walkers work on this one
                const uint32_t nx = atomicAdd(ticket, 1u);
                S.tile_idx_next = nx;
                const uint64_t tn = tile_begin + (uint64_t)nx;
                if (tn < tile_end) {
                    const uint64_t ns = tn * (uint64_t)TB;
                    const uint32_t nbytes = (uint32_t)min((uint64_t)TB, P.n - ns) & ~15u;
                    if (nbytes) bulk_prefetch_l2(P.bytes + ns, nbytes);
                }
#endif
            }
            __syncwarp();
#if NTG_STATS
            st_mark = clock64();
            st_d += st_mark - st_ph0;
#endif
            const bool had_pending = S.pend_valid != 0;
            __syncwarp();                                       // (every lane has read the flag before lane 0 sets it again below)
            if (had_pending) {
                resolve_pending(P, S, epoch, lane, acc, slow);
#if NTG_STATS
                st_nlb++;
#endif
            }
            if (defer) {
                if (lane == 0) {
                    S.pend_agg = agg; S.pend_t = t; S.pend_guess = guess; S.pend_cs = Cs; S.pend_avail = avail;
                    S.pend_line0 = line0_starts_here ? 1u : 0u;
                    for (uint32_t j = 0; j < 4; j++) S.pend_nl4[j] = j < Cs ? S.nl[j] : 0u;
                    S.pend_valid = 1;
                }
            } else {
#if NTG_STATS == 3
                pre = tile_prefix(P, t, epoch, lane, fasta, slow, st_dbg);
#else
                pre = tile_prefix(P, t, epoch, lane, fasta, slow);
#endif
                if (lane == 0) {
                    publish_inclusive(P, t, epoch, pre, agg, fasta);
                    S.prefix = pre;
                    __threadfence_block();
                    S.prefix_seq = my_seq + 1;
                }
            }
#if NTG_STATS
            st_lb += clock64() - st_mark; if (!defer && t > 0) st_nlb++;
#endif
            __syncwarp();
        }
        bool have_pre = is_coord && !defer;
        if (!spec) { tile_sync(); pre = S.prefix; have_pre = true; }      // everyone needs the prefix before going on
#if NTG_STATS == 2
        if (fasta) { const long long c = clock64(); st_wait += c - st_ph; st_ph = c; }     // P2b + look-back wait (FASTA: every thread waits)
#endif
        // previous newline (global position, NONE if none) `back` newlines before newline i of this tile (back >= 1)
        auto prev_nl = [&](uint32_t i, uint32_t back) -> uint64_t {
            if (i >= back) return tile_start + S.nl[i - back];
            const uint32_t r = back - i - 1;
            return r < 4 ? pre.last[r] : NONE;
        };
        auto cr_before = [&](uint64_t q, uint64_t prevq) -> uint32_t {      // trim_cr on the line (prevq, q)
            const uint64_t ls = prevq == NONE ? 0 : prevq + 1;
            if (q <= ls) return 0u;
            if (q - 1 + halo < tile_start && q - 1 < P.gmin) { slow |= FLAG_HALO_OVERFLOW; return 0u; }   // (streamed window no longer holds it)
            return byte_at(P, sb, tile_start, halo, q - 1) == '\r' ? 1u : 0u;
        };
        // warm-up bound of a fragment of line i: the line start (FASTQ) / the sequence-region start (FASTA)
        auto fastq_bound = [&](uint32_t i, int a, int& lo, bool& lo_exact) {
            if (i > 0 || line0_starts_here) { lo = line_start_rel(i); lo_exact = true; (void)a; }
            else if (!have_pre) halo_line_start(sb, halo, lo, lo_exact);
            else {
                const uint64_t p1 = pre.last[0];
                const int64_t ls = (p1 == NONE ? 0 : (int64_t)p1 + 1) - (int64_t)tile_start;
                lo_exact = ls >= -(int64_t)halo;
                lo = lo_exact ? (int)ls : -(int)halo;
            }
        };
        auto fasta_bound = [&](uint32_t i, int a, int& lo, bool& lo_exact) {
            const uint32_t rs = S.rstart[i];
            if (rs) { lo = (int)rs - HALO; lo_exact = true; }
            else if (pre.hdr == NONE || pre.hdr == INHDR) { lo = a; lo_exact = true; }     // (only before any header: malformed)
            else {
                const int64_t ls = (int64_t)pre.hdr + 1 - (int64_t)tile_start;
                lo_exact = ls >= -(int64_t)halo;
                lo = lo_exact ? (int)ls : -(int)halo;
            }
        };

        if (!fasta) {
            // ---------------------------------------------------------------------------- FASTQ
            uint32_t ord0;
            if (have_pre) ord0 = (uint32_t)(pre.count & 3);
            else if (guess != 4) ord0 = guess;
            else {                                                      // no unique local evidence: wait for the coordinator
                while (S.prefix_seq <= my_seq) __nanosleep(32);
                __threadfence_block();
                pre = S.prefix; have_pre = true;
                ord0 = (uint32_t)(pre.count & 3);
            }
            if (spec && is_coord && guess != 4 && guess != ord0) slow |= FLAG_SPEC_MISS;
            // (A) line events: start bytes, n_bases, record completion (fastq.rs:240-285).
            auto line_events = [&](uint32_t i) {
                const uint32_t role = (ord0 + i) & 3;                     // 0 header, 1 sequence, 2 separator, 3 quality
                const int s = line_start_rel(i);
                const bool starts = (i > 0 || line0_starts_here) && (uint32_t)s < avail;
                auto rec_start = [&]() -> uint64_t { const uint64_t p = prev_nl(i, role + 1); return p == NONE ? 0 : p + 1; };
                if (starts) {
                    if (role == 0 && sb[s] != '@') note_parse_error(P.err_key, slow, tile_start + (uint64_t)s, 0);
                    if (role == 2 && sb[s] != '+') note_parse_error(P.err_key, slow, rec_start(), 1);
                }
                if (i < Cs) {                                             // the line ends in this tile at newline q
                    const uint64_t q = tile_start + S.nl[i];
                    if (role == 1) {
                        const uint64_t p1 = prev_nl(i, 1);
                        const uint64_t ls = p1 == NONE ? 0 : p1 + 1;
                        acc.n_bases += (q - ls) - cr_before(q, p1);
                    } else if (role == 3) {
                        const uint64_t q2 = prev_nl(i, 1), q1 = prev_nl(i, 2), q0 = prev_nl(i, 3);
                        if (q2 == NONE || q1 == NONE || q0 == NONE) note_parse_error(P.err_key, slow, 0, 0);   // inconsistent state
                        else {
                            const uint64_t seq_len = (q1 - q0 - 1) - cr_before(q1, q0);
                            const uint64_t qual_len = (q - q2 - 1) - cr_before(q, q2);
                            if (seq_len != qual_len) note_parse_error(P.err_key, slow, rec_start(), 2);
                            acc.n_records++;
                        }
                    }
                }
            };
            // Without speculation every thread has the prefix and takes lines tid, tid+NT, ...  With speculation the
            // coordinator warp (which has the prefix) takes the lines that may need it — the first four of the tile —
            // and each walker takes the lines >= 4 of "its" record, whose previous newlines are all in the tile's list.
            if (!spec) { for (uint32_t i = tid; i <= Cs; i += NWK) line_events(i); }
            else if (is_coord && have_pre) { for (uint32_t i = lane; i <= Cs && i < 4; i += 32) line_events(i); }   // (deferred: resolve_pending)
            // (B) sequence lines only: walker thread j takes the j-th role-1 line of the tile (every 4th line)
            const uint32_t i_first = (1u - ord0) & 3u;
#if NTG_STATS
            st_mark = clock64();
#endif
            if (!is_coord) {
                for (uint32_t i = i_first + 4u * tid; i <= Cs + 1; i += 4u * NTW) {
                    if (spec) {                                             // events of this record's lines (header .. quality)
                        for (uint32_t e = (i ? i - 1 : 0); e <= i + 2 && e <= Cs; e++) if (e >= 4) line_events(e);
                    }
                    if (i > Cs) continue;
                    const int a = line_start_rel(i), b = line_end_rel(i);
                    if (b <= a) continue;
                    if (b - a > SEG) { const uint32_t li = atomicAdd(&S.n_long, 1u); if (li < LONGMAX) S.long_line[li] = i; continue; }
                    int lo; bool lo_exact;
                    fastq_bound(i, a, lo, lo_exact);
                    run_item<KW, MINI, W, FK, FM>(sb, S.lut, S.rins, S.comb, a, b, lo, lo_exact, P, acc, false, slow, mode);
                }
            }
        } else {
            // ---------------------------------------------------------------------------- FASTA
            const bool line0_hdr_cont = !line0_starts_here && pre.hdr == INHDR;
            auto is_header = [&](uint32_t i) -> bool {
                if (i == 0 && !line0_starts_here) return line0_hdr_cont;
                const int s = line_start_rel(i);
                return (uint32_t)s < avail && sb[s] == '>';
            };
            // exclusive max-scan over lines of (end of header line + 1 + HALO): start of the sequence region
            const uint32_t per = (Cs + 1 + NWK - 1) / NWK;
            const uint32_t i0 = min(tid * per, Cs + 1), i1 = min(i0 + per, Cs + 1);
            uint32_t lm = 0;
            for (uint32_t i = i0; i < i1; i++) if (i < Cs && is_header(i)) lm = max(lm, (uint32_t)S.nl[i] + 1 + HALO);
            const uint32_t incl = block_incl_max(lm, S.warp_tmp[3]);
            uint32_t run = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) run = 0;
            __shared__ uint32_t s_prev[NT / 32 + 1];
            if (lane == 31) s_prev[tid >> 5] = incl;
            tile_sync();
            if (lane == 0 && tid > 0) run = s_prev[(tid >> 5) - 1];
            for (uint32_t i = i0; i < i1; i++) {
                S.rstart[i] = run;
                if (i < Cs && is_header(i)) run = max(run, (uint32_t)S.nl[i] + 1 + HALO);
            }
            tile_sync();
            for (uint32_t i = tid; i <= Cs; i += NWK) {
                if (is_header(i)) continue;
                const int a = line_start_rel(i), b = line_end_rel(i);
                if (b <= a) continue;
                if (b - a > SEG) { const uint32_t li = atomicAdd(&S.n_long, 1u); if (li < LONGMAX) S.long_line[li] = i; continue; }
                int lo; bool lo_exact;
                fasta_bound(i, a, lo, lo_exact);
                run_item<KW, MINI, W, FK, FM>(sb, S.lut, S.rins, S.comb, a, b, lo, lo_exact, P, acc, true, slow, mode);
            }
        }
#if NTG_STATS
        const long long st_done = clock64();
#endif
#if !NTG_EARLY_TICKET
        // The next tile of this CTA (read after a barrier).  A ticket obliges its holder to publish that tile's aggregate soon:
        // every later tile's look-back waits for it.  So the ticket is taken by the LAST warp to finish the walk: the warps of a
        // CTA finish far apart (the schedulers favour the oldest warp), and thread 0 taking it after its own items made the
        // successors of the next tile poll through most of this tile's walk (NTG_STATS == 3 on the 10 kbp FASTA shape: 12 000
        // cycles from loop top to aggregate, yet 70 000 of 153 000 cycles per tile spent polling for predecessors' aggregates).
        // A tile made of long lines does its walking in the piece loop below and takes the ticket after that loop.
        const bool late_claim = (uint64_t)avail > (uint64_t)(Cs + 1) * SEG;      // mean line longer than a piece: some line is long
        // (speculative FASTQ defers its look-backs by a tile, which hides that wait: thread 0 takes the ticket — 1 % faster there)
        if (spec) { if (tid == 0 && !late_claim) S.tile_idx_next = atomicAdd(ticket, 1u); }
        else if (!late_claim) claim_when_last(S, ticket, lane);
#else
        const bool late_claim = false;
#endif
        tile_sync();
#if NTG_STATS
        if (!fasta && !is_coord) { st_walk += st_done - st_mark; st_wait += clock64() - st_done; }
#endif
        // ---- long lines: SEG-byte pieces shared by the whole CTA
        const uint32_t n_long = min(S.n_long, (uint32_t)LONGMAX);
        // piece length: about one piece per thread when the tile is made of long lines (at least 128 B, at most SEG)
        // (every long line ends with a partial piece: leave one thread per line for it, or a few threads walk a second round)
        const uint32_t pthreads = n_long < NWK / 2 ? NWK - n_long : NWK / 2;
        const int PSEG = max(128, min(SEG, (int)(((avail + pthreads - 1) / pthreads + 15) & ~15u)));
        if (n_long) {
            if (tid == 0) {
                // (order of long_line[] is arbitrary: atomics) -> prefix of piece counts
                uint32_t sacc = 0;
                for (uint32_t j = 0; j < n_long; j++) {
                    const uint32_t i = S.long_line[j];
                    S.long_pref[j] = sacc; sacc += (uint32_t)(line_end_rel(i) - line_start_rel(i) + PSEG - 1) / PSEG;
                }
                S.long_pref[n_long] = sacc;
            }
            tile_sync();
            const uint32_t n_pieces = S.long_pref[n_long];
            for (uint32_t pc = tid; pc < n_pieces; pc += NWK) {
                uint32_t j = 0;
                while (j + 1 < n_long && S.long_pref[j + 1] <= pc) j++;
                const uint32_t i = S.long_line[j];
                const int la = line_start_rel(i), lb = line_end_rel(i);
                const int a = la + (int)(pc - S.long_pref[j]) * PSEG, b = min(a + PSEG, lb);
                int lo; bool lo_exact;
                if (!fasta) fastq_bound(i, la, lo, lo_exact); else fasta_bound(i, la, lo, lo_exact);
                run_item<KW, MINI, W, FK, FM>(sb, S.lut, S.rins, S.comb, a, b, lo, lo_exact, P, acc, fasta, slow, mode);
            }
        }
        if (n_long || late_claim) {                        // (no long lines: nothing read the tile since the barrier above, and
#if !NTG_EARLY_TICKET
            if (spec) { if (tid == 0 && late_claim) S.tile_idx_next = atomicAdd(ticket, 1u); }
            else if (late_claim) claim_when_last(S, ticket, lane);
#endif
            tile_sync();                                   //  the reset of S.n_long at the next tile stores the value it holds)
        }
        next_ticket = S.tile_idx_next;
#if NTG_STATS == 2
        if (fasta) st_walk += clock64() - st_ph;            // region scan + short lines + long-line pieces
#endif
        my_seq++;
    }

    if (is_coord && S.pend_valid) resolve_pending(P, S, epoch, lane, acc, slow);      // the last deferred tile of this CTA

#if NTG_STATS
    {
        const unsigned long long life = (unsigned long long)(clock64() - st_t0);
        if (tid == 0) { atomicAdd(&P.tallies[9], life); atomicMax(&P.tallies[12], life);
                        atomicAdd(&P.tallies[14], (unsigned long long)st_wait); atomicAdd(&P.tallies[15], (unsigned long long)st_walk); }
#if NTG_STATS == 3
        if (tid == NTW) { atomicAdd(&P.tallies[10], (unsigned long long)st_lb); atomicAdd(&P.tallies[11], (unsigned long long)st_nlb);
                          atomicAdd(&P.tallies[14], (unsigned long long)st_d); atomicAdd(&P.tallies[15], (unsigned long long)st_dbg[0]); atomicAdd(&P.tallies[6], (unsigned long long)st_dbg[1]); }
#elif NTG_STATS == 2
        if (tid == 0) { atomicAdd(&P.tallies[10], (unsigned long long)st_p0); atomicAdd(&P.tallies[11], (unsigned long long)st_p1); atomicAdd(&P.tallies[6], (unsigned long long)st_p2); }   // ([6] = n_query: unused without a query)
#else
        if (tid == NTW) { atomicAdd(&P.tallies[10], (unsigned long long)st_lb); atomicAdd(&P.tallies[11], (unsigned long long)st_nlb); }
#endif
    }
#endif
    // ---- P4: block reduction of the register tallies, 9 atomics per CTA
    uint64_t v[9] = {acc.n_records, acc.n_bases, acc.n_kmers, acc.n_not_rc, acc.ksum_lo, acc.ksum_hi, acc.n_query, acc.n_mini, acc.msum};
#pragma unroll
    for (int q = 0; q < 9; q++) {
#pragma unroll
        for (int d = 16; d; d >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], d);
        if (lane == 0) S.red[tid >> 5][q] = v[q];
    }
    slow = __reduce_or_sync(0xffffffffu, slow);
    if (lane == 0 && slow) atomicOr(P.flags, slow);
    __syncthreads();
    if (tid < 9) {
        uint64_t sres = 0;
        for (int wi = 0; wi < NT / 32; wi++) sres += S.red[wi][tid];
        if (sres) atomicAdd(&P.tallies[tid], (unsigned long long)sres);
    }
}

// ---- end-of-stream rules, one thread (fastq.rs:337-356, fasta.rs:200-216,348-356) ----------------
__global__ void k_finalize(const Params P) {
    const SState st = *P.final_state;
    uint32_t err = 0, slow = 0;
    auto byte = [&](uint64_t p) -> uint8_t {                        // (a streamed window holds the stream's tail only)
        if (p < P.gmin || p >= P.n) { slow |= FLAG_HALO_OVERFLOW; return 0; }
        return P.bytes[p];
    };
    if (P.format == NTG_FMT_FASTQ) {
        const uint32_t r = (uint32_t)(st.count & 3);
        const bool have_complete = st.count >= 4;
        uint64_t start = 0;
        if (have_complete) start = st.last[r] + 1;                 // behind the 4th newline of the last complete record
        auto trim = [&](uint64_t b, uint64_t e) { return (e > b && byte(e - 1) == '\r') ? e - 1 : e; };
        if (r == 3) {                                              // last record without trailing newline
            const uint64_t seq = st.last[2] + 1, sep = st.last[1] + 1, qual = st.last[0] + 1, end = P.n;
            if (byte(start) != '@' || byte(sep) != '+' || trim(seq, sep - 1) - seq != trim(qual, end) - qual) err = 1;
            else atomicAdd(&P.tallies[0], 1ull);
        } else {
            uint64_t ls = start;
            for (uint32_t i = 0; i <= r; i++) {                     // leftover must be empty / "\r" lines only
                const uint64_t le = i < r ? st.last[r - 1 - i] : P.n;
                const uint64_t len = le - ls;
                if (len > 1 || (len == 1 && byte(ls) != '\r')) err = 1;
                ls = le + 1;
            }
        }
        if (err) note_parse_error(P.err_key, slow, start, 3);       // the host replays [0, start) and classifies the record at `start`
    } else {
        // FASTA: the last record needs a pushed newline (one that is not the final byte); without one it is an UnexpectedEnd
        // (fasta.rs:205-213,348-356) and is not delivered.  Its header line is the last line of the stream: nothing of it was tallied.
        const bool bad = st.hdr == INHDR || st.hdr == NONE || st.hdr == P.n - 1;
        const uint64_t n_starts = P.fa_totals[0], count = P.fa_totals[1];           // sums over every tile of the pass
        P.tallies[0] = n_starts - (bad && n_starts ? 1 : 0);
        if (bad) {
            slow |= FLAG_PARSE_ERROR;
            P.fin[0] = NTG_EUNEXPECTED_END;
            P.fin[1] = 1 + count - ((st.hdr != INHDR && st.hdr != NONE) ? 1 : 0);      // line of that header: newlines before it + 1
            P.fin[2] = n_starts ? n_starts - 1 : 0;                                    // its record index
        }
    }
    // (the resident entry point caches the sniffed format per buffer: a buffer whose content changed format is caught here)
    if (P.gmin == 0 && P.n && P.bytes[0] != (P.format == NTG_FMT_FASTQ ? '@' : '>')) slow |= FLAG_FORMAT;
    if (slow) atomicOr(P.flags, slow);
    if (P.reduce_buf) {
#pragma unroll 1
        for (int i = 0; i < 9; i++) P.reduce_buf[i] = P.tallies[i];
        P.reduce_buf[9] = *P.flags ? 1ull : 0ull;                  // ranks whose result needs the host (error replay, exact path)
        for (int i = 10; i < 16; i++) P.reduce_buf[i] = 0;
    }
}

// record-owned passes with NTG_TALLY_ALLREDUCE: tallies -> NCCL send buffer ([9] = this rank needs the host: flags, or a tail)
__global__ void k_reduce_copy(const unsigned long long* tallies, const uint32_t* flags, const unsigned long long* next, uint64_t n_vis,
                              unsigned long long* reduce_buf) {
    const uint32_t i = threadIdx.x;
    if (i < 9) reduce_buf[i] = tallies[i];
    else if (i == 9) reduce_buf[9] = (*flags || *next < n_vis) ? 1ull : 0ull;
    else if (i < 16) reduce_buf[i] = 0;
}

// ---- exact fallback: one thread per parsed record walks its raw_seq from global memory ---------
template <int KW, bool MINI>
__global__ void __launch_bounds__(128) k_tally_records(const Params P, const ntg_record* __restrict__ recs, uint64_t n_recs) {
    __shared__ uint8_t lut[256];
    __shared__ uint64_t red[4][9];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = class_of(i);
    __syncthreads();
    Acc acc; uint32_t slow = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_recs; r += (uint64_t)gridDim.x * blockDim.x) {
        const ntg_record rec = recs[r];
        acc.n_records++; acc.n_bases += rec.num_bases;
        const uint8_t* base = P.bytes + rec.seq_b;
        const uint64_t len = rec.seq_e - rec.seq_b;
        for (uint64_t o = 0; o < len; o += 0x40000000ull) {          // int-sized windows (records beyond 1 GiB)
            const uint64_t rem = len - o;
            const uint64_t wlen = rem < 0x40000000ull ? rem : 0x40000000ull;
            const int lo = o ? -0x100000 : 0;                         // warm-up room inside the same record
            const int ws = find_ws(base + o, lut, 0, lo, true, (int)P.k, slow);
            walk<KW, MINI, 0>(base + o, lut, ws, 0, (int)wlen, P, acc, false);
        }
    }
    uint64_t v[9] = {acc.n_records, acc.n_bases, acc.n_kmers, acc.n_not_rc, acc.ksum_lo, acc.ksum_hi, acc.n_query, acc.n_mini, acc.msum};
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 9; q++) {
#pragma unroll
        for (int d = 16; d; d >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], d);
        if (lane == 0) red[threadIdx.x >> 5][q] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        uint64_t s = 0;
        for (int wi = 0; wi < (int)(blockDim.x >> 5); wi++) s += red[wi][threadIdx.x];
        if (s) atomicAdd(&P.tallies[threadIdx.x], (unsigned long long)s);
    }
}
}  // namespace fused
