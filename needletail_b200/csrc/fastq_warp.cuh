// fastq_warp.cuh — the short-read FASTQ fast path: RECORD-OWNED chunks, one warp per chunk, no CTA barriers (unity build).
//
// fused::k_fused synchronises 320 threads four times per 84 KiB tile and its INT pipe idles while a tile is loaded, scanned
// and listed (≈ 40 % of a CTA's time).  Here every WARP works alone: it claims a chunk of ≈ 10 KB, brings it
// (plus 2 KiB of slack and 16 bytes of back halo) into its own shared-memory buffer with one TMA bulk copy, lists the newlines,
// and owns the records that START in the chunk — a record that crosses the chunk end is finished from the slack.  A record is
// therefore always whole inside one warp's buffer: validation (fastq.rs:240-285) is local, the walk of its sequence line needs
// no warm-up, and nothing is looked up across chunks while the chunk is resident.
//
// The one global dependency — which line starts are record starts (newline ordinal mod 4; quality lines may start with '@' or
// '+') — is SPECULATED from local evidence (the same rule as fused::guess_phase) and VERIFIED afterwards: every chunk leaves the
// number of line starts it holds (mod 4) and its guess; k_verify takes the prefix over the chunks of the launch (a launch
// starts at a record start, so the prefix starts at 0), raises FLAG_SPEC_MISS on a wrong guess and hands chunks without unique
// evidence to a second launch of the same kernel with their true phase (fix-up).
//
// What the fast path does not do is left to the callers in fused_host.cuh, which fall back to fused::k_fused for the pass:
// wrong guesses, records longer than the slack, newline-dense chunks.  The end of the stream is not interpreted here either:
// the first record that is not complete inside the visible bytes is reported (`carry`); at the end of a stream the host runs
// the exact record-table path on that tail (at most one record and blank lines), inside a stream the next launch starts there.
#pragma once
#include "fused.cuh"

namespace fqw {
constexpr int WARPS = 4;                       // warps per CTA, each on its own chunk
constexpr int NT = WARPS * 32;
constexpr int CTAS_PER_SM = 4;
constexpr int CBMAX = 10240;                   // chunk capacity (bytes, multiple of 256)
constexpr int SLACK = 2048;                    // a record may extend this far behind its chunk
constexpr int HALO = 16;                       // back halo: is the first byte of the chunk a line start?
constexpr int WIN = CBMAX + SLACK;             // bytes scanned per chunk
constexpr int ROWS = WIN / fused::ROWB;        // 48 rows of 256 B: at most two per lane
constexpr int NLW = 318;                       // newline capacity per chunk window
static_assert(WIN % fused::ROWB == 0 && ROWS <= 64, "two scan rounds per lane");

enum : uint32_t { FLAG_LONG = 64, FLAG_FQ_NL = 128 };      // (beside fused::FLAG_*) record longer than the slack / newline-dense chunk
enum : uint8_t { INFO_DONE = 0x80, INFO_UNRESOLVED = 0x40 };

struct Params {
    fused::Params W;                           // walker parameters (k, m, w, query, spectrum binding), tallies / flags / err_key pointers
    const uint8_t* bytes;                      // virtual base: stream byte p is bytes[p]
    const unsigned long long* start;           // stream position of the launch's first record (a record start), device word
    uint64_t n_vis;                            // stream bytes visible to this launch
    uint32_t cb;                               // chunk bytes in use (multiple of 16, <= CBMAX)
    uint32_t frags;                            // fragments a sequence line is cut into (1, 2 or 4): items = records x frags
    uint8_t* info;                             // per chunk: line starts in the chunk mod 4 | guess << 2 | INFO_*
    uint32_t* ticket;
    unsigned long long* carry;                 // min over the starts of records that are not complete inside the visible bytes
    // fix-up launches
    const uint32_t* fix_list; const uint32_t* fix_count; const uint8_t* fix_phase; uint32_t fix_cap;
};

struct __align__(16) WarpSmem {
    uint8_t halo[HALO];
    uint8_t win[WIN];
    uint16_t nl[NLW + 2];
    uint64_t bar;
};
struct __align__(16) Smem {
    WarpSmem w[WARPS];
    uint8_t lut[256];
    uint32_t rins[256];
    uint32_t comb[256];
    uint64_t red[WARPS][9];
};

// Newlines of one 256-byte row, as a loop (fused::scan_row is the same test fully unrolled: 8 KB of code per call site, and
// with every warp of an SM in a different phase of its chunk the instruction cache was this kernel's largest stall).  The
// lane reads 16-byte column (j + lane) & 15 at step j (bank-conflict free for rows 256 B apart) and sets the mask bits of
// that column directly.
__device__ __forceinline__ void scan_row_loop(const uint32_t* __restrict__ row, uint32_t lane, uint32_t& cnt, uint64_t& mask) {
    const uint4* __restrict__ row4 = reinterpret_cast<const uint4*>(row);
    uint32_t acc = 0; uint64_t m = 0;
    auto test = [](uint32_t w) -> uint32_t {
        const uint32_t x = w ^ 0x0A0A0A0Au;
        return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;      // 0x80 in every byte that is exactly '\n'
    };
#pragma unroll 1
    for (uint32_t j = 0; j < 16; j++) {
        const uint32_t col = (j + lane) & 15u;
        const uint4 v = row4[col];
        const uint32_t z0 = test(v.x), z1 = test(v.y), z2 = test(v.z), z3 = test(v.w);
        acc += (z0 >> 7) + (z1 >> 7) + (z2 >> 7) + (z3 >> 7);                // <= 64 per byte lane over the row
        const uint32_t nib = (z0 ? 1u : 0u) | (z1 ? 2u : 0u) | (z2 ? 4u : 0u) | (z3 ? 8u : 0u);
        m |= (uint64_t)nib << (4u * col);
    }
    const uint32_t t = (acc & 0x00FF00FFu) + ((acc >> 8) & 0x00FF00FFu);
    cnt = (t & 0xFFFFu) + (t >> 16);
    mask = m;
}

// One fragment [ap, bp) of the sequence line that starts at a (ap == a for the first fragment: no warm-up; later fragments
// warm up over the k-1 bases in front of them, inside their own line).  fused::run_item without the warm-up-from-codes variant
// (a line start has nothing in front of it), i.e. with one copy of the clean walker instead of two.
template <int KW, bool MINI, int W, int FK, int FM>
__device__ __forceinline__ void run_fragment(const uint8_t* sb, const uint8_t* lut, const uint32_t* rins, const uint32_t* comb, int a, int ap, int bp,
                                             const fused::Params& P, fused::Acc& acc, uint32_t& slow, uint32_t& mode) {
    int b = bp;
    if (b > ap && sb[b - 1] == '\r') b--;                  // a trailing '\r' is deleted by normalize: nothing to walk
    if (b <= ap) return;
    const int ws = ap > a ? fused::find_ws(sb, lut, ap, a, true, (int)P.k, slow) : a;
    constexpr bool ONE = FK >= 21 && FK <= 31;
    if (FK > 32) {
        if (!__any_sync(__activemask(), mode != 0u) && fused::walk_clean2<(FK > 32 ? FK : 51)>(sb, comb, ws, b, acc)) return;
        mode = 1u;
        uint32_t any_bad = 0;
        for (int q = ws; q < b; q++) any_bad |= lut[sb[q]];
        if (any_bad <= 3u) mode = 0u;
        fused::walk<KW, MINI, W>(sb, lut, ws, ap, b, P, acc, false);
    } else if (ONE) {
        constexpr int CK = ONE ? FK : 21, CM = ONE ? FM : 0;
        bool done = false;
        if (!__any_sync(__activemask(), mode != 0u)) done = fused::walk_clean<CK, CM, false>(sb, comb, ws, b, 0, acc);
        if (!done) {
            uint32_t seen = 0x80u;
            done = fused::walk_fast_cold<CK, CM>(sb, lut, rins, ws, b, acc, &seen);
            mode = seen > 3u ? 1u : 0u;
        }
        if (!done) fused::walk_slow<KW, MINI, W>(sb, lut, ws, ap, b, P, acc, false);
    } else {
        fused::walk<KW, MINI, W>(sb, lut, ws, ap, b, P, acc, false);
    }
}

template <int KW, bool MINI, int W, int FK, int FM>
__global__ void __launch_bounds__(NT, CTAS_PER_SM) k_records(const Params P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t lane = tid & 31, wid = tid >> 5;
    WarpSmem& B = S.w[wid];
    for (int i = tid; i < 256; i += NT) {
        const uint8_t c = fused::class_of(i);
        S.lut[i] = c;
        const uint32_t ri = (FK > 32) ? ((3u - (c & 3u)) << fused::Clean2Shape<(FK > 32 ? FK : 51)>::SH)
                                      : (FK >= 17) ? ((3u - (c & 3u)) << (2 * ((FK >= 17 && FK <= 32 ? FK : 17) - 1) - 32)) : 0u;
        S.rins[i] = ri;
        S.comb[i] = ri | c;
    }
    if (lane == 0) { fused::mbar_init(&B.bar, 1); fused::fence_mbar_init(); }
    __syncthreads();
    const uint64_t S0 = *P.start;
    const uint64_t A = S0 & ~15ull;                                   // chunk grid origin (TMA sources are 16-byte aligned)
    const uint32_t CB = P.cb;
    const uint32_t WINB = min((uint32_t)WIN, (CB + (uint32_t)SLACK + 255u) & ~255u);   // bytes of window in use: chunk + slack, whole rows
    const uint8_t* sb = B.win;
    uint32_t parity = 0, slow = 0, mode = 0;
    fused::Acc acc;
    const bool FIX = P.fix_list != nullptr;                           // fix-up launch: chunks and their true phases come from k_verify
    const uint32_t n_fix = FIX ? min(*P.fix_count, P.fix_cap) : 0u;
    // Chunks are independent, so a warp claims its NEXT chunk while it starts on the current one and asks for it to be pulled into
    // L2 (cp.async.bulk.prefetch.L2): the bulk copy that later stages it then pays an L2 round trip instead of an HBM one.
    uint32_t c_next = 0;
    if (lane == 0) c_next = atomicAdd(P.ticket, 1u);
    c_next = __shfl_sync(0xffffffffu, c_next, 0);
    for (;;) {
        uint32_t c = c_next;
        uint32_t given = 4;
        if (FIX) { if (c >= n_fix) break; given = P.fix_phase[c]; c = P.fix_list[c]; }
        const uint64_t lo = A + (uint64_t)c * CB;
        if (lo >= P.n_vis) break;
        if (lane == 0) {
            c_next = atomicAdd(P.ticket, 1u);
            if (!FIX) {
                const uint64_t nlo = A + (uint64_t)c_next * CB;
                if (nlo < P.n_vis) { const uint32_t nb = (uint32_t)min((uint64_t)WINB, P.n_vis - nlo) & ~15u; if (nb) fused::bulk_prefetch_l2(P.bytes + nlo, nb); }
            }
        }
        c_next = __shfl_sync(0xffffffffu, c_next, 0);
        const uint32_t avail = (uint32_t)min((uint64_t)WINB, P.n_vis - lo);
        const uint32_t halo = lo ? HALO : 0;
        const uint32_t bulk = avail & ~15u;
        // ---- the chunk window (+ halo) with one bulk async copy; the unaligned tail and the bytes behind the stream by hand
        if (lane == 0 && halo + bulk) {
            fused::mbar_expect_tx(&B.bar, halo + bulk);
            fused::bulk_g2s(B.halo + (HALO - halo), P.bytes + lo - halo, halo + bulk, &B.bar);
        }
        for (uint32_t i = bulk + lane; i < WINB; i += 32) B.win[i] = i < avail ? P.bytes[lo + i] : 0;
        if (halo + bulk) { fused::mbar_wait(&B.bar, parity); parity ^= 1; }
        __syncwarp();
        // ---- newlines of the window: rows lane and lane + 32, ordered list by two warp scans
        uint32_t cnt[2]; uint64_t wm[2];
#pragma unroll 1
        for (int rd = 0; rd < 2; rd++) {
            cnt[rd] = 0; wm[rd] = 0;
            const uint32_t row = rd * 32 + lane;
            if (row < (uint32_t)ROWS && row * fused::ROWB < avail) scan_row_loop(reinterpret_cast<const uint32_t*>(B.win) + row * fused::ROWW, lane, cnt[rd], wm[rd]);
        }
        uint32_t inc = cnt[0] | (cnt[1] << 16);                      // (a row holds at most 256 newlines, a round at most 8192)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
        const uint32_t C0 = tot & 0xFFFFu, C = C0 + (tot >> 16);
        const uint32_t exc = inc - (cnt[0] | (cnt[1] << 16));
        if (C > (uint32_t)NLW) {                                     // newline-dense chunk: not this kernel's business
            slow |= FLAG_FQ_NL;
            if (lane == 0 && !FIX) P.info[c] = INFO_DONE;
            continue;
        }
#pragma unroll
        for (int rd = 0; rd < 2; rd++) {
            if (!cnt[rd]) continue;
            const uint32_t row = rd * 32 + lane;
            const uint32_t* rowp = reinterpret_cast<const uint32_t*>(B.win) + row * fused::ROWW;
            uint32_t o = rd == 0 ? (exc & 0xFFFFu) : C0 + (exc >> 16);
            uint64_t mk = wm[rd];
            while (mk) {
                const int jj = __ffsll((long long)mk) - 1;
                mk &= mk - 1;
                const uint32_t wv = rowp[jj];
#pragma unroll
                for (int bs = 0; bs < 4; bs++)
                    if (((wv >> (8 * bs)) & 0xFF) == '\n') B.nl[o++] = (uint16_t)(row * fused::ROWB + jj * 4 + bs);
            }
        }
        __syncwarp();
        // ---- line starts of the chunk.  f0: first byte that belongs to this launch (chunk 0 begins up to 15 bytes early);
        // u: newlines in front of it; fis: is f0 itself a line start.  LS[j] for j >= 1 (or 0 without fis) = nl[..] + 1 < CB.
        const uint32_t f0 = (c == 0) ? (uint32_t)(S0 - A) : 0u;
        uint32_t u = 0;
        if (f0) { for (uint32_t j = lane; j < C && j < 32; j += 32) u += B.nl[j] < f0 ? 1u : 0u; u = __reduce_add_sync(0xffffffffu, u); }
        const bool fis = (c == 0) || (lo == 0) || B.halo[HALO - 1] == '\n';
        uint32_t in_chunk = 0;                                        // newlines j >= u whose successor byte is a line start inside the chunk
        for (uint32_t j = u + lane; j < C; j += 32) in_chunk += (uint32_t)B.nl[j] + 1u < CB ? 1u : 0u;
        in_chunk = __reduce_add_sync(0xffffffffu, in_chunk);
        const uint32_t n_ls = (fis && f0 < CB ? 1u : 0u) + in_chunk;
        const uint32_t e_off = u + (fis ? 0u : 1u);                  // line start j ends at newline nl[e_off + j]
        auto line_start = [&](uint32_t j) -> uint32_t { return (fis && j == 0) ? f0 : (uint32_t)B.nl[e_off + j - 1] + 1u; };
        // ---- which line starts are record starts: g = ordinal mod 4 of LS[0].  Chunk 0 starts at a record start by definition;
        // elsewhere the unique g for which the first <= 8 line starts (slack included) have '@' at role 0 and '+' at role 2.
        uint32_t g;
        if (FIX) g = given;
        else if (c == 0) g = 0;
        else {
            uint32_t ok = 0xF; bool valid = false;
            if (lane < 8 && e_off + lane <= C) {                      // (LS[j] exists as long as its predecessor newline does)
                const bool have = (fis && lane == 0) || e_off + lane >= 1;
                const uint32_t s = have ? line_start(lane) : avail;
                if (s < avail) {
                    valid = true;
                    const uint8_t ch = sb[s];
                    ok &= ~((ch != '@' ? 1u : 0u) << ((0u - lane) & 3u));
                    ok &= ~((ch != '+' ? 1u : 0u) << ((2u - lane) & 3u));
                }
            }
            const uint32_t seen = (uint32_t)__popc(__ballot_sync(0xffffffffu, valid));
            ok = __reduce_and_sync(0xffffffffu, ok);
            g = (seen < 4 || __popc(ok) != 1) ? 4u : (uint32_t)__ffs((int)ok) - 1u;
        }
        if (!FIX && lane == 0) P.info[c] = (uint8_t)((n_ls & 3u) | (g << 2) | (g == 4 ? INFO_UNRESOLVED : INFO_DONE));
        if (g == 4) continue;                                         // resolved by the fix-up launch with the true phase
        // ---- items: record r = line starts j0 + 4r .. ; fragment `part` of its sequence line
        const uint32_t j0 = (4u - g) & 3u;
        const uint32_t n_rec = n_ls > j0 ? (n_ls - j0 + 3u) / 4u : 0u;
        const uint32_t frags = P.frags;
        for (uint32_t t = lane; t < n_rec * frags; t += 32) {
            const uint32_t r = t / frags, part = t - r * frags;
            const uint32_t j = j0 + 4u * r;
            const uint32_t e0i = e_off + j;
            const uint32_t p = line_start(j);
            if (e0i + 3 >= C) {
                // not complete inside the window: behind the visible bytes (the next launch / the host's tail pass starts here)
                // or longer than the slack
                if (part == 0) {
                    if (avail < WINB) atomicMin(P.carry, (unsigned long long)(lo + p));
                    else slow |= FLAG_LONG;
                }
                continue;
            }
            const uint32_t e0 = B.nl[e0i], e1 = B.nl[e0i + 1], e2 = B.nl[e0i + 2], e3 = B.nl[e0i + 3];
            const int a = (int)e0 + 1, b = (int)e1;
            if (part == 0) {
                // validate (fastq.rs:240-285): '@', '+', equal lengths after trim_cr
                auto cr = [&](uint32_t ls, uint32_t e) -> uint32_t { return (e > ls && sb[e - 1] == '\r') ? 1u : 0u; };
                if (sb[p] != '@') fused::note_parse_error(P.W.err_key, slow, lo + p, 0);
                if (sb[e1 + 1] != '+') fused::note_parse_error(P.W.err_key, slow, lo + p, 1);
                const uint32_t seq_len = (e1 - (e0 + 1)) - cr(e0 + 1, e1), qual_len = (e3 - (e2 + 1)) - cr(e2 + 1, e3);
                if (seq_len != qual_len) fused::note_parse_error(P.W.err_key, slow, lo + p, 2);
                acc.n_records++; acc.n_bases += seq_len;
            }
            const int len = b - a;
            const int ap = a + (int)(((int64_t)len * part) / frags), bp = a + (int)(((int64_t)len * (part + 1)) / frags);
            if (P.W.qmask) {
                // QualitySequence::quality_mask (sequence.rs:280-297), fused: the record's quality line sits in the same buffer, so
                // low-quality bases become 'N' in place before the walk (this lane owns the bytes; the two lanes of a split
                // line write the same values where their ranges overlap).  Only the bases proper: trim_cr'd lengths.
                const int nb = len - ((len > 0 && sb[b - 1] == '\r') ? 1 : 0);
                const int q0 = (int)e2 + 1 - a;                             // offset from a base to its quality byte
                const int from = max(a, ap - (int)P.W.k + 1), to = min(bp, a + nb);
                for (int q = from; q < to; q++) if (sb[q + q0] < P.W.qmask) B.win[q] = 'N';
            }
            run_fragment<KW, MINI, W, FK, FM>(sb, S.lut, S.rins, S.comb, a, ap, bp, P.W, acc, slow, mode);
        }
        __syncwarp();
    }
    // ---- tallies: warp reduction, one block reduction, 9 atomics per CTA
    uint64_t v[9] = {acc.n_records, acc.n_bases, acc.n_kmers, acc.n_not_rc, acc.ksum_lo, acc.ksum_hi, acc.n_query, acc.n_mini, acc.msum};
#pragma unroll
    for (int q = 0; q < 9; q++) {
#pragma unroll
        for (int d = 16; d; d >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], d);
        if (lane == 0) S.red[wid][q] = v[q];
    }
    slow = __reduce_or_sync(0xffffffffu, slow);
    if (lane == 0 && slow) atomicOr(P.W.flags, slow);
    __syncthreads();
    if (tid < 9) {
        uint64_t sres = 0;
        for (int wi = 0; wi < WARPS; wi++) sres += S.red[wi][tid];
        if (sres) atomicAdd(&P.W.tallies[tid], (unsigned long long)sres);
    }
}

// Prefix over the chunks of a launch: true phase of every chunk = line starts before it (mod 4).  Wrong guesses raise
// FLAG_SPEC_MISS, chunks without a guess go to the fix-up list with their true phase; the launch's carry becomes the start of the
// next launch (`next_start`), or n_vis when every record was complete.  One CTA.
constexpr int VT = 1024;
__global__ void __launch_bounds__(VT) k_verify(const uint8_t* __restrict__ info, const unsigned long long* start, uint32_t cb, uint32_t* fix_list,
                                               uint8_t* fix_phase, uint32_t* fix_count, uint32_t fix_cap, uint32_t* flags, const unsigned long long* carry,
                                               uint64_t n_vis, unsigned long long* next_start, unsigned long long* ctl_next) {
    __shared__ uint32_t s_sum[VT];
    const uint32_t tid = threadIdx.x;
    const uint64_t A = *start & ~15ull;
    const uint32_t n_chunks = n_vis > A ? (uint32_t)((n_vis - A + cb - 1) / cb) : 0u;
    // 16 chunks per load (the info array is zero-padded to whole vectors by the host)
    const uint4* __restrict__ info4 = reinterpret_cast<const uint4*>(info);
    const uint32_t nvec = (n_chunks + 15u) / 16u;
    const uint32_t per = (nvec + VT - 1) / VT;
    const uint32_t b = min(tid * per, nvec), e = min(b + per, nvec);
    auto s3 = [](uint32_t w) -> uint32_t { return ((w & 0x03030303u) * 0x01010101u) >> 24; };      // sum of the four 2-bit counts
    uint32_t sum = 0;
#pragma unroll 4
    for (uint32_t i = b; i < e; i++) { const uint4 x = info4[i]; sum += s3(x.x) + s3(x.y) + s3(x.z) + s3(x.w); }
    s_sum[tid] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < VT; d <<= 1) {                          // inclusive scan (Hillis-Steele; once per launch)
        const uint32_t t = tid >= d ? s_sum[tid - d] : 0u;
        __syncthreads();
        s_sum[tid] += t;
        __syncthreads();
    }
    uint32_t pre = s_sum[tid] - sum, bad = 0;
#pragma unroll 2
    for (uint32_t i = b; i < e; i++) {
        const uint4 x = info4[i];
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const uint32_t ci = i * 16u + q;
            const uint32_t v = (w[q >> 2] >> (8 * (q & 3))) & 0xFFu;
            if (ci < n_chunks) {
                const uint32_t g = (v >> 2) & 7u, truth = pre & 3u;
                if (v & INFO_UNRESOLVED) {
                    const uint32_t o = atomicAdd(fix_count, 1u);
                    if (o < fix_cap) { fix_list[o] = ci; fix_phase[o] = (uint8_t)truth; } else bad |= fused::FLAG_SPEC_MISS;
                } else if (!(v & INFO_DONE) || (g < 4 && g != truth)) bad |= fused::FLAG_SPEC_MISS;      // wrong guess (or a chunk nobody claimed)
            }
            pre += v & 3u;
        }
    }
    if (bad) atomicOr(flags, bad);
    if (tid == 0) {
        const unsigned long long cv = *carry;
        const unsigned long long nx = cv == ~0ull ? (unsigned long long)n_vis : cv;
        *next_start = nx; *ctl_next = nx;
    }
}
}  // namespace fqw
