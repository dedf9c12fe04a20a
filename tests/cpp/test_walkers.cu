// Host-side differential test of the fused kernel's row scanner (scan_row) and per-item walkers
// (needletail_b200/csrc/fused.cuh: find_ws, walk, walk_fast, walk_clean) against the oracle's tally_sequence (oracle/ntref.hpp).  The walkers are __host__ __device__:
// this file compiles the very same source for the CPU (no device code is launched) and checks, on random lines,
//   1. every walker that accepts an item returns exactly the oracle's tallies for it,
//   2. an item cut into fragments (each warmed up over its k-1 predecessors, as the kernel does at tile / piece
//      boundaries) tallies to the same totals as the whole item  — k-mers are owned by their last base.
//   nvcc -std=c++17 --expt-relaxed-constexpr -o /tmp/test_walkers tests/cpp/test_walkers.cu && /tmp/test_walkers
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../needletail_b200/csrc/fused.cuh"
#include "../../oracle/ntref.hpp"

using fused::Acc; using fused::Params;

static uint8_t g_cls[256];
static void build_cls() {
    for (int b = 0; b < 256; b++) {
        uint8_t one = (uint8_t)b; std::vector<uint8_t> o;
        ntref::normalize(&one, 1, false, o);
        if (o.empty()) g_cls[b] = (b == '\r' || b == '\n') ? 0x86 : 0x85;
        else g_cls[b] = o[0] == 'A' ? 0 : o[0] == 'C' ? 1 : o[0] == 'G' ? 2 : o[0] == 'T' ? 3 : 4;
    }
}
struct Luts { uint32_t rins[256], comb[256]; };
template <int K> static Luts make_luts() {
    Luts l;
    for (int b = 0; b < 256; b++) { l.rins[b] = (3u - (g_cls[b] & 3u)) << (2 * (K - 1) - 32); l.comb[b] = l.rins[b] | g_cls[b]; }
    return l;
}
static bool same(const Acc& a, const ntref::Tallies& t, bool mini) {
    return a.n_kmers == t.n_kmers && a.n_not_rc == t.n_not_rc && a.ksum_lo == t.kmer_sum_lo &&
           (!mini || (a.n_mini == t.n_minimizers && a.msum == t.minimizer_sum));
}

enum Which { CLEAN = 1, FAST = 2, GENERIC = 4, CLEANFP = 8 };   // CLEANFP: walk_clean_fp (the NTG_FP64_MIN experiment) in place of walk_clean
// the kernel's run_item chain on sb[a..b) with warm-up not below lo; `which` selects the walkers that may be used
template <int K, int M>
static bool item(const uint8_t* sb, const Luts& L, int a, int b, int lo, int which, Acc& acc, int* used = nullptr) {
    constexpr bool MINI = M > 0; constexpr int W = MINI ? K - M + 1 : 0;
    if (b > a && sb[b - 1] == '\r') b--;
    if (b <= a) return true;
    uint32_t slow = 0;
    const int ws = fused::find_ws(sb, g_cls, a, lo, true, K, slow);
    if ((which & CLEAN) && fused::walk_clean<K, M>(sb, L.comb, ws, b, acc)) { if (used) *used = CLEAN; return true; }
    if constexpr (M > 0 && 2 * M <= 42 && K - M >= 8 && K - M <= 16) {
        if ((which & CLEANFP) && fused::walk_clean_fp<K, M>(sb, L.comb, ws, b, acc)) { if (used) *used = CLEANFP; return true; }
    }
    fused::FastLuts FL{g_cls, L.rins, L.comb};
    if ((which & FAST) && fused::walk_fast<K, M>(sb, FL, ws, b, acc)) { if (used) *used = FAST; return true; }
    if (which & GENERIC) {
        Params P{}; P.k = K; P.m = M; P.w = MINI ? K - M + 1 : 0; P.has_query = 0;
        fused::walk<1, MINI, W>(sb, g_cls, ws, a, b, P, acc, false);
        if (used) *used = GENERIC;
        return true;
    }
    return false;
}

template <int K, int M>
static int run(std::mt19937_64& rng, int iters, const char* name) {
    const Luts L = make_luts<K>();
    int fails = 0; long used_cnt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<uint8_t> norm, rc;
    for (int it = 0; it < iters && fails < 5; it++) {
        const int flavour = (int)(rng() % 5);               // 0: clean, 1: a few non-ACGT, 2: + whitespace, 3: dirty, 4: clean with T runs
        const int n = (int)(rng() % (it % 7 == 0 ? 700 : 220));
        std::string s(n, 'A');
        for (auto& c : s) {
            const uint64_t r = rng();
            c = "ACGTacgt"[r & 7];
            const unsigned roll = (unsigned)((r >> 8) % 1000);
            if (flavour >= 1 && roll < (flavour == 3 ? 80u : 8u)) c = "NnRYKM-.~UuXx*"[(r >> 20) % 14];
            if (flavour >= 2 && roll >= 990) c = " \t\r"[(r >> 30) % 3];
        }
        if (flavour == 4)                                     // runs of T: RC_k(x) < x happens (walk_clean's rare case)
            for (int q = 0; q < 3 && n > 0; q++) { const int at = (int)(rng() % (unsigned)n), len = 4 + (int)(rng() % 30); for (int j = at; j < n && j < at + len; j++) s[j] = (rng() & 1) ? 'T' : 't'; }
        // (a line never contains '\n'; the kernel strips one trailing '\r' itself)
        const uint8_t* sb = (const uint8_t*)s.data();
        ntref::Tallies t;
        ntref::tally_sequence(sb, (size_t)n, K, M, false, nullptr, t, norm, rc);
        // 1. whole item through every admissible chain
        for (int which : std::vector<int>{CLEAN | FAST | GENERIC, CLEANFP | FAST | GENERIC, FAST | GENERIC, GENERIC}) {
            Acc acc; int used = 0;
            item<K, M>(sb, L, 0, n, 0, which, acc, &used);
            if (which == (CLEAN | FAST | GENERIC) || used == CLEANFP) used_cnt[used]++;
            if (!same(acc, t, M > 0)) {
                std::printf("%s: whole item mismatch (chain %d, walker %d) n=%d: kmers %llu/%llu not_rc %llu/%llu ksum %llx/%llx mini %llu/%llu msum %llx/%llx\n  %s\n",
                            name, which, used, n, (unsigned long long)acc.n_kmers, (unsigned long long)t.n_kmers, (unsigned long long)acc.n_not_rc,
                            (unsigned long long)t.n_not_rc, (unsigned long long)acc.ksum_lo, (unsigned long long)t.kmer_sum_lo, (unsigned long long)acc.n_mini,
                            (unsigned long long)t.n_minimizers, (unsigned long long)acc.msum, (unsigned long long)t.minimizer_sum, s.c_str());
                fails++;
            }
        }
        // 2. fragments: k-mers are owned by their last base
        if (n >= 2) {
            int c1 = (int)(rng() % (unsigned)n), c2 = (int)(rng() % (unsigned)n);
            if (c1 > c2) std::swap(c1, c2);
            for (int which : std::vector<int>{CLEAN | FAST | GENERIC, CLEANFP | FAST | GENERIC, GENERIC}) {
                Acc acc;
                // (a '\r' ending a fragment is a deleted byte wherever it is: run_item's trim is harmless for inner fragments)
                auto frag = [&](int a, int b, bool) { if (b > a) item<K, M>(sb, L, a, b, 0, which, acc); };
                frag(0, c1, false); frag(c1, c2, false); frag(c2, n, true);
                if (!same(acc, t, M > 0)) { std::printf("%s: fragment sum mismatch (chain %d) n=%d cuts %d,%d\n  %s\n", name, which, n, c1, c2, s.c_str()); fails++; }
            }
        }
    }
    std::printf("%s: %d items, walker used: clean %ld (fp variant %ld), fast %ld, generic %ld; %s\n", name, iters, used_cnt[CLEAN], used_cnt[CLEANFP],
                used_cnt[FAST], used_cnt[GENERIC], fails ? "FAIL" : "ok");
    return fails;
}

// scan_row: newline count and word mask of a 256 B row, for every lane rotation
static int test_scan_row(std::mt19937_64& rng) {
    int fails = 0;
    alignas(16) uint8_t row[256];
    for (int it = 0; it < 4000 && fails < 5; it++) {
        const int dens = (int)(rng() % 5);                    // newline density: none .. all
        for (auto& c : row) {
            const uint64_t r = rng();
            const bool nl = dens == 0 ? false : dens == 4 ? true : (r % (dens == 1 ? 150 : dens == 2 ? 8 : 2)) == 0;
            c = nl ? '\n' : (uint8_t)((r >> 16) % 5 == 0 ? ((r >> 24) & 0xFF) : "ACGT@+\r\x0b\x8a\x0aIII"[(r >> 8) % 9]);
            if (!nl && c == '\n') c = 0x8A;                    // look-alikes: 0x8A, 0x0B, 0x09 must not count
        }
        uint32_t want_cnt = 0; uint64_t want_mask = 0;
        for (int i = 0; i < 256; i++) if (row[i] == '\n') { want_cnt++; want_mask |= 1ull << (i / 4); }
        for (uint32_t lane = 0; lane < 32; lane++) {
            uint32_t cnt; uint64_t mask;
            fused::scan_row(reinterpret_cast<const uint32_t*>(row), lane, cnt, mask);
            if (cnt != want_cnt || mask != want_mask) { std::printf("scan_row lane %u: cnt %u/%u mask %llx/%llx\n", lane, cnt, want_cnt, (unsigned long long)mask, (unsigned long long)want_mask); fails++; break; }
        }
    }
    std::printf("scan_row: %s\n", fails ? "FAIL" : "ok");
    return fails;
}

int main() {
    build_cls();
    std::mt19937_64 rng(20240917);
    int fails = test_scan_row(rng);
    fails += run<31, 21>(rng, 6000, "k31 m21");
    fails += run<21, 11>(rng, 6000, "k21 m11");
    fails += run<31, 0>(rng, 6000, "k31 m0");
    fails += run<31, 31>(rng, 1500, "k31 m31");
    fails += run<25, 24>(rng, 1500, "k25 m24");
    if (fails) { std::printf("walkers FAILED\n"); return 1; }
    std::printf("walkers ok\n");
    return 0;
}
