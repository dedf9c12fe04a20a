// Host-side differential test of the fused kernel's row scanner (scan_row), the deferred first-lines events
// (first_lines_event) and the per-item walkers
// (needletail_b200/csrc/fused.cuh: find_ws, walk, walk_fast, walk_clean) against the oracle's tally_sequence (oracle/ntref.hpp).  The walkers are __host__ __device__:
// this file compiles the very same source for the CPU (no device code is launched) and checks, on random lines,
//   1. every walker that accepts an item returns exactly the oracle's tallies for it,
//   2. an item cut into fragments (each warmed up over its k-1 predecessors, as the kernel does at tile / piece
//      boundaries) tallies to the same totals as the whole item  — k-mers are owned by their last base.
//   nvcc -std=c++17 --expt-relaxed-constexpr -o /tmp/test_walkers tests/cpp/test_walkers.cu && /tmp/test_walkers
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
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
    return a.n_kmers == t.n_kmers && a.n_not_rc == t.n_not_rc && a.ksum_lo == t.kmer_sum_lo && a.ksum_hi == t.kmer_sum_hi &&
           (!mini || (a.n_mini == t.n_minimizers && a.msum == t.minimizer_sum));
}

enum Which { CLEAN = 1, FAST = 2, GENERIC = 4 };
// the kernel's run_item chain on sb[a..b) with warm-up not below lo; `which` selects the walkers that may be used
template <int K, int M>
static bool item(const uint8_t* sb, const Luts& L, int a, int b, int lo, int which, Acc& acc, int* used = nullptr) {
    constexpr bool MINI = M > 0; constexpr int W = MINI ? K - M + 1 : 0;
    if (b > a && sb[b - 1] == '\r') b--;
    if (b <= a) return true;
    uint32_t slow = 0;
    const int ws = fused::find_ws(sb, g_cls, a, lo, true, K, slow);
    if ((which & CLEAN) && fused::walk_clean<K, M, false>(sb, L.comb, ws, b, 0, acc)) { if (used) *used = CLEAN; return true; }
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
        if (flavour == 4 || (flavour == 1 && it % 2 == 0))    // runs of T: RC_k(x) < x happens (the rare case of walk_clean / walk_fast), also next to non-ACGT bases
            for (int q = 0; q < 3 && n > 0; q++) { const int at = (int)(rng() % (unsigned)n), len = 4 + (int)(rng() % 30); for (int j = at; j < n && j < at + len; j++) s[j] = (rng() & 1) ? 'T' : 't'; }
        // (a line never contains '\n'; the kernel strips one trailing '\r' itself)
        const uint8_t* sb = (const uint8_t*)s.data();
        ntref::Tallies t;
        ntref::tally_sequence(sb, (size_t)n, K, M, false, nullptr, t, norm, rc);
        // 1. whole item through every admissible chain
        for (int which : std::vector<int>{CLEAN | FAST | GENERIC, FAST | GENERIC, GENERIC}) {
            Acc acc; int used = 0;
            item<K, M>(sb, L, 0, n, 0, which, acc, &used);
            if (which == (CLEAN | FAST | GENERIC)) used_cnt[used]++;
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
            for (int which : std::vector<int>{CLEAN | FAST | GENERIC, GENERIC}) {
                Acc acc;
                // (a '\r' ending a fragment is a deleted byte wherever it is: run_item's trim is harmless for inner fragments)
                auto frag = [&](int a, int b, bool) { if (b > a) item<K, M>(sb, L, a, b, 0, which, acc); };
                frag(0, c1, false); frag(c1, c2, false); frag(c2, n, true);
                if (!same(acc, t, M > 0)) { std::printf("%s: fragment sum mismatch (chain %d) n=%d cuts %d,%d\n  %s\n", name, which, n, c1, c2, s.c_str()); fails++; }
            }
        }
    }
    std::printf("%s: %d items, walker used: clean %ld, fast %ld, generic %ld; %s\n", name, iters, used_cnt[CLEAN],
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

// first_lines_event: the line events of a tile's first four lines, evaluated by the deferred look-back from global memory
static int test_first_lines_event(std::mt19937_64& rng) {
    using fused::SState; using fused::NONE;
    int fails = 0; long n_err = 0, n_rec = 0;
    for (int it = 0; it < 30000 && fails < 5; it++) {
        // a FASTQ-like text: mostly valid records, sometimes CRLF, sometimes damaged
        std::string s;
        const int nrec = 1 + (int)(rng() % 6);
        const bool crlf = rng() % 4 == 0;
        for (int r = 0; r < nrec; r++) {
            const int L = (int)(rng() % 12);
            std::string seq(L, 'A'), qual(L, 'I');
            for (auto& c : seq) c = "ACGT"[rng() & 3];
            for (auto& c : qual) c = "!+@I5"[rng() % 5];
            if (rng() % 12 == 0) qual += 'I';                                  // unequal lengths
            const char h = rng() % 15 == 0 ? 'x' : '@', p = rng() % 15 == 0 ? '-' : '+';
            const std::string eol = crlf && rng() % 8 ? "\r\n" : "\n";
            s += std::string(1, h) + "r" + eol + seq + eol + std::string(1, p) + eol + qual + eol;
        }
        if (rng() % 5 == 0 && !s.empty()) s.pop_back();                        // no final newline
        const uint64_t n = s.size();
        const uint8_t* bytes = (const uint8_t*)s.data();
        std::vector<uint64_t> gnl;                                             // all newline positions
        for (uint64_t q = 0; q < n; q++) if (bytes[q] == '\n') gnl.push_back(q);
        const uint64_t tile_start = rng() % n;
        const uint32_t TB = 8 + (uint32_t)(rng() % 200);
        const uint32_t avail = (uint32_t)std::min<uint64_t>(TB, n - tile_start);
        SState pre = fused::identity_state();
        size_t g0 = 0;                                                         // index of the tile's first newline in gnl
        while (g0 < gnl.size() && gnl[g0] < tile_start) g0++;
        pre.count = g0;
        for (int j = 0; j < 4 && j < (int)g0; j++) pre.last[j] = gnl[g0 - 1 - j];
        uint32_t Cs = 0, nl4[4] = {0, 0, 0, 0};
        for (size_t g = g0; g < gnl.size() && gnl[g] < tile_start + avail; g++) { if (Cs < 4) nl4[Cs] = (uint32_t)(gnl[g] - tile_start); Cs++; }
        const bool line0 = tile_start == 0 || bytes[tile_start - 1] == '\n';
        Acc got; uint32_t slow = 0; unsigned long long got_key = ~0ull, want_key = ~0ull;
        for (uint32_t i = 0; i < 4 && i <= Cs; i++) fused::first_lines_event(bytes, 0, tile_start, nl4, Cs, avail, line0, pre, i, got, slow, &got_key);
        auto want_note = [&](size_t g, uint32_t role, uint32_t check) {         // the record of line g starts at line g - role
            const size_t g0r = g - role;
            const uint64_t rs = g0r ? gnl[g0r - 1] + 1 : 0;
            want_key = std::min<unsigned long long>(want_key, (rs << 2) | check);
        };
        // the definition, on global line ordinals
        uint64_t want_bases = 0, want_rec = 0; bool want_err = false;
        auto line_len = [&](uint64_t ls, uint64_t q) { return (q - ls) - ((q > ls && bytes[q - 1] == '\r') ? 1 : 0); };
        for (uint32_t i = 0; i < 4 && i <= Cs; i++) {
            const size_t g = g0 + i;                                           // global ordinal of this line
            const uint32_t role = (uint32_t)(g & 3);
            const uint64_t ls = g ? gnl[g - 1] + 1 : 0;                        // global start of the line
            const bool starts_in_tile = ls >= tile_start && ls - tile_start < avail;
            if (starts_in_tile && role == 0 && bytes[ls] != '@') { want_err = true; want_note(g, 0, 0); }
            if (starts_in_tile && role == 2 && bytes[ls] != '+') { want_err = true; want_note(g, 2, 1); }
            if (i < Cs) {
                const uint64_t q = gnl[g];
                if (role == 1) want_bases += line_len(ls, q);
                else if (role == 3) {
                    if (g < 3) { want_err = true; want_key = 0; }
                    else {
                        const uint64_t q0 = gnl[g - 3], q1 = gnl[g - 2], q2 = gnl[g - 1];
                        if (line_len(q0 + 1, q1) != line_len(q2 + 1, q)) { want_err = true; want_note(g, 3, 2); }
                        want_rec++;
                    }
                }
            }
        }
        const bool got_err = (slow & fused::FLAG_PARSE_ERROR) != 0;
        n_err += want_err; n_rec += (long)want_rec;
        if (got.n_bases != want_bases || got.n_records != want_rec || got_err != want_err || got_key != want_key) {
            std::printf("first_lines_event: tile [%llu,+%u) Cs %u: bases %llu/%llu records %llu/%llu err %d/%d key %llx/%llx\n", (unsigned long long)tile_start, avail, Cs,
                        (unsigned long long)got.n_bases, (unsigned long long)want_bases, (unsigned long long)got.n_records, (unsigned long long)want_rec, got_err, want_err,
                        got_key, want_key);
            fails++;
        }
    }
    std::printf("first_lines_event: %ld records completed, %ld tiles with an error; %s\n", n_rec, n_err, fails ? "FAIL" : "ok");
    return fails;
}

// walk_clean2 (33 <= K <= 63, two-word k-mers) and the generic two-word walker against the oracle
template <int K>
static int run2(std::mt19937_64& rng, int iters, const char* name) {
    uint32_t comb2[256];
    for (int b = 0; b < 256; b++) comb2[b] = ((3u - (g_cls[b] & 3u)) << fused::Clean2Shape<K>::SH) | g_cls[b];
    int fails = 0; long n_clean = 0;
    std::vector<uint8_t> norm, rc;
    for (int it = 0; it < iters && fails < 5; it++) {
        const int flavour = (int)(rng() % 3);                 // 0: clean, 1: a few non-ACGT, 2: + whitespace
        const int n = (int)(rng() % (it % 5 == 0 ? 900 : 300));
        std::string s(n, 'A');
        for (auto& c : s) {
            const uint64_t r = rng();
            c = "ACGTacgt"[r & 7];
            const unsigned roll = (unsigned)((r >> 8) % 1000);
            if (flavour >= 1 && roll < 6u) c = "NnRY-.Xx"[(r >> 20) % 8];
            if (flavour >= 2 && roll >= 992) c = " \t\r"[(r >> 30) % 3];
        }
        if (it % 9 == 0) for (auto& c : s) c = (rng() % 40) ? 'A' : 'T';      // low complexity: long runs, F == R prefixes
        const uint8_t* sb = (const uint8_t*)s.data();
        ntref::Tallies t;
        ntref::tally_sequence(sb, (size_t)n, K, 0, false, nullptr, t, norm, rc);
        auto frag = [&](int a, int b, bool clean_first, Acc& acc) {
            if (b > a && sb[b - 1] == '\r') b--;
            if (b <= a) return;
            uint32_t slow = 0;
            const int ws = fused::find_ws(sb, g_cls, a, 0, true, K, slow);
            if (clean_first && fused::walk_clean2<K>(sb, comb2, ws, b, acc)) { n_clean++; return; }
            Params P{}; P.k = K; P.m = 0; P.w = 1; P.has_query = 0;
            fused::walk<2, false, 0>(sb, g_cls, ws, a, b, P, acc, false);
        };
        int c1 = n ? (int)(rng() % (unsigned)n) : 0, c2 = n ? (int)(rng() % (unsigned)n) : 0;
        if (c1 > c2) std::swap(c1, c2);
        for (int mode = 0; mode < 4; mode++) {                // whole / fragments  x  clean-first / generic only
            Acc acc;
            if (mode < 2) frag(0, n, mode == 0, acc);
            else { frag(0, c1, mode == 2, acc); frag(c1, c2, mode == 2, acc); frag(c2, n, mode == 2, acc); }
            if (!same(acc, t, false)) {
                std::printf("%s: mismatch (mode %d) n=%d cuts %d,%d: kmers %llu/%llu not_rc %llu/%llu lo %llx/%llx hi %llx/%llx\n", name, mode, n, c1, c2,
                            (unsigned long long)acc.n_kmers, (unsigned long long)t.n_kmers, (unsigned long long)acc.n_not_rc, (unsigned long long)t.n_not_rc,
                            (unsigned long long)acc.ksum_lo, (unsigned long long)t.kmer_sum_lo, (unsigned long long)acc.ksum_hi, (unsigned long long)t.kmer_sum_hi);
                fails++;
            }
        }
    }
    std::printf("%s: %d items, clean two-word walks %ld; %s\n", name, iters, n_clean, fails ? "FAIL" : "ok");
    return fails;
}

// wrapped sequences: items are the lines; a line's warm-up crosses the previous line break and is fed to walk_clean<WARM> as codes
template <int K, int M>
static int run_wrapped(std::mt19937_64& rng, int iters, const char* name) {
    const Luts L = make_luts<K>();
    int fails = 0; long n_w = 0, n_other = 0;
    std::vector<uint8_t> norm, rc;
    for (int it = 0; it < iters && fails < 5; it++) {
        const int n = 1 + (int)(rng() % 600), width = 20 + (int)(rng() % 70);
        const bool crlf = rng() % 3 == 0, dirty = rng() % 4 == 0;
        std::string s;
        std::vector<std::pair<int, int>> lines;             // [a, b) of every line, line break excluded
        for (int i = 0, col = 0, a = 0; i < n; i++) {
            char c = "ACGTacgt"[rng() & 7];
            if (dirty && rng() % 97 == 0) c = 'N';
            s += c;
            if (++col == width || i == n - 1) {
                lines.push_back({a, (int)s.size()});
                if (i != n - 1 || rng() % 2) s += crlf ? "\r\n" : "\n";
                a = (int)s.size(); col = 0;
            }
        }
        const uint8_t* sb = (const uint8_t*)s.data();
        ntref::Tallies t;
        ntref::tally_sequence(sb, s.size(), K, M, false, nullptr, t, norm, rc);
        Acc acc;
        for (auto [a, b] : lines) {
            if (b <= a) continue;
            uint32_t slow = 0; int got = 0; uint64_t wcodes = 0;
            const int ws = fused::find_ws_codes(sb, g_cls, a, 0, true, K, slow, got, wcodes);
            if (got == K - 1 && a - ws != got && fused::walk_clean<K, M, true>(sb, L.comb, a, b, wcodes, acc)) { n_w++; continue; }
            n_other++;
            item<K, M>(sb, L, a, b, 0, CLEAN | FAST | GENERIC, acc);
        }
        if (!same(acc, t, M > 0)) {
            std::printf("%s: wrapped mismatch n=%d width=%d: kmers %llu/%llu mini %llu/%llu msum %llx/%llx\n", name, n, width, (unsigned long long)acc.n_kmers,
                        (unsigned long long)t.n_kmers, (unsigned long long)acc.n_mini, (unsigned long long)t.n_minimizers, (unsigned long long)acc.msum,
                        (unsigned long long)t.minimizer_sum);
            fails++;
        }
    }
    std::printf("%s wrapped: %d sequences, %ld lines through walk_clean<WARM>, %ld through the other walkers; %s\n", name, iters, n_w, n_other, fails ? "FAIL" : "ok");
    return fails;
}

int main() {
    build_cls();
    std::mt19937_64 rng(20240917);
    int fails = test_scan_row(rng);
    fails += test_first_lines_event(rng);
    fails += run<31, 21>(rng, 6000, "k31 m21");
    fails += run<21, 11>(rng, 6000, "k21 m11");
    fails += run<31, 0>(rng, 6000, "k31 m0");
    fails += run<31, 31>(rng, 1500, "k31 m31");
    fails += run<25, 24>(rng, 1500, "k25 m24");
    fails += run_wrapped<31, 21>(rng, 3000, "k31 m21");
    fails += run_wrapped<21, 11>(rng, 3000, "k21 m11");
    fails += run_wrapped<31, 0>(rng, 2000, "k31 m0");
    fails += run2<51>(rng, 4000, "k51");
    fails += run2<63>(rng, 1500, "k63");
    fails += run2<35>(rng, 1500, "k35");
    if (fails) { std::printf("walkers FAILED\n"); return 1; }
    std::printf("walkers ok\n");
    return 0;
}
