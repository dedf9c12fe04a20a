// Host-side property test of the carried scan state of the fused kernel's decoupled look-back
// (needletail_b200/csrc/fused.cuh: SState / combine / identity_state): the look-back combines tile aggregates in
// arbitrary groupings (serial walk, 32-wide shuffle tree, chunk chaining), which is only valid if combine() is
// associative with identity_state() as a two-sided identity, and if folding per-tile aggregates reproduces the state
// of the concatenated span.  Compiled with nvcc, runs on the CPU (no device code is launched).
//   nvcc -std=c++17 -o /tmp/test_state_monoid tests/cpp/test_state_monoid.cu && /tmp/test_state_monoid
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../needletail_b200/csrc/fused.cuh"

using fused::SState; using fused::combine; using fused::identity_state; using fused::NONE; using fused::INHDR;

static bool eq(const SState& a, const SState& b) { return std::memcmp(&a, &b, sizeof(SState)) == 0; }

// state of a byte span [b0, b1) computed directly (the definition): newline count, last four newline positions,
// FASTA record starts ('>' at line start), header state at the end of the span, first newline
static SState direct(const std::string& s, size_t b0, size_t b1) {
    SState st = identity_state();
    std::vector<uint64_t> nls;
    bool have_start = false; uint64_t hdr = NONE;
    for (size_t p = b0; p < b1; p++) {
        const bool line_start = (p == 0) || s[p - 1] == '\n';
        if (line_start && s[p] == '>') { st.n_starts++; have_start = true; hdr = INHDR; }
        if (s[p] == '\n') {
            nls.push_back(p);
            if (st.first_nl == NONE) st.first_nl = p;
            if (have_start && hdr == INHDR) hdr = p;
        }
    }
    st.count = nls.size();
    for (int j = 0; j < 4 && j < (int)nls.size(); j++) st.last[j] = nls[nls.size() - 1 - j];
    st.hdr = have_start ? hdr : NONE;
    return st;
}

int main() {
    std::mt19937_64 rng(12345);
    int fails = 0;
    for (int it = 0; it < 20000 && fails < 5; it++) {
        const size_t n = 1 + rng() % 120;
        std::string s(n, 'A');
        for (auto& c : s) { const uint64_t r = rng() % 10; c = r < 2 ? '\n' : (r < 4 ? '>' : "ACGT"[r & 3]); }
        s[0] = '>';
        // random cut points
        size_t c1 = rng() % (n + 1), c2 = rng() % (n + 1);
        if (c1 > c2) std::swap(c1, c2);
        const SState A = direct(s, 0, c1), B = direct(s, c1, c2), Cc = direct(s, c2, n), all = direct(s, 0, n);
        const SState l = combine(combine(A, B), Cc), r = combine(A, combine(B, Cc));
        // (the aggregate of a span that does not start at byte 0 cannot know whether its first line fragment is a
        //  header; combine() resolves that from the earlier span — exactly what `direct` over the whole string gives)
        if (!eq(l, r)) { std::printf("associativity fails for %zu/%zu/%zu\n", c1, c2, n); fails++; }
        if (!eq(l, all)) { std::printf("fold != direct for cuts %zu,%zu of %zu: count %llu vs %llu hdr %llx vs %llx\n", c1, c2, n,
                                       (unsigned long long)l.count, (unsigned long long)all.count, (unsigned long long)l.hdr, (unsigned long long)all.hdr); fails++; }
        if (!eq(combine(identity_state(), B), B) || !eq(combine(B, identity_state()), B)) { std::printf("identity law fails\n"); fails++; }
        // the one-word FASTA look-back folds (kind, position) pairs: same header state as the full combine, in every grouping
        {
            using fused::FaState; using fused::fa_combine; using fused::fa_of;
            const FaState fl = fa_combine(fa_combine(fa_of(A), fa_of(B)), fa_of(Cc)), fr = fa_combine(fa_of(A), fa_combine(fa_of(B), fa_of(Cc)));
            const FaState want = fa_of(all);
            auto same = [](const FaState& x, const FaState& y) { return x.kind == y.kind && (x.kind == 0 || x.kind == 3 || x.pos == y.pos); };
            if (!same(fl, want) || !same(fr, want)) { std::printf("fa_combine != combine for cuts %zu,%zu of %zu: kind %u/%u/%u\n", c1, c2, n, fl.kind, fr.kind, want.kind); fails++; }
        }
    }
    if (fails) return 1;
    std::puts("state monoid ok");
    return 0;
}
