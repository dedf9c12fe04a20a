// synth.hpp — CPU side of the deterministic synthetic FASTQ/FASTA generator
// (ORACLE / test infrastructure).  The product has its own device generator
// (needletail_b200/csrc/synth.cu) that must produce identical bytes; the spec is
// written down in DESIGN.md §"Synthetic inputs" and follows SURVEY.md §8(d).
//
//   rnd(seed, stream, rec, w) = mix64(seed + GOLDEN * ((rec << 26) | (stream << 24) | w))
//   bases   : stream 0, 32 bases per word, base j -> bits [2*(j%32), +2) of word j/32 -> "ACGT"
//   N inject: stream 1,  4 lanes per word, base j is 'N' iff 16-bit lane (j%4) of word j/4 < n_thresh
//   quality : stream 2,  8 bytes per word, '!' + ((byte * 42) >> 8)   (Phred33 '!'..'J')
//   FASTQ record: "@r%09llu\n" bases "\n+\n" quals "\n"      (2L+16 bytes)
//   FASTA record: ">r%08llu\n" bases "\n"                    (L+12 bytes, unwrapped)
#pragma once
#include <cstdint>
#include <cstddef>

namespace ntsynth {

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t rec, uint64_t w) {
    return mix64(seed + 0x9E3779B97F4A7C15ull * ((rec << 26) | (stream << 24) | w));
}
static inline size_t fastq_record_bytes(size_t L) { return 2 * L + 16; }
static inline size_t fasta_record_bytes(size_t L) { return L + 12; }

static inline void write_digits(uint8_t* p, uint64_t v, int nd) {
    for (int i = nd - 1; i >= 0; i--) { p[i] = (uint8_t)('0' + v % 10); v /= 10; }
}
static inline uint8_t base_at(uint64_t seed, uint64_t rec, size_t j, uint32_t n_thresh) {
    uint64_t w = rnd(seed, 0, rec, j >> 5);
    uint8_t b = (uint8_t)"ACGT"[(w >> (2 * (j & 31))) & 3];
    if (n_thresh) {
        uint64_t nw = rnd(seed, 1, rec, j >> 2);
        if (((nw >> (16 * (j & 3))) & 0xFFFF) < n_thresh) b = 'N';
    }
    return b;
}
static inline uint8_t qual_at(uint64_t seed, uint64_t rec, size_t j) {
    uint64_t w = rnd(seed, 2, rec, j >> 3);
    uint32_t by = (uint32_t)((w >> (8 * (j & 7))) & 0xFF);
    return (uint8_t)('!' + ((by * 42u) >> 8));
}

// writes records [rec0, rec0+nrec) ; out must hold nrec * fastq_record_bytes(L)
static inline void gen_fastq(uint8_t* out, uint64_t seed, uint64_t rec0, uint64_t nrec, size_t L, uint32_t n_thresh) {
    for (uint64_t r = 0; r < nrec; r++) {
        uint64_t rec = rec0 + r;
        uint8_t* p = out + r * fastq_record_bytes(L);
        p[0] = '@'; p[1] = 'r'; write_digits(p + 2, rec, 9); p[11] = '\n';
        uint8_t* s = p + 12;
        for (size_t j = 0; j < L; j++) s[j] = base_at(seed, rec, j, n_thresh);
        s[L] = '\n'; s[L + 1] = '+'; s[L + 2] = '\n';
        uint8_t* q = s + L + 3;
        for (size_t j = 0; j < L; j++) q[j] = qual_at(seed, rec, j);
        q[L] = '\n';
    }
}
static inline void gen_fasta(uint8_t* out, uint64_t seed, uint64_t rec0, uint64_t nrec, size_t L, uint32_t n_thresh) {
    for (uint64_t r = 0; r < nrec; r++) {
        uint64_t rec = rec0 + r;
        uint8_t* p = out + r * fasta_record_bytes(L);
        p[0] = '>'; p[1] = 'r'; write_digits(p + 2, rec, 8); p[10] = '\n';
        uint8_t* s = p + 11;
        for (size_t j = 0; j < L; j++) s[j] = base_at(seed, rec, j, n_thresh);
        s[L] = '\n';
    }
}

}  // namespace ntsynth
