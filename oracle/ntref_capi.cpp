// ntref_capi.cpp — flat C ABI over the CPU oracle (ntref.hpp) for ctypes.
// ORACLE / test infrastructure only: see the header of ntref.hpp.
#include "ntref.hpp"
#include "ntref_incremental.hpp"
#include "synth.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <sched.h>
#include <thread>

using namespace ntref;

extern "C" {

int ntref_normalize(const uint8_t* seq, size_t n, int iupac, uint8_t* out, size_t* out_n) {
    std::vector<uint8_t> v;
    bool ch = normalize(seq, n, iupac != 0, v);
    if (out && !v.empty()) std::memcpy(out, v.data(), v.size());
    *out_n = v.size();
    return ch ? 1 : 0;
}
uint8_t ntref_complement(uint8_t c) { return complement(c); }
void ntref_reverse_complement(const uint8_t* seq, size_t n, uint8_t* out) {
    std::vector<uint8_t> v; reverse_complement(seq, n, v);
    if (n) std::memcpy(out, v.data(), n);
}
int ntref_strip_returns(const uint8_t* seq, size_t n, uint8_t* out, size_t* out_n) {
    std::vector<uint8_t> v; bool ch = strip_returns(seq, n, v);
    if (!v.empty()) std::memcpy(out, v.data(), v.size());
    *out_n = v.size();
    return ch ? 1 : 0;
}
void ntref_str_canonical(const uint8_t* seq, size_t n, uint8_t* out) {
    std::vector<uint8_t> v; str_canonical(seq, n, v);
    if (n) std::memcpy(out, v.data(), n);
}
void ntref_str_minimizer(const uint8_t* seq, size_t n, size_t len, uint8_t* out) {
    std::vector<uint8_t> v; str_minimizer(seq, n, len, v);
    if (len) std::memcpy(out, v.data(), len);
}
void ntref_quality_mask(const uint8_t* seq, const uint8_t* qual, size_t n, uint8_t score, uint8_t* out) {
    std::vector<uint8_t> v; quality_mask(seq, qual, n, score, v);
    if (n) std::memcpy(out, v.data(), n);
}
int ntref_decode_phred(const uint8_t* q, size_t n, int base64, uint8_t* out) {
    std::vector<uint8_t> v;
    if (!decode_phred(q, n, base64, v)) return 0;
    if (n) std::memcpy(out, v.data(), n);
    return 1;
}

// Kmers (src/kmer.rs:13-41): plain windows; returns the count (positions are 0..count-1)
size_t ntref_kmers_count(size_t n, unsigned k) { return (k <= n) ? n - k + 1 : 0; }

// CanonicalKmers over (seq, rc).  pos/was_rc may be null (count only).  Returns total count.
size_t ntref_canonical_kmers(const uint8_t* seq, size_t n, const uint8_t* rc, size_t rc_n, unsigned k,
                             uint64_t* pos, uint8_t* was_rc, size_t cap) {
    CanonicalKmers it(seq, n, rc, rc_n, (uint8_t)k);
    size_t p; const uint8_t* km; bool f; size_t c = 0;
    while (it.next(p, km, f)) {
        if (pos && c < cap) { pos[c] = p; was_rc[c] = f ? 1 : 0; }
        c++;
    }
    return c;
}

size_t ntref_bit_kmers(const uint8_t* seq, size_t n, unsigned k, int canonical,
                       uint64_t* pos, uint64_t* kmer, uint8_t* was_rc, size_t cap) {
    BitNuclKmer it(seq, n, (uint8_t)k, canonical != 0);
    size_t p; BitKmer bk; bool f; size_t c = 0;
    while (it.next(p, bk, f)) {
        if (pos && c < cap) { pos[c] = p; kmer[c] = bk.v; was_rc[c] = f ? 1 : 0; }
        c++;
    }
    return c;
}
uint64_t ntref_bit_reverse_complement(uint64_t v, unsigned k) { return bit_reverse_complement(BitKmer{v, (uint8_t)k}).v; }
uint64_t ntref_bit_canonical(uint64_t v, unsigned k, int* was_rc) {
    bool f; BitKmer r = bit_canonical(BitKmer{v, (uint8_t)k}, f); *was_rc = f ? 1 : 0; return r.v;
}
uint64_t ntref_bit_minimizer(uint64_t v, unsigned k, unsigned m) { return bit_minimizer(BitKmer{v, (uint8_t)k}, (uint8_t)m).v; }
void ntref_bitmer_to_bytes(uint64_t v, unsigned k, uint8_t* out) {
    std::vector<uint8_t> b; bitmer_to_bytes(BitKmer{v, (uint8_t)k}, b);
    if (k) std::memcpy(out, b.data(), k);
}
uint64_t ntref_bytes_to_bitmer(const uint8_t* s, size_t n) { return bytes_to_bitmer(s, n).v; }

// Parse a whole (already decompressed) FASTX buffer.  recs: cap rows of 12 u64
// {start,id_b,id_e,seq_b,seq_e,qual_b,qual_e,all_e,num_bases,pos_line,pos_byte,0}.
// info (u64[8]): {format, line_ending, err_kind, err_line, err_has_id, final_line, final_byte, 0}
// Returns the number of records parsed before the first error / EOF.
static size_t export_parse(const ParseResult& pr, uint64_t* recs, size_t cap, uint64_t* info, char* err_id, size_t err_id_cap);

// Same output, but through the reference's incremental readers (buffer of `capacity` bytes that is refilled,
// shifted and grown; the underlying Read hands out at most `max_read` bytes per call).
size_t ntref_parse_fastx_incremental(const uint8_t* buf, size_t n, size_t capacity, size_t max_read, uint64_t* recs, size_t cap,
                                     uint64_t* info, char* err_id, size_t err_id_cap) {
    ParseResult pr;
    parse_fastx_incremental(buf, n, capacity, max_read ? max_read : (size_t)-1, pr);
    return export_parse(pr, recs, cap, info, err_id, err_id_cap);
}

size_t ntref_parse_fastx(const uint8_t* buf, size_t n, uint64_t* recs, size_t cap, uint64_t* info,
                         char* err_id, size_t err_id_cap) {
    ParseResult pr;
    parse_fastx(buf, n, pr);
    return export_parse(pr, recs, cap, info, err_id, err_id_cap);
}

static size_t export_parse(const ParseResult& pr, uint64_t* recs, size_t cap, uint64_t* info, char* err_id, size_t err_id_cap) {
    size_t c = pr.records.size();
    for (size_t i = 0; i < c && i < cap; i++) {
        const Record& r = pr.records[i];
        uint64_t* o = recs + 12 * i;
        o[0] = r.start; o[1] = r.id_b; o[2] = r.id_e; o[3] = r.seq_b; o[4] = r.seq_e; o[5] = r.qual_b;
        o[6] = r.qual_e; o[7] = r.all_e; o[8] = r.num_bases; o[9] = r.pos_line; o[10] = r.pos_byte; o[11] = 0;
    }
    info[0] = pr.format; info[1] = pr.line_ending; info[2] = pr.err.kind; info[3] = pr.err.line;
    info[4] = pr.err.has_id ? 1 : 0; info[5] = pr.final_line; info[6] = pr.final_byte; info[7] = 0;
    if (err_id && err_id_cap) {
        size_t l = std::min(err_id_cap - 1, pr.err.id.size());
        std::memcpy(err_id, pr.err.id.data(), l); err_id[l] = 0;
    }
    return c;
}

// The hot-path loop over one whole FASTX buffer: parse -> per record
// normalize -> reverse_complement -> canonical_kmers(k) [+ bit_kmers(k,false) -> minimizer(m)].
// out: u64[9] = {n_records,n_bases,n_kmers,n_not_rc,kmer_sum_lo,kmer_sum_hi,n_query,n_minimizers,minimizer_sum}
// Records before the first error are tallied (iterator semantics).  Returns the error kind (0 = none).
static int tally_one(const uint8_t* buf, size_t n, unsigned k, unsigned m, int iupac, const uint8_t* query, Tallies& t) {
    ParseResult pr;
    if (n == 0) return 0;
    if (buf[0] == '>') parse_fasta(buf, n, pr);
    else if (buf[0] == '@') parse_fastq(buf, n, pr);
    else return ERR_UNKNOWN_FORMAT;
    std::vector<uint8_t> norm, rc;
    for (const Record& r : pr.records) {
        t.n_records++;
        t.n_bases += r.num_bases;
        tally_sequence(buf + r.seq_b, r.seq_e - r.seq_b, k, m, iupac != 0, query, t, norm, rc);
    }
    return pr.err.kind;
}
static void store_tallies(const Tallies& t, uint64_t* out) {
    out[0] = t.n_records; out[1] = t.n_bases; out[2] = t.n_kmers; out[3] = t.n_not_rc; out[4] = t.kmer_sum_lo;
    out[5] = t.kmer_sum_hi; out[6] = t.n_query; out[7] = t.n_minimizers; out[8] = t.minimizer_sum;
}
int ntref_tally_fastx(const uint8_t* buf, size_t n, unsigned k, unsigned m, int iupac, const uint8_t* query,
                      uint64_t* out) {
    Tallies t;
    int e = (n < 2) ? (int)ERR_EMPTY_FILE : tally_one(buf, n, k, m, iupac, query, t);
    store_tallies(t, out);
    return e;
}

// Streaming variant that fuses parse + tally record by record (no record vector): the shape of the
// reference's bench loop (benches/benchmark.rs:32-41).  FASTQ only, used for CPU-baseline timing.
static void tally_fastq_stream(const uint8_t* buf, size_t n, unsigned k, unsigned m, int iupac, Tallies& t) {
    std::vector<uint8_t> norm, rc;
    size_t start = 0;
    while (start < n) {
        const uint8_t* p0 = (const uint8_t*)std::memchr(buf + start, '\n', n - start); if (!p0) break;
        size_t seq = p0 - buf + 1;
        const uint8_t* p1 = (const uint8_t*)std::memchr(buf + seq, '\n', n - seq); if (!p1) break;
        size_t sep = p1 - buf + 1;
        const uint8_t* p2 = (const uint8_t*)std::memchr(buf + sep, '\n', n - sep); if (!p2) break;
        size_t qual = p2 - buf + 1;
        const uint8_t* p3 = (const uint8_t*)std::memchr(buf + qual, '\n', n - qual);
        size_t end = p3 ? (size_t)(p3 - buf) : n;
        if (buf[start] != '@' || buf[sep] != '+') break;
        size_t se = trim_cr_end(buf, seq, sep - 1), qe = trim_cr_end(buf, qual, end);
        if (se - seq != qe - qual) break;
        t.n_records++; t.n_bases += se - seq;
        tally_sequence(buf + seq, se - seq, k, m, iupac != 0, nullptr, t, norm, rc);
        start = end + 1;
    }
}

// Multi-threaded CPU baseline: the buffer is cut at the caller-supplied record-aligned offsets
// (noff = nthreads+1 entries); each thread runs the full per-record loop over its slice.
// Returns elapsed seconds (steady_clock around the parallel region).
double ntref_bench_fastq(const uint8_t* buf, const uint64_t* offs, unsigned nthreads, unsigned k, unsigned m,
                         int iupac, uint64_t* out) {
    // Each thread tallies into its OWN stack-local Tallies (the per-k-mer increments must not share cache lines
    // between threads) and publishes it once at the end; the threads are created and parked on a start gate
    // before the clock starts, so thread creation is not timed.
    struct alignas(128) Slot { Tallies t; };
    std::vector<Slot> parts(nthreads);
    std::atomic<unsigned> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nthreads; i++)
        th.emplace_back([&, i] {
            Tallies local;
            ready.fetch_add(1);
            while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
            tally_fastq_stream(buf + offs[i], offs[i + 1] - offs[i], k, m, iupac, local);
            parts[i].t = local;
        });
    while (ready.load() < nthreads) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    Tallies t;
    for (auto& p : parts) t.add(p.t);
    store_tallies(t, out);
    return std::chrono::duration<double>(t1 - t0).count();
}
// Host threads this process may actually run on (cgroup / affinity mask), not the machine's core count.
unsigned ntref_usable_cores(void) {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) { int c = CPU_COUNT(&set); if (c > 0) return (unsigned)c; }
    unsigned h = std::thread::hardware_concurrency();
    return h ? h : 1;
}

// Synthetic generator (CPU side), multi-threaded over records.
void ntref_gen_fastq(uint8_t* out, uint64_t seed, uint64_t rec0, uint64_t nrec, size_t L, uint32_t n_thresh, unsigned nthreads) {
    if (nthreads <= 1) { ntsynth::gen_fastq(out, seed, rec0, nrec, L, n_thresh); return; }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nthreads; i++) {
        uint64_t a = nrec * i / nthreads, b = nrec * (i + 1) / nthreads;
        th.emplace_back([=] { ntsynth::gen_fastq(out + a * ntsynth::fastq_record_bytes(L), seed, rec0 + a, b - a, L, n_thresh); });
    }
    for (auto& x : th) x.join();
}
void ntref_gen_fasta(uint8_t* out, uint64_t seed, uint64_t rec0, uint64_t nrec, size_t L, uint32_t n_thresh, unsigned nthreads) {
    if (nthreads <= 1) { ntsynth::gen_fasta(out, seed, rec0, nrec, L, n_thresh); return; }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nthreads; i++) {
        uint64_t a = nrec * i / nthreads, b = nrec * (i + 1) / nthreads;
        th.emplace_back([=] { ntsynth::gen_fasta(out + a * ntsynth::fasta_record_bytes(L), seed, rec0 + a, b - a, L, n_thresh); });
    }
    for (auto& x : th) x.join();
}

}  // extern "C"
