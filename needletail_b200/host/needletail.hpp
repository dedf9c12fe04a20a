// needletail.hpp — C++17 host-side mirror of needletail's public Rust surface (src/lib.rs:56-57) on top
// of libntgpu's C ABI (include/ntgpu.h).  The reference is compiled Rust and there is no Rust toolchain
// in this image, so the host layer above the boundary is written in C++ with the reference's names,
// argument meaning and error behaviour:
//
//   parse_fastx_file / parse_fastx_reader / parse_fastx_stdin   src/parser/mod.rs:85-163
//   FastxReader { next(), position(), line_ending() }            src/parser/utils.rs:119-130
//   SequenceRecord { id, raw_seq, seq, qual, all, num_bases, start_line_number, position, format,
//                    line_ending }                               src/parser/record.rs:57-154
//   Sequence { strip_returns, reverse_complement, normalize, canonical_kmers, bit_kmers }
//                                                                src/sequence.rs:156-253
//   bitkmer::{reverse_complement, canonical, minimizer}          src/bitkmer.rs:112-162
//
// The record scan runs on the GPU once per buffer (ntg_parse_fastx); next() then hands records out one by
// one in the reference's order, returning the ParseError exactly where the reference's iterator would.
// Per-record Sequence calls are batches of one; use the *_batch forms to amortise the FFI crossing.
#pragma once
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <iterator>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <variant>
#include <vector>

#include "../../include/ntgpu.h"

namespace needletail {

using Bytes = std::vector<uint8_t>;
using ByteView = std::basic_string_view<uint8_t>;

enum class ParseErrorKind { Io = 1, UnknownFormat, InvalidStart, InvalidSeparator, UnequalLengths, UnexpectedEnd, EmptyFile };  // errors.rs:28-43
enum class Format { Fasta = 1, Fastq = 2 };          // parser/utils.rs:75-88
enum class LineEnding { Windows = 2, Unix = 1 };     // parser/utils.rs:91-104

struct ErrorPosition { uint64_t line = 0; std::optional<std::string> id; };   // errors.rs:7-25
struct ParseError : std::runtime_error {                                       // errors.rs:46-56
    ParseErrorKind kind; ErrorPosition position; std::optional<Format> format;
    ParseError(ParseErrorKind k, ErrorPosition p, std::optional<Format> f)
        : std::runtime_error("needletail parse error kind " + std::to_string((int)k) + " at line " + std::to_string(p.line)),
          kind(k), position(std::move(p)), format(f) {}
};
struct GpuError : std::runtime_error { int status; GpuError(int s, const std::string& m) : std::runtime_error(m), status(s) {} };

struct Position { uint64_t line_ = 0, byte_ = 0; uint64_t line() const { return line_; } uint64_t byte() const { return byte_; } };   // utils.rs:52-72

// One libntgpu context per process/device; not thread-safe (same contract as `&mut` readers).
class Gpu {
public:
    explicit Gpu(int device = 0) { int s = ntg_create(device, &ctx_); if (s != NTG_OK) throw GpuError(s, "ntg_create failed: libntgpu has no CPU path"); }
    ~Gpu() { ntg_destroy(ctx_); }
    Gpu(const Gpu&) = delete; Gpu& operator=(const Gpu&) = delete;
    ntg_ctx* ctx() const { return ctx_; }
    void check(int s) const { if (s != NTG_OK) throw GpuError(s, ntg_last_error(ctx_)); }
    static Gpu& instance() { static Gpu g(0); return g; }
private:
    ntg_ctx* ctx_ = nullptr;
};

// ---- Sequence trait, batch forms ------------------------------------------------------------------
struct KmerItem { size_t pos; uint64_t lo, hi; bool was_rc; };      // (usize, kmer, bool): kmer is the 2-bit pack
struct Batch {
    Bytes cat; std::vector<uint64_t> offs{0};
    void push(ByteView s) { cat.insert(cat.end(), s.begin(), s.end()); offs.push_back(cat.size()); }
    size_t size() const { return offs.size() - 1; }
};
inline std::vector<std::pair<Bytes, bool>> normalize_batch(const Batch& b, bool iupac, Gpu& g = Gpu::instance()) {   // sequence.rs:19-62
    Bytes out(b.cat.size() + 1); std::vector<uint64_t> oo(b.size() + 1); Bytes ch(b.size() + 1);
    g.check(ntg_normalize(g.ctx(), b.cat.data(), b.offs.data(), b.size(), iupac, out.data(), oo.data(), ch.data()));
    std::vector<std::pair<Bytes, bool>> r;
    for (size_t i = 0; i < b.size(); i++) r.emplace_back(Bytes(out.begin() + oo[i], out.begin() + oo[i + 1]), ch[i] != 0);
    return r;
}
inline std::vector<Bytes> reverse_complement_batch(const Batch& b, Gpu& g = Gpu::instance()) {                     // sequence.rs:202-208
    Bytes out(b.cat.size() + 1);
    g.check(ntg_reverse_complement(g.ctx(), b.cat.data(), b.offs.data(), b.size(), out.data()));
    std::vector<Bytes> r;
    for (size_t i = 0; i < b.size(); i++) r.emplace_back(out.begin() + b.offs[i], out.begin() + b.offs[i + 1]);
    return r;
}
inline std::vector<std::vector<KmerItem>> items_to_vec(ntg_items* it) {
    std::vector<std::vector<KmerItem>> r(it->n_seqs);
    for (uint64_t s = 0; s < it->n_seqs; s++)
        for (uint64_t j = it->item_offs[s]; j < it->item_offs[s + 1]; j++)
            r[s].push_back(KmerItem{it->pos[j], it->val_lo[j], it->val_hi ? it->val_hi[j] : 0, it->was_rc ? it->was_rc[j] != 0 : false});
    ntg_items_free(it);
    return r;
}
inline std::vector<std::vector<KmerItem>> canonical_kmers_batch(const Batch& b, uint8_t k, const Batch* rc = nullptr, Gpu& g = Gpu::instance()) {  // kmer.rs:73-129
    ntg_items* it = nullptr;
    g.check(ntg_canonical_kmers(g.ctx(), b.cat.data(), rc ? rc->cat.data() : nullptr, b.offs.data(), b.size(), k, &it));
    return items_to_vec(it);
}
inline std::vector<std::vector<KmerItem>> bit_kmers_batch(const Batch& b, uint8_t k, bool canonical, Gpu& g = Gpu::instance()) {                // bitkmer.rs:72-109
    ntg_items* it = nullptr;
    g.check(ntg_bit_kmers(g.ctx(), b.cat.data(), b.offs.data(), b.size(), k, canonical, &it));
    return items_to_vec(it);
}
inline std::vector<std::vector<KmerItem>> bit_minimizers_batch(const Batch& b, uint8_t k, uint8_t m, Gpu& g = Gpu::instance()) {                // bitkmer.rs:146-162 per item
    ntg_items* it = nullptr;
    g.check(ntg_bit_minimizers(g.ctx(), b.cat.data(), b.offs.data(), b.size(), k, m, &it));
    return items_to_vec(it);
}

// `Sequence` for any byte slice (sequence.rs:255-271): methods named as in the trait
struct Sequence {
    ByteView s;
    explicit Sequence(ByteView v) : s(v) {}
    ByteView sequence() const { return s; }
    Bytes strip_returns() const {                                                                      // sequence.rs:165-191
        Gpu& g = Gpu::instance(); Batch b; b.push(s);
        Bytes out(s.size() + 1); uint64_t oo[2]; uint8_t ch[1];
        g.check(ntg_strip_returns(g.ctx(), b.cat.data(), b.offs.data(), 1, out.data(), oo, ch));
        out.resize(oo[1]); return out;
    }
    Bytes reverse_complement() const { Batch b; b.push(s); return reverse_complement_batch(b)[0]; }
    Bytes normalize(bool iupac) const { Batch b; b.push(s); return normalize_batch(b, iupac)[0].first; }   // Cow: equals the input when unchanged
    std::vector<KmerItem> canonical_kmers(uint8_t k, ByteView rc) const {                               // sequence.rs:237-239
        Batch b, r; b.push(s); r.push(rc); return canonical_kmers_batch(b, k, &r)[0];
    }
    std::vector<KmerItem> bit_kmers(uint8_t k, bool canonical) const { Batch b; b.push(s); return bit_kmers_batch(b, k, canonical)[0]; }
};
namespace bitkmer {   // element-wise helpers on BitKmer = (u64, k)
inline uint64_t reverse_complement(uint64_t v, uint8_t k) { uint64_t o; Gpu& g = Gpu::instance(); g.check(ntg_bitkmer_reverse_complement(g.ctx(), &v, 1, k, &o)); return o; }
inline std::pair<uint64_t, bool> canonical(uint64_t v, uint8_t k) { uint64_t o; uint8_t f; Gpu& g = Gpu::instance(); g.check(ntg_bitkmer_canonical(g.ctx(), &v, 1, k, &o, &f)); return {o, f != 0}; }
inline uint64_t minimizer(uint64_t v, uint8_t k, uint8_t m) { uint64_t o; Gpu& g = Gpu::instance(); g.check(ntg_bitkmer_minimizer(g.ctx(), &v, 1, k, m, &o)); return o; }
}  // namespace bitkmer

// ---- SequenceRecord / FastxReader -------------------------------------------------------------------
class FastxReader;
class SequenceRecord {                                                                                  // parser/record.rs:57-154
public:
    ByteView id() const { return view(r_.id_b, r_.id_e); }
    ByteView raw_seq() const { return view(r_.seq_b, r_.seq_e); }
    Bytes seq() const { Bytes o; for (uint8_t c : raw_seq()) if (c != '\r' && c != '\n') o.push_back(c); return o; }   // record.rs:84-89
    std::optional<ByteView> qual() const { return fmt_ == Format::Fastq ? std::optional<ByteView>(view(r_.qual_b, r_.qual_e)) : std::nullopt; }
    ByteView all() const { return view(r_.start, r_.all_e); }
    size_t num_bases() const { return r_.num_bases; }
    uint64_t start_line_number() const { return r_.line; }
    Position position() const { return Position{r_.line, r_.start}; }
    Format format() const { return fmt_; }
    LineEnding line_ending() const { return le_; }
    // impl Sequence for SequenceRecord: sequence() == raw_seq()  (record.rs:181-185)
    Sequence as_sequence() const { return Sequence(raw_seq()); }
    Bytes normalize(bool iupac) const { return as_sequence().normalize(iupac); }
private:
    friend class FastxReader;
    SequenceRecord(const uint8_t* buf, const ntg_record& r, Format f, LineEnding le) : buf_(buf), r_(r), fmt_(f), le_(le) {}
    ByteView view(uint64_t b, uint64_t e) const { return ByteView(buf_ + b, e - b); }
    const uint8_t* buf_; ntg_record r_; Format fmt_; LineEnding le_;
};

class FastxReader {                                                                                     // parser/utils.rs:119-130
public:
    explicit FastxReader(Bytes data, Gpu& g = Gpu::instance()) : data_(std::move(data)) {
        g.check(ntg_parse_fastx(g.ctx(), data_.data(), data_.size(), &recs_));
        // parse_fastx_reader fails up front on < 2 bytes / unknown first byte (mod.rs:88-91,44)
        if (recs_->error.kind == NTG_EEMPTY_FILE || recs_->error.kind == NTG_EUNKNOWN_FORMAT) { auto e = make_error(); ntg_records_free(recs_); recs_ = nullptr; throw e; }
    }
    ~FastxReader() { ntg_records_free(recs_); }
    FastxReader(const FastxReader&) = delete; FastxReader& operator=(const FastxReader&) = delete;
    // Option<Result<SequenceRecord, ParseError>>: nullopt == None; the variant is Ok / Err
    std::optional<std::variant<SequenceRecord, ParseError>> next() {
        if (i_ < recs_->n_records) {
            const ntg_record& r = recs_->records[i_++];
            pos_ = Position{r.line, r.start};
            if (!le_) le_ = recs_->line_ending ? std::optional<LineEnding>((LineEnding)recs_->line_ending) : std::nullopt;
            return std::variant<SequenceRecord, ParseError>(SequenceRecord(data_.data(), r, (Format)recs_->format, le_.value_or(LineEnding::Unix)));
        }
        if (recs_->error.kind && !error_given_) { error_given_ = true; pos_ = Position{recs_->final_line, recs_->final_byte}; return std::variant<SequenceRecord, ParseError>(make_error()); }
        if (!finished_) { finished_ = true; pos_ = Position{recs_->final_line, recs_->final_byte}; }
        return std::nullopt;
    }
    const Position& position() const { return pos_; }
    std::optional<LineEnding> line_ending() const { return le_; }
private:
    ParseError make_error() const {
        const ntg_parse_error& e = recs_->error;
        ErrorPosition p; p.line = e.line; if (e.has_id) p.id = std::string(e.id);
        return ParseError((ParseErrorKind)e.kind, p, e.format ? std::optional<Format>((Format)e.format) : std::nullopt);
    }
    Bytes data_; ntg_records* recs_ = nullptr; uint64_t i_ = 0; bool error_given_ = false, finished_ = false;
    Position pos_{1, 0}; std::optional<LineEnding> le_;
};

// parse_fastx_reader: any byte source already in memory (decompression stays on the host side of the boundary;
// gzip/bz2/xz/zstd streams must be inflated by the caller in this C++ face — the Python face does gzip/bz2/xz).
inline std::unique_ptr<FastxReader> parse_fastx_reader(Bytes data) { return std::make_unique<FastxReader>(std::move(data)); }   // mod.rs:85-150
inline std::unique_ptr<FastxReader> parse_fastx_file(const std::string& path) {                                                 // mod.rs:161-163
    std::ifstream f(path, std::ios::binary);
    if (!f) throw ParseError(ParseErrorKind::Io, {}, std::nullopt);
    Bytes data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return parse_fastx_reader(std::move(data));
}
inline std::unique_ptr<FastxReader> parse_fastx_stdin() {                                                                        // mod.rs:154-157
    Bytes data((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
    return parse_fastx_reader(std::move(data));
}
}  // namespace needletail
