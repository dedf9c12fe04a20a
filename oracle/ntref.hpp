// ntref.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A literal C++17 restatement of needletail v0.7.3's per-record hot path.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.  The product (libntgpu.so) never links,
// loads or calls this.
//
// Parity status: PINNED.  Every function below is checked by
// tests/test_oracle_vectors.py against the reference's own known-answer tests
// (file:line cited per test) and the constants asserted in
// benches/benchmark.rs:43-44,66-67,151,166,180 (570 records / 738 580 bases /
// 718 007 k-mers / 350 983 forward-canonical on tests/data/28S.fasta).
// The Rust reference itself cannot be built here (no rustc/cargo in the image).
// ntref_incremental.hpp restates the readers' incremental buffer management; tests/test_oracle_incremental.py
// shows the whole-buffer parsers below give the same results for every buffer capacity.
//
// All `ref:` citations are relative to /root/reference/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace ntref {

// ---------------------------------------------------------------------------
// sequence.rs
// ---------------------------------------------------------------------------

// ref: src/sequence.rs:19-62  normalize(seq, allow_iupac) -> Option<Vec<u8>>
// Returns true when something changed (Some), false for None; `out` always
// receives the normalized bytes (== input when unchanged).
inline bool normalize(const uint8_t* seq, size_t n, bool allow_iupac, std::vector<uint8_t>& out) {
    out.clear();
    out.reserve(n);
    bool changed = false;
    for (size_t i = 0; i < n; i++) {
        uint8_t c = seq[i];
        uint8_t nc;
        bool ch;
        switch (c) {
            case 'A': case 'C': case 'G': case 'T': case 'N': case '-':
                nc = c; ch = false; break;
            case 'a': nc = 'A'; ch = true; break;
            case 'c': nc = 'C'; ch = true; break;
            case 'g': nc = 'G'; ch = true; break;
            case 't': case 'u': case 'U': nc = 'T'; ch = true; break;
            case '.': case '~': nc = '-'; ch = true; break;
            case 'B': case 'D': case 'H': case 'V': case 'R':
            case 'Y': case 'S': case 'W': case 'K': case 'M':
                if (allow_iupac) { nc = c; ch = false; } else { nc = 'N'; ch = true; }
                break;
            case 'b': case 'd': case 'h': case 'v': case 'r':
            case 'y': case 's': case 'w': case 'k': case 'm':
                if (allow_iupac) { nc = (uint8_t)(c - 32); ch = true; } else { nc = 'N'; ch = true; }
                break;
            case ' ': case '\t': case '\r': case '\n':
                nc = ' '; ch = true; break;
            default:
                nc = 'N'; ch = true; break;
        }
        changed = changed || ch;
        if (nc != ' ') out.push_back(nc);
    }
    return changed;
}

// ref: src/sequence.rs:67-105  complement(n)
inline uint8_t complement(uint8_t n) {
    switch (n) {
        case 'a': return 't'; case 'A': return 'T';
        case 'c': return 'g'; case 'C': return 'G';
        case 'g': return 'c'; case 'G': return 'C';
        case 't': return 'a'; case 'T': return 'A';
        case 'r': return 'y'; case 'y': return 'r';
        case 'k': return 'm'; case 'm': return 'k';
        case 'b': return 'v'; case 'v': return 'b';
        case 'd': return 'h'; case 'h': return 'd';
        case 's': return 's'; case 'w': return 'w';
        case 'R': return 'Y'; case 'Y': return 'R';
        case 'K': return 'M'; case 'M': return 'K';
        case 'B': return 'V'; case 'V': return 'B';
        case 'D': return 'H'; case 'H': return 'D';
        case 'S': return 'S'; case 'W': return 'W';
        default: return n;
    }
}

// ref: src/sequence.rs:202-208  Sequence::reverse_complement
inline void reverse_complement(const uint8_t* seq, size_t n, std::vector<uint8_t>& out) {
    out.resize(n);
    for (size_t i = 0; i < n; i++) out[i] = complement(seq[n - 1 - i]);
}

// ref: src/sequence.rs:165-191  Sequence::strip_returns (removes every \r and \n)
inline bool strip_returns(const uint8_t* seq, size_t n, std::vector<uint8_t>& out) {
    out.clear();
    out.reserve(n);
    bool changed = false;
    for (size_t i = 0; i < n; i++) {
        if (seq[i] == '\r' || seq[i] == '\n') changed = true;
        else out.push_back(seq[i]);
    }
    return changed;
}

// lexicographic byte-slice `<` (Rust `&[u8] < &[u8]`)
inline bool slice_lt(const uint8_t* a, size_t na, const uint8_t* b, size_t nb) {
    size_t n = na < nb ? na : nb;
    int c = n ? std::memcmp(a, b, n) : 0;
    if (c != 0) return c < 0;
    return na < nb;
}

// ref: src/sequence.rs:110-134  canonical(seq) -> Cow (original on ties)
inline void str_canonical(const uint8_t* seq, size_t n, std::vector<uint8_t>& out) {
    std::vector<uint8_t> buf;
    buf.reserve(n);
    bool enough = false, original_was_canonical = false;
    for (size_t i = 0; i < n; i++) {
        uint8_t rn = complement(seq[n - 1 - i]);
        uint8_t nn = seq[i];
        buf.push_back(rn);
        if (!enough && nn < rn) { original_was_canonical = true; break; }
        else if (!enough && rn < nn) enough = true;
    }
    if (!original_was_canonical && enough) out = buf;
    else out.assign(seq, seq + n);
}

// ref: src/sequence.rs:139-152  minimizer(seq, length)
inline void str_minimizer(const uint8_t* seq, size_t n, size_t length, std::vector<uint8_t>& out) {
    std::vector<uint8_t> rc;
    reverse_complement(seq, n, rc);
    out.assign(seq, seq + length);
    for (size_t i = 0; i + length <= n; i++) {
        if (slice_lt(seq + i, length, out.data(), length)) out.assign(seq + i, seq + i + length);
        if (slice_lt(rc.data() + i, length, out.data(), length)) out.assign(rc.data() + i, rc.data() + i + length);
    }
}

// ref: src/sequence.rs:280-297  QualitySequence::quality_mask
inline void quality_mask(const uint8_t* seq, const uint8_t* qual, size_t n, uint8_t score,
                         std::vector<uint8_t>& out) {
    out.resize(n);
    for (size_t i = 0; i < n; i++) out[i] = qual[i] < score ? (uint8_t)'N' : seq[i];
}

// ---------------------------------------------------------------------------
// kmer.rs
// ---------------------------------------------------------------------------

// ref: src/kmer.rs:6-8
inline bool is_good_base(uint8_t c) {
    return c == 'a' || c == 'c' || c == 'g' || c == 't' || c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

// ref: src/kmer.rs:48-130  CanonicalKmers (same control flow as update_position/next)
struct CanonicalKmers {
    size_t k, start_pos;
    const uint8_t* buffer; size_t n;
    const uint8_t* rc_buffer; size_t rc_n;

    CanonicalKmers(const uint8_t* buf, size_t n_, const uint8_t* rc, size_t rc_n_, uint8_t k_)
        : k(k_), start_pos(0), buffer(buf), n(n_), rc_buffer(rc), rc_n(rc_n_) {
        update_position(true);
    }
    // ref: src/kmer.rs:84-108
    bool update_position(bool initial) {
        if (start_pos + k > n) return false;
        size_t kmer_len = initial ? 0 : k - 1;
        size_t stop_len = initial ? k - 1 : k;
        while (kmer_len < stop_len) {
            if (is_good_base(buffer[start_pos + kmer_len])) {
                kmer_len += 1;
            } else {
                kmer_len = 0;
                start_pos += kmer_len + 1;  // (sic) always advances by one: kmer_len was just zeroed
                if (start_pos + k > n) return false;
            }
        }
        return true;
    }
    // ref: src/kmer.rs:114-129 ; returns false for None
    bool next(size_t& pos, const uint8_t*& kmer, bool& was_rc) {
        if (!update_position(false)) return false;
        pos = start_pos;
        start_pos += 1;
        const uint8_t* result = buffer + pos;
        const uint8_t* rc_result = rc_buffer + (rc_n - pos - k);
        if (slice_lt(result, k, rc_result, k)) { kmer = result; was_rc = false; }
        else { kmer = rc_result; was_rc = true; }   // ties => rc slice, was_rc = true
        return true;
    }
};

// ---------------------------------------------------------------------------
// bitkmer.rs
// ---------------------------------------------------------------------------
typedef uint64_t BitKmerSeq;
struct BitKmer { BitKmerSeq v; uint8_t k; };

// ref: src/bitkmer.rs:5-18  NUC2BIT_LOOKUP ; returns -1 for None
inline int nuc2bit(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}
// 2^(2k) - 1 with Rust release-mode wrapping pow (k == 32 -> 2^64 wraps to 0, minus 1 -> all ones)
inline uint64_t mask2k(unsigned k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1); }

// ref: src/bitkmer.rs:26-36  extend_kmer
inline bool extend_kmer(BitKmer& kmer, uint8_t new_char) {
    int c = nuc2bit(new_char);
    if (c < 0) return false;
    uint64_t nk = (kmer.v << 2) + (uint64_t)c;
    kmer.v = nk & mask2k(kmer.k);
    return true;
}

// ref: src/bitkmer.rs:39-70  update_position
inline bool bit_update_position(size_t& start_pos, BitKmer& kmer, const uint8_t* buffer, size_t n, bool initial) {
    if (start_pos + kmer.k > n) return false;
    size_t kmer_len = initial ? 0 : (size_t)kmer.k - 1;
    size_t stop_len = initial ? (size_t)kmer.k - 1 : kmer.k;
    while (kmer_len < stop_len) {
        if (extend_kmer(kmer, buffer[start_pos + kmer_len])) {
            kmer_len += 1;
        } else {
            kmer_len = 0;
            kmer.v = 0;
            start_pos += kmer_len + 1;
            if (start_pos + kmer.k > n) return false;
        }
    }
    return true;
}

// ref: src/bitkmer.rs:112-132  reverse_complement
inline BitKmer bit_reverse_complement(BitKmer kmer) {
    uint64_t x = kmer.v;
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = ((x >> 32) & 0x00000000FFFFFFFFull) | ((x & 0x00000000FFFFFFFFull) << 32);
    x ^= 0xFFFFFFFFFFFFFFFFull;
    unsigned sh = 2 * (32 - kmer.k);
    x = sh >= 64 ? 0 : (x >> sh);
    return BitKmer{x, kmer.k};
}

// ref: src/bitkmer.rs:136-143  canonical (ties => original, false)
inline BitKmer bit_canonical(BitKmer kmer, bool& was_rc) {
    BitKmer rc = bit_reverse_complement(kmer);
    if (kmer.v > rc.v) { was_rc = true; return rc; }
    was_rc = false; return kmer;
}

// ref: src/bitkmer.rs:146-162  minimizer(kmer, minmer_size) ; note RC at width k (quirk A.9)
inline BitKmer bit_minimizer(BitKmer kmer, uint8_t minmer_size) {
    uint64_t new_kmer = kmer.v;
    uint64_t lowest = ~0ull;
    uint64_t bitmask = mask2k(minmer_size);
    for (unsigned i = 0; i <= (unsigned)(kmer.k - minmer_size); i++) {
        uint64_t cur = bitmask & new_kmer;
        if (cur < lowest) lowest = cur;
        BitKmer cur_rev = bit_reverse_complement(BitKmer{bitmask & new_kmer, kmer.k});
        if (cur_rev.v < lowest) lowest = cur_rev.v;
        new_kmer >>= 2;
    }
    return BitKmer{lowest, kmer.k};
}

// ref: src/bitkmer.rs:164-186  bitmer_to_bytes
inline void bitmer_to_bytes(BitKmer kmer, std::vector<uint8_t>& out) {
    out.clear();
    for (int i = (int)kmer.k - 1; i >= 0; i--) out.push_back("ACGT"[(kmer.v >> (2 * i)) & 3]);
}
// ref: src/bitkmer.rs:288-296  (test helper) bytes_to_bitmer
inline BitKmer bytes_to_bitmer(const uint8_t* s, size_t n) {
    BitKmer b{0, (uint8_t)n};
    for (size_t i = 0; i < n; i++) extend_kmer(b, s[i]);
    return b;
}

// ref: src/bitkmer.rs:72-109  BitNuclKmer iterator
struct BitNuclKmer {
    size_t start_pos; BitKmer cur; const uint8_t* buffer; size_t n; bool canonical;
    BitNuclKmer(const uint8_t* buf, size_t n_, uint8_t k, bool canon)
        : start_pos(0), cur{0, k}, buffer(buf), n(n_), canonical(canon) {
        bit_update_position(start_pos, cur, buffer, n, true);
    }
    bool next(size_t& pos, BitKmer& kmer, bool& was_rc) {
        if (!bit_update_position(start_pos, cur, buffer, n, false)) return false;
        start_pos += 1;
        pos = start_pos - 1;
        if (canonical) kmer = bit_canonical(cur, was_rc);
        else { kmer = cur; was_rc = false; }
        return true;
    }
};

// ---------------------------------------------------------------------------
// quality.rs
// ---------------------------------------------------------------------------
// ref: src/quality.rs:15-28  decode_phred ; returns false on PhredOffsetError
inline bool decode_phred(const uint8_t* qual, size_t n, int base64, std::vector<uint8_t>& out) {
    uint8_t off = base64 ? 64 : 33;
    out.clear();
    for (size_t i = 0; i < n; i++) {
        if (qual[i] < off) return false;
        out.push_back((uint8_t)(qual[i] - off));
    }
    return true;
}

// ---------------------------------------------------------------------------
// parser/  (whole-buffer instantiation: the reader's buffer capacity exceeds the
// input, so fill_buf loads everything and `buffer.len() < capacity` == EOF known.
// The reference's results do not depend on capacity — see SURVEY.md facts table.)
// ---------------------------------------------------------------------------
enum ErrKind : int {            // ref: src/errors.rs:28-43
    ERR_NONE = 0, ERR_IO = 1, ERR_UNKNOWN_FORMAT = 2, ERR_INVALID_START = 3,
    ERR_INVALID_SEPARATOR = 4, ERR_UNEQUAL_LENGTHS = 5, ERR_UNEXPECTED_END = 6, ERR_EMPTY_FILE = 7
};
enum Format : int { FMT_NONE = 0, FMT_FASTA = 1, FMT_FASTQ = 2 };
enum LineEnding : int { LE_NONE = 0, LE_UNIX = 1, LE_WINDOWS = 2 };

struct ParseError { int kind = ERR_NONE; uint64_t line = 0; std::string id; bool has_id = false; };

// One parsed record, as offsets into the input buffer (half-open ranges).
struct Record {
    uint64_t start;              // '@' / '>'
    uint64_t id_b, id_e;         // id()
    uint64_t seq_b, seq_e;       // raw_seq()
    uint64_t qual_b, qual_e;     // qual() (FASTQ only; 0,0 for FASTA)
    uint64_t all_e;              // all() == buf[start .. all_e)
    uint64_t num_bases;          // num_bases()
    uint64_t pos_line, pos_byte; // position()
};

// ref: src/parser/utils.rs:12-18
inline size_t trim_cr_end(const uint8_t* buf, size_t b, size_t e) {
    return (e > b && buf[e - 1] == '\r') ? e - 1 : e;
}
// ref: src/parser/utils.rs:106-117
inline int find_line_ending(const uint8_t* bytes, size_t n) {
    if (n) {
        const void* p = std::memchr(bytes, '\n', n);
        if (p) {
            size_t idx = (const uint8_t*)p - bytes;
            if (idx > 0 && bytes[idx - 1] == '\r') return LE_WINDOWS;
            return LE_UNIX;
        }
    }
    return LE_NONE;
}

struct ParseResult {
    int format = FMT_NONE;
    int line_ending = LE_NONE;
    std::vector<Record> records;
    ParseError err;                 // first error (records before it are still delivered)
    uint64_t final_line = 0, final_byte = 0;   // reader.position() after the last next()
};

// ref: src/parser/fastq.rs  Reader (find :155-187, validate :240-285, check_end :337-356, next :388-449)
inline void parse_fastq(const uint8_t* buf, size_t n, ParseResult& out) {
    out.format = FMT_FASTQ;
    size_t start = 0, end = 0, seq = 0, sep = 0, qual = 0;
    uint64_t line = 1, byte = 0;
    bool finished = false;
    if (n == 0) { out.final_line = line; out.final_byte = byte; return; }   // :397-401

    auto find_line = [&](size_t search_start, size_t& res) -> bool {        // :306-308
        if (search_start >= n) return false;
        const void* p = std::memchr(buf + search_start, '\n', n - search_start);
        if (!p) return false;
        res = ((const uint8_t*)p - buf) + 1;
        return true;
    };
    auto id_for_error = [&](std::string& id) -> bool {                       // :287-303 (parse_id=true branch)
        if (seq - start > 1) {
            size_t b = start + 1, e = trim_cr_end(buf, start + 1, seq - 1);
            size_t sp = b;
            while (sp < e && buf[sp] != ' ') sp++;
            id.assign((const char*)buf + b, sp - b);
            return true;
        }
        return false;
    };
    auto validate = [&]() -> bool {                                          // :240-285
        uint8_t sb = buf[start];
        if (sb != '@') { finished = true; out.err.kind = ERR_INVALID_START; out.err.line = line; return false; }
        uint8_t pb = buf[sep];
        if (pb != '+') {
            finished = true; out.err.kind = ERR_INVALID_SEPARATOR; out.err.line = line + 2;
            out.err.has_id = id_for_error(out.err.id); return false;
        }
        size_t seq_len = trim_cr_end(buf, seq, sep - 1) - seq;
        size_t qual_len = trim_cr_end(buf, qual, end) - qual;
        if (seq_len != qual_len) {
            finished = true; out.err.kind = ERR_UNEQUAL_LENGTHS; out.err.line = line;
            out.err.has_id = id_for_error(out.err.id); return false;
        }
        return true;
    };

    bool is_new = true;
    while (!finished) {
        if (!is_new) { byte += end + 1 - start; line += 4; start = end + 1; }  // :411-415
        int search_pos = 0;   // Id=0, Sequence=1, Separator=2, Quality=3
        bool complete = false;
        size_t p;
        // find() :155-187
        if (!find_line(start, p)) search_pos = 0;
        else { seq = p;
            if (!find_line(seq, p)) search_pos = 1;
            else { sep = p;
                if (!find_line(sep, p)) search_pos = 2;
                else { qual = p;
                    if (!find_line(qual, p)) search_pos = 3;
                    else { end = p - 1; complete = true; } } } }
        if (complete) { if (!validate()) break; }
        else {
            // next_complete() -> whole input is in the buffer -> check_end() :337-356
            finished = true;
            if (search_pos == 3) {
                end = n;
                if (!validate()) break;
            } else {
                bool all_blank = true;
                size_t ls = start;
                for (;;) {
                    size_t le = ls;
                    while (le < n && buf[le] != '\n') le++;
                    if (trim_cr_end(buf, ls, le) != ls) { all_blank = false; break; }
                    if (le >= n) break;
                    ls = le + 1;
                }
                if (all_blank) break;   // Ok(false) -> None
                out.err.kind = ERR_UNEXPECTED_END;
                out.err.line = line + (uint64_t)search_pos;
                if (search_pos > 0) out.err.has_id = id_for_error(out.err.id);
                break;
            }
        }
        is_new = false;   // is_new() is `end == 0`; end > 0 after any found record
        if (out.line_ending == LE_NONE) out.line_ending = find_line_ending(buf + start, end - start);  // :434-436
        Record r{};
        r.start = start;
        r.id_b = start + 1; r.id_e = trim_cr_end(buf, start + 1, seq - 1);
        r.seq_b = seq; r.seq_e = trim_cr_end(buf, seq, sep - 1);
        r.qual_b = qual; r.qual_e = trim_cr_end(buf, qual, end);
        r.all_e = end;
        r.num_bases = r.seq_e - r.seq_b;
        r.pos_line = line; r.pos_byte = byte;
        out.records.push_back(r);
    }
    out.final_line = line; out.final_byte = byte;
}

// ref: src/parser/fasta.rs  Reader (_find :220-243, find :200-216, next :291-367) + BufferPosition :16-108
inline void parse_fasta(const uint8_t* buf, size_t n, ParseResult& out) {
    out.format = FMT_FASTA;
    uint64_t line = 0, byte = 0;
    if (n == 0) return;                                                  // :299-302
    if (buf[0] != '>') { out.err.kind = ERR_INVALID_START; out.err.line = 0; return; }  // :315-323
    line = 1; byte = 0;
    size_t start = 0, search_pos = 1;
    std::vector<size_t> seq_pos;
    bool finished = false;
    bool first = true;
    while (!finished) {
        if (!first) {   // next_pos() :190-195 (is_new() false after a record was produced)
            line += seq_pos.size();
            byte += search_pos - start;
            start = search_pos;
            seq_pos.clear();
        }
        first = false;
        // _find() :220-243
        bool found = false;
        {
            size_t from = search_pos;
            bool broke = false;
            while (from < n) {
                const void* p = std::memchr(buf + from, '\n', n - from);
                if (!p) break;
                size_t pos = (const uint8_t*)p - buf;
                size_t next_line_start = pos + 1;
                if (next_line_start == n) { search_pos = pos; broke = true; break; }
                seq_pos.push_back(pos);
                if (buf[next_line_start] == '>') { search_pos = next_line_start; found = true; broke = true; break; }
                from = pos + 1;
            }
            if (!broke) search_pos = n;
        }
        if (!found) {   // find() EOF branch :205-213
            finished = true;
            if (!seq_pos.empty()) seq_pos.push_back(search_pos);
        }
        if (seq_pos.empty()) {   // :348-356
            out.err.kind = ERR_UNEXPECTED_END; out.err.line = line; break;
        }
        size_t last = seq_pos.back(), firstp = seq_pos.front();
        if (out.line_ending == LE_NONE) out.line_ending = find_line_ending(buf + start, last - start);
        Record r{};
        r.start = start;
        r.id_b = start + 1; r.id_e = trim_cr_end(buf, start + 1, firstp);
        if (seq_pos.size() > 1) { r.seq_b = firstp + 1; r.seq_e = trim_cr_end(buf, firstp + 1, last); }
        else { r.seq_b = r.seq_e = firstp; }   // b"" (offset is arbitrary for an empty slice)
        r.qual_b = r.qual_e = 0;
        r.all_e = last;
        uint64_t nb = r.seq_e - r.seq_b;       // num_bases :102-107
        for (size_t i = r.seq_b; i < r.seq_e; i++) if (buf[i] == '\n' || buf[i] == '\r') nb--;
        r.num_bases = nb;
        r.pos_line = line; r.pos_byte = byte;
        out.records.push_back(r);
    }
    out.final_line = line; out.final_byte = byte;
}

// ref: src/parser/mod.rs:85-150 (sniff on already-decompressed bytes) + :37-46
inline void parse_fastx(const uint8_t* buf, size_t n, ParseResult& out) {
    if (n < 2) { out.err.kind = ERR_EMPTY_FILE; return; }
    if (buf[0] == '>') parse_fasta(buf, n, out);
    else if (buf[0] == '@') parse_fastq(buf, n, out);
    else out.err.kind = ERR_UNKNOWN_FORMAT;
}

// ---------------------------------------------------------------------------
// Hot-path composition: the loop the README / benches run.
//   norm = rec.normalize(iupac)                 lib.rs:24 / benchmark.rs:34
//   rc   = norm.reverse_complement()            lib.rs:27 / benchmark.rs:35
//   for (pos,kmer,was_rc) in norm.canonical_kmers(k,&rc)   lib.rs:33 / benchmark.rs:36-41
//   for (pos,bk,_) in norm.bit_kmers(k,false): bitkmer::minimizer(bk,m)   (k<=32, m>0)
// ---------------------------------------------------------------------------
struct Tallies {
    uint64_t n_records = 0, n_bases = 0;
    uint64_t n_kmers = 0, n_not_rc = 0;
    uint64_t kmer_sum_lo = 0;    // wrapping sum of the low 64 bits of the 2-bit packed canonical k-mer
    uint64_t kmer_sum_hi = 0;    // wrapping sum of bits 64..127 of the pack (k > 32 only)
    uint64_t n_query = 0;        // canonical k-mers equal to the query (lib.rs:31-35)
    uint64_t n_minimizers = 0;   // bit-k-mer items fed to bitkmer::minimizer
    uint64_t minimizer_sum = 0;  // wrapping sum of minimizer values
    void add(const Tallies& o) {
        n_records += o.n_records; n_bases += o.n_bases; n_kmers += o.n_kmers; n_not_rc += o.n_not_rc;
        kmer_sum_lo += o.kmer_sum_lo; kmer_sum_hi += o.kmer_sum_hi; n_query += o.n_query;
        n_minimizers += o.n_minimizers; minimizer_sum += o.minimizer_sum;
    }
};

// 2-bit pack of an ASCII ACGT k-mer (first base most significant), k <= 64 -> (hi, lo)
inline void pack_kmer(const uint8_t* s, size_t k, uint64_t& hi, uint64_t& lo) {
    hi = 0; lo = 0;
    for (size_t i = 0; i < k; i++) {
        hi = (hi << 2) | (lo >> 62);
        lo = (lo << 2) | (uint64_t)nuc2bit(s[i]);
    }
}

inline void tally_sequence(const uint8_t* raw_seq, size_t n, unsigned k, unsigned m, bool iupac,
                           const uint8_t* query /* k bytes or null */, Tallies& t,
                           std::vector<uint8_t>& norm, std::vector<uint8_t>& rc) {
    normalize(raw_seq, n, iupac, norm);
    reverse_complement(norm.data(), norm.size(), rc);
    CanonicalKmers it(norm.data(), norm.size(), rc.data(), rc.size(), (uint8_t)k);
    size_t pos; const uint8_t* kmer; bool was_rc;
    while (it.next(pos, kmer, was_rc)) {
        t.n_kmers++;
        if (!was_rc) t.n_not_rc++;
        uint64_t hi, lo; pack_kmer(kmer, k, hi, lo);
        t.kmer_sum_lo += lo; t.kmer_sum_hi += hi;
        if (query && std::memcmp(kmer, query, k) == 0) t.n_query++;
    }
    if (m > 0 && k <= 32) {
        BitNuclKmer bit(norm.data(), norm.size(), (uint8_t)k, false);
        BitKmer bk; bool f;
        while (bit.next(pos, bk, f)) {
            t.n_minimizers++;
            t.minimizer_sum += bit_minimizer(bk, (uint8_t)m).v;
        }
    }
}

}  // namespace ntref
