// ntref_incremental.hpp — ORACLE (test infrastructure): the reference's FASTQ / FASTA readers restated
// *with* their incremental buffer management (refill, shift, grow), over a small model of
// buffer_redux::BufReader.  ntref.hpp's parse_fastq / parse_fasta instantiate the same readers with a
// buffer larger than the input; tests/test_oracle_incremental.py checks that both give identical results
// for capacities from 3 bytes up and for short reads — i.e. that the reference's results do not depend on
// its buffer capacity, which is what allows a whole-buffer (GPU) formulation to be bit-exact.
// All `ref:` citations are relative to /root/reference/.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "ntref.hpp"

namespace ntref {

// Model of buffer_redux::BufReader as the readers use it: capacity(), buffer(), read_into_buf(),
// consume(), make_room(), reserve().
struct BufModel {
    const uint8_t* src; size_t n, srcpos = 0;
    std::vector<uint8_t> mem; size_t pos = 0, end = 0, cap;
    size_t max_read;                      // the underlying Read returns at most this many bytes per call
    size_t dropped = 0;                   // bytes consumed so far (stream offset of mem[0] ... of buffer()[0] is dropped)
    BufModel(const uint8_t* s, size_t n_, size_t capacity, size_t max_read_) : src(s), n(n_), mem(capacity), cap(capacity), max_read(max_read_) {}
    size_t capacity() const { return cap; }
    const uint8_t* buf() const { return mem.data() + pos; }
    size_t len() const { return end - pos; }
    size_t read_into_buf() {
        size_t space = cap - end;
        size_t k = std::min(std::min(space, n - srcpos), max_read);
        if (k) { std::memcpy(mem.data() + end, src + srcpos, k); end += k; srcpos += k; }
        return k;
    }
    void consume(size_t k) { pos += k; dropped += k; }
    void make_room() { if (pos) { std::memmove(mem.data(), mem.data() + pos, end - pos); end -= pos; pos = 0; } }
    void reserve(size_t additional) { make_room(); cap += additional; mem.resize(cap); }
    uint64_t stream_off() const { return dropped; }     // stream offset of buf()[0]
};
// ref: src/parser/utils.rs:24-30
inline size_t grow_to(size_t current) { return current < (size_t(1) << 23) ? current * 2 : current + (size_t(1) << 23); }
// ref: src/parser/utils.rs:34-49
inline size_t fill_buf(BufModel& r) {
    size_t initial = r.len(), num_read = 0;
    while (initial + num_read < r.capacity()) {
        size_t k = r.read_into_buf();
        if (k == 0) break;
        num_read += k;
    }
    return num_read;
}

// ---------------------------------------------------------------------------------------------
// ref: src/parser/fastq.rs  Reader
inline void parse_fastq_incremental(const uint8_t* data, size_t n, size_t capacity, size_t max_read, ParseResult& out) {
    out.format = FMT_FASTQ;
    BufModel br(data, n, capacity, max_read);
    size_t start = 0, end = 0, seq = 0, sep = 0, qual = 0;      // BufferPosition (buffer-relative)
    int search_pos = 0;                                         // Id, Sequence, Separator, Quality
    uint64_t line = 1, byte = 0;
    bool finished = false;

    auto find_line = [&](size_t from, size_t& res) -> bool {    // :306-308
        if (from >= br.len()) return false;
        const void* p = std::memchr(br.buf() + from, '\n', br.len() - from);
        if (!p) return false;
        res = ((const uint8_t*)p - br.buf()) + 1;
        return true;
    };
    auto id_for_error = [&](std::string& id) -> bool {          // :287-303
        if (seq - start > 1) {
            const uint8_t* b = br.buf();
            size_t ib = start + 1, ie = trim_cr_end(b, start + 1, seq - 1), sp = ib;
            while (sp < ie && b[sp] != ' ') sp++;
            id.assign((const char*)b + ib, sp - ib);
            return true;
        }
        return false;
    };
    auto validate = [&]() -> bool {                             // :240-285
        const uint8_t* b = br.buf();
        if (b[start] != '@') { finished = true; out.err.kind = ERR_INVALID_START; out.err.line = line; return false; }
        if (b[sep] != '+') { finished = true; out.err.kind = ERR_INVALID_SEPARATOR; out.err.line = line + 2; out.err.has_id = id_for_error(out.err.id); return false; }
        size_t sl = trim_cr_end(b, seq, sep - 1) - seq, ql = trim_cr_end(b, qual, end) - qual;
        if (sl != ql) { finished = true; out.err.kind = ERR_UNEQUAL_LENGTHS; out.err.line = line; out.err.has_id = id_for_error(out.err.id); return false; }
        return true;
    };
    // :155-187 ; returns 1 complete, 0 incomplete, -1 error
    auto find = [&]() -> int {
        size_t p;
        if (!find_line(start, p)) { search_pos = 0; return 0; } seq = p;
        if (!find_line(seq, p)) { search_pos = 1; return 0; } sep = p;
        if (!find_line(sep, p)) { search_pos = 2; return 0; } qual = p;
        if (!find_line(qual, p)) { search_pos = 3; return 0; } end = p - 1;
        return validate() ? 1 : -1;
    };
    // :192-234
    auto find_incomplete = [&]() -> int {
        size_t p;
        if (search_pos == 0) { if (!find_line(start, p)) { search_pos = 0; return 0; } seq = p; }
        if (search_pos <= 1) { if (!find_line(seq, p)) { search_pos = 1; return 0; } sep = p; }
        if (search_pos <= 2) { if (!find_line(sep, p)) { search_pos = 2; return 0; } qual = p; }
        if (search_pos <= 3) { if (!find_line(qual, p)) { search_pos = 3; return 0; } end = p - 1; }
        search_pos = 0;
        return validate() ? 1 : -1;
    };
    // :337-356 ; 1 record, 0 none, -1 error
    auto check_end = [&]() -> int {
        finished = true;
        if (search_pos == 3) { end = br.len(); return validate() ? 1 : -1; }
        const uint8_t* b = br.buf();
        bool all_blank = true;
        size_t ls = start, L = br.len();
        for (;;) {
            size_t le = ls;
            while (le < L && b[le] != '\n') le++;
            if (trim_cr_end(b, ls, le) != ls) { all_blank = false; break; }
            if (le >= L) break;
            ls = le + 1;
        }
        if (all_blank) return 0;
        out.err.kind = ERR_UNEXPECTED_END; out.err.line = line + (uint64_t)search_pos;
        if (search_pos > 0) out.err.has_id = id_for_error(out.err.id);
        return -1;
    };
    // :312-333
    auto next_complete = [&]() -> int {
        for (;;) {
            if (br.len() < br.capacity()) return check_end();
            if (start == 0) { size_t c = br.capacity(); br.reserve(grow_to(c) - c); }       // grow :360-365
            else {                                                                         // make_room :368-384
                size_t consumed = start;
                br.consume(consumed); br.make_room();
                start = 0;
                if (search_pos >= 1) seq -= consumed;
                if (search_pos >= 2) sep -= consumed;
                if (search_pos >= 3) qual -= consumed;
            }
            fill_buf(br);
            int r = find_incomplete();
            if (r != 0) return r;
        }
    };

    bool is_new = true;                                          // BufferPosition::is_new(): end == 0
    for (;;) {                                                   // next() :388-449
        if (finished) break;
        if (br.len() == 0) { if (fill_buf(br) == 0) { finished = true; break; } }
        if (!is_new) { byte += end + 1 - start; line += 4; start = end + 1; }
        int r = find();
        if (r < 0) break;
        if (r == 0) { r = next_complete(); if (r <= 0) break; }
        is_new = (end == 0);
        const uint8_t* b = br.buf();
        if (out.line_ending == LE_NONE) out.line_ending = find_line_ending(b + start, end - start);
        const uint64_t off = br.stream_off();
        Record rec{};
        rec.start = off + start;
        rec.id_b = off + start + 1; rec.id_e = off + trim_cr_end(b, start + 1, seq - 1);
        rec.seq_b = off + seq; rec.seq_e = off + trim_cr_end(b, seq, sep - 1);
        rec.qual_b = off + qual; rec.qual_e = off + trim_cr_end(b, qual, end);
        rec.all_e = off + end;
        rec.num_bases = rec.seq_e - rec.seq_b;
        rec.pos_line = line; rec.pos_byte = byte;
        out.records.push_back(rec);
    }
    out.final_line = line; out.final_byte = byte;
}

// ---------------------------------------------------------------------------------------------
// ref: src/parser/fasta.rs  Reader
inline void parse_fasta_incremental(const uint8_t* data, size_t n, size_t capacity, size_t max_read, ParseResult& out) {
    out.format = FMT_FASTA;
    BufModel br(data, n, capacity, max_read);
    size_t start = 0, search_pos = 0;
    std::vector<size_t> seq_pos;
    uint64_t line = 0, byte = 0;
    bool finished = false;

    auto _find = [&]() -> bool {                                  // :220-243
        const size_t bufsize = br.len();
        const uint8_t* b = br.buf();
        size_t from = search_pos;
        while (from < bufsize) {
            const void* p = std::memchr(b + from, '\n', bufsize - from);
            if (!p) break;
            size_t pos = (const uint8_t*)p - b, next_line_start = pos + 1;
            if (next_line_start == bufsize) { search_pos = pos; return false; }
            seq_pos.push_back(pos);
            if (b[next_line_start] == '>') { search_pos = next_line_start; return true; }
            from = pos + 1;
        }
        search_pos = bufsize;
        return false;
    };
    auto find = [&]() -> bool {                                   // :200-216
        if (_find()) return true;
        if (br.len() < br.capacity()) {
            finished = true;
            if (!seq_pos.empty()) seq_pos.push_back(search_pos);
            return true;
        }
        return false;
    };
    auto next_complete = [&]() {                                  // :250-265
        for (;;) {
            if (start == 0) { size_t c = br.capacity(); br.reserve(grow_to(c) - c); }       // grow
            else {                                                                         // make_room :277-287
                size_t consumed = start;
                br.consume(consumed); br.make_room();
                start = 0; search_pos -= consumed;
                for (auto& s : seq_pos) s -= consumed;
            }
            fill_buf(br);
            if (find()) return;
        }
    };

    for (;;) {                                                    // next() :291-367
        if (finished) break;
        if (line == 0) {
            if (fill_buf(br) == 0) { finished = true; break; }
            if (br.buf()[0] == '>') { line = 1; byte = 0; start = 0; search_pos = 1; }
            else { out.err.kind = ERR_INVALID_START; out.err.line = 0; break; }
        }
        if (!seq_pos.empty()) {                                   // next_pos() :190-195
            line += seq_pos.size(); byte += search_pos - start; start = search_pos; seq_pos.clear();
        }
        if (!find()) next_complete();
        if (seq_pos.empty()) { out.err.kind = ERR_UNEXPECTED_END; out.err.line = line; break; }
        const uint8_t* b = br.buf();
        const size_t last = seq_pos.back(), firstp = seq_pos.front();
        if (out.line_ending == LE_NONE) out.line_ending = find_line_ending(b + start, last - start);
        const uint64_t off = br.stream_off();
        Record rec{};
        rec.start = off + start;
        rec.id_b = off + start + 1; rec.id_e = off + trim_cr_end(b, start + 1, firstp);
        if (seq_pos.size() > 1) { rec.seq_b = off + firstp + 1; rec.seq_e = off + trim_cr_end(b, firstp + 1, last); }
        else { rec.seq_b = rec.seq_e = off + firstp; }
        rec.all_e = off + last;
        uint64_t nb = rec.seq_e - rec.seq_b;
        for (uint64_t i = rec.seq_b; i < rec.seq_e; i++) if (data[i] == '\n' || data[i] == '\r') nb--;
        rec.num_bases = nb;
        rec.pos_line = line; rec.pos_byte = byte;
        out.records.push_back(rec);
    }
    out.final_line = line; out.final_byte = byte;
}

inline void parse_fastx_incremental(const uint8_t* buf, size_t n, size_t capacity, size_t max_read, ParseResult& out) {
    if (n < 2) { out.err.kind = ERR_EMPTY_FILE; return; }
    if (buf[0] == '>') parse_fasta_incremental(buf, n, capacity, max_read, out);
    else if (buf[0] == '@') parse_fastq_incremental(buf, n, capacity, max_read, out);
    else out.err.kind = ERR_UNKNOWN_FORMAT;
}
}  // namespace ntref
