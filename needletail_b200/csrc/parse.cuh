// parse.cuh — FASTX record scanner producing the record table (materialising path).
//
// Whole-buffer formulation of the reference's readers: delimiter flags ->
// device-wide scans (newline ordinal per byte) -> compaction of newline / record-start
// positions -> one thread per record fills the ntg_record row and validates it.
//   FASTQ: record r owns newlines 4r..4r+3          (src/parser/fastq.rs:155-187)
//   FASTA: a record starts at byte 0 and after every "\n>"   (src/parser/fasta.rs:220-243)
// Only the O(1) end-of-stream rules (fastq.rs:337-356, fasta.rs:200-216) run on the host, on the
// few newline positions the device hands back.  Part of the unity build (ntgpu.cu).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace parse {
constexpr int BLOCK = 256;
static inline unsigned grid_for(size_t n) { return (unsigned)((n + BLOCK - 1) / BLOCK); }

__global__ void __launch_bounds__(BLOCK) k_flags(const uint8_t* __restrict__ bytes, uint32_t n, int fasta,
                                                 uint8_t* __restrict__ f_nl, uint8_t* __restrict__ f_cr, uint8_t* __restrict__ f_st) {
    uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g >= n) return;
    uint8_t b = bytes[g];
    f_nl[g] = b == '\n';
    if (fasta) {
        f_cr[g] = b == '\r';
        f_st[g] = (b == '>') && (g == 0 || bytes[g - 1] == '\n');
    }
}
__global__ void __launch_bounds__(BLOCK) k_compact(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ idx, uint32_t n,
                                                   uint32_t* __restrict__ outpos) {
    uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g < n && flags[g]) outpos[idx[g]] = g;
}

__device__ __forceinline__ uint32_t trim_cr_end(const uint8_t* __restrict__ bytes, uint32_t b, uint32_t e) {
    return (e > b && bytes[e - 1] == '\r') ? e - 1 : e;      // parser/utils.rs:12-18
}

// one thread per complete FASTQ record (4 newlines): BufferPosition + validate (fastq.rs:17-65,240-285)
__global__ void __launch_bounds__(BLOCK) k_fastq_records(const uint8_t* __restrict__ bytes, const uint32_t* __restrict__ nlpos,
                                                         uint32_t n_complete, ntg_record* __restrict__ recs,
                                                         uint8_t* __restrict__ errkind, uint32_t* __restrict__ first_err) {
    uint32_t r = blockIdx.x * BLOCK + threadIdx.x;
    if (r >= n_complete) return;
    uint32_t start = r ? nlpos[4 * r - 1] + 1 : 0;
    uint32_t seq = nlpos[4 * r] + 1, sep = nlpos[4 * r + 1] + 1, qual = nlpos[4 * r + 2] + 1, end = nlpos[4 * r + 3];
    ntg_record o;
    o.start = start;
    o.id_b = start + 1; o.id_e = trim_cr_end(bytes, start + 1, seq - 1);
    o.seq_b = seq; o.seq_e = trim_cr_end(bytes, seq, sep - 1);
    o.qual_b = qual; o.qual_e = trim_cr_end(bytes, qual, end);
    o.all_e = end;
    o.num_bases = o.seq_e - o.seq_b;
    o.line = 1 + 4ull * r;                                    // fastq.rs:116,411-415
    recs[r] = o;
    uint8_t ek = 0;
    if (bytes[start] != '@') ek = NTG_EINVALID_START;
    else if (bytes[sep] != '+') ek = NTG_EINVALID_SEPARATOR;
    else if (o.seq_e - o.seq_b != o.qual_e - o.qual_b) ek = NTG_EUNEQUAL_LENGTHS;
    errkind[r] = ek;
    if (ek) atomicMin(first_err, r);
}

// one thread per FASTA record start (fasta.rs:16-108,190-195)
__global__ void __launch_bounds__(BLOCK) k_fasta_records(const uint8_t* __restrict__ bytes, uint32_t n,
                                                         const uint32_t* __restrict__ stpos, uint32_t n_starts,
                                                         const uint32_t* __restrict__ nlidx, const uint32_t* __restrict__ cridx,
                                                         const uint32_t* __restrict__ nlpos, uint32_t n_nl,
                                                         ntg_record* __restrict__ recs, uint32_t* __restrict__ last_bad,
                                                         uint32_t* __restrict__ first_le) {
    uint32_t r = blockIdx.x * BLOCK + threadIdx.x;
    if (r >= n_starts) return;
    uint32_t start = stpos[r];
    bool is_last = (r + 1 == n_starts);
    uint32_t ord = nlidx[start];                              // newlines before `start` (none at start: it is '>')
    uint32_t first_nl = ord < n_nl ? nlpos[ord] : 0xFFFFFFFFu;
    uint32_t last;
    if (!is_last) last = stpos[r + 1] - 1;                    // the '\n' in front of the next '>'
    else {
        // EOF: a newline that is the final byte is not pushed; seq_pos empty => UnexpectedEnd (fasta.rs:205-213,348-356)
        if (first_nl == 0xFFFFFFFFu || first_nl == n - 1) { *last_bad = 1; first_nl = n; }
        last = (bytes[n - 1] == '\n') ? n - 1 : n;
    }
    ntg_record o;
    o.start = start;
    o.id_b = start + 1; o.id_e = trim_cr_end(bytes, start + 1, first_nl < n ? first_nl : n);
    if (last > first_nl) { o.seq_b = first_nl + 1; o.seq_e = trim_cr_end(bytes, first_nl + 1, last); }
    else { o.seq_b = o.seq_e = (first_nl < n ? first_nl : n); }
    o.qual_b = o.qual_e = 0;
    o.all_e = last;
    uint32_t sb = (uint32_t)o.seq_b, se = (uint32_t)o.seq_e;
    o.num_bases = (uint64_t)(se - sb) - (nlidx[se] - nlidx[sb]) - (cridx[se] - cridx[sb]);   // fasta.rs:102-107
    o.line = 1 + (uint64_t)ord;
    recs[r] = o;
    // line_ending(): taken from the first record whose all() contains a newline (fasta.rs:358-360, utils.rs:106-117)
    if (first_nl < last) atomicMin(&first_le[0], r);
}
}  // namespace parse

struct RecordsPriv { PinBuf<ntg_record> recs; };

// Host-side access to a few input bytes: straight from the caller's host copy when there is one,
// else small device-to-host reads (device-resident inputs).
struct Peek {
    const uint8_t* host; const uint8_t* dev;
    bool get(uint64_t b, size_t len, uint8_t* dst) const {
        if (len == 0) return true;
        if (host) { std::memcpy(dst, host + b, len); return true; }
        return cudaMemcpy(dst, dev + b, len, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    uint8_t at(uint64_t p) const { uint8_t v = 0; get(p, 1, &v); return v; }
};
// parser/utils.rs:106-117 on all() of the first record: `first_nl` is the first newline inside it
static int host_line_ending(const Peek& pk, uint64_t b, uint64_t e, uint64_t first_nl) {
    if (first_nl < b || first_nl >= e) return NTG_LE_NONE;
    return (first_nl > b && pk.at(first_nl - 1) == '\r') ? NTG_LE_WINDOWS : NTG_LE_UNIX;
}
// ErrorPosition.id: first space-delimited token of the id (fastq.rs:287-303)
static void host_error_id(const Peek& pk, uint64_t start, uint64_t seq, ntg_parse_error* err) {
    err->has_id = 0; err->id[0] = 0;
    if (seq - start > 1) {
        uint64_t b = start + 1, e = seq - 1;
        size_t l = (size_t)(e - b);
        if (l > sizeof(err->id) - 1) l = sizeof(err->id) - 1;
        uint8_t tmp[sizeof(err->id)];
        pk.get(b, l, tmp);
        if (l == (size_t)(e - b) && l > 0 && tmp[l - 1] == '\r') l--;
        size_t sp = 0;
        while (sp < l && tmp[sp] != ' ') sp++;
        std::memcpy(err->id, tmp, sp); err->id[sp] = 0;
        err->has_id = 1;
    }
}

// One window of a stream.  `force_format` = 0: sniff (the window is the start of the stream), else the format of the stream
// this window continues.  `at_eof` = false: the stream continues behind the window - only records that are complete inside
// it are delivered (FASTQ: four newlines; FASTA: followed by another record start), no end-of-stream rule runs, and
// *consumed = offset of the first byte not covered by a delivered record (where the next window must start).  Offsets and
// line numbers are relative to the window.  keep_drecs / keep_dbytes: the caller takes over the device copies.
static int run_parse_device(ntg_ctx* ctx, const uint8_t* bytes, const uint8_t* dev_in, size_t n, ntg_records** out,
                            DevBuf<ntg_record>* keep_drecs, int force_format = 0, bool at_eof = true, uint64_t* consumed = nullptr,
                            DevBuf<uint8_t>* keep_dbytes = nullptr) {
    using namespace parse;
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output pointer");
    *out = nullptr;
    if (n && !bytes && !dev_in) return ntg_set_error(ctx, NTG_EINVAL, "null input");
    const Peek pk{bytes, dev_in};
    if (n >= 0xFFFFFFF0ull) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "ntg_parse_fastx: feed at most 4 GiB per call");
    auto* res = new ntg_records();
    auto* priv = new RecordsPriv();
    std::memset(res, 0, sizeof(*res));
    res->_priv = priv;
    auto done = [&](int st) { if (st != NTG_OK) { delete priv; delete res; } else *out = res; return st; };

    if (consumed) *consumed = 0;
    if (force_format) {
        res->format = force_format;
        if (n == 0) return done(NTG_OK);
    } else {
        // sniff: parse_fastx_reader / get_fastx_reader (parser/mod.rs:85-93,37-46)
        if (n < 2) { res->error.kind = NTG_EEMPTY_FILE; return done(NTG_OK); }
        const uint8_t b0 = pk.at(0);
        if (b0 == '>') res->format = NTG_FMT_FASTA;
        else if (b0 == '@') res->format = NTG_FMT_FASTQ;
        else { res->error.kind = NTG_EUNKNOWN_FORMAT; return done(NTG_OK); }
    }
    res->error.format = res->format;
    const bool fasta = res->format == NTG_FMT_FASTA;
    const uint32_t n32 = (uint32_t)n;

    DevBuf<uint8_t> dbytes_local, f_nl, f_cr, f_st;
    DevBuf<uint8_t>& dbytes_own = keep_dbytes ? *keep_dbytes : dbytes_local;
    DevBuf<uint32_t> nlidx, cridx, stidx, tmp, nlpos, stpos;
    struct { const uint8_t* p; } dbytes{dev_in};
    if (!dev_in) {
        if (dbytes_own.alloc(n)) return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
        dbytes.p = dbytes_own.p;
    }
    if (f_nl.alloc(n) || nlidx.alloc(n + 1) || tmp.alloc(scan_tmp_count(n)))
        return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
    if (fasta && (f_cr.alloc(n) || f_st.alloc(n) || cridx.alloc(n + 1) || stidx.alloc(n + 1)))
        return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
#define PCUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return done(ntg_set_error(ctx, NTG_ECUDA, "%s:%d %s", __FILE__, __LINE__, cudaGetErrorString(_e))); } while (0)
#define PTRY(expr) do { int _s = (expr); if (_s != NTG_OK) return done(_s); } while (0)
    if (!dev_in) PCUDA(cudaMemcpyAsync(dbytes_own.p, bytes, n, cudaMemcpyHostToDevice, ctx->stream));
    k_flags<<<grid_for(n), BLOCK, 0, ctx->stream>>>(dbytes.p, n32, fasta ? 1 : 0, f_nl.p, f_cr.p, f_st.p);
    ctx->launches++;
    PTRY(exclusive_scan_u8(ctx, f_nl.p, nlidx.p, n, tmp.p));
    uint32_t n_nl = 0, n_st = 0;
    PCUDA(cudaMemcpyAsync(&n_nl, nlidx.p + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (fasta) {
        PTRY(exclusive_scan_u8(ctx, f_cr.p, cridx.p, n, tmp.p));
        PTRY(exclusive_scan_u8(ctx, f_st.p, stidx.p, n, tmp.p));
        PCUDA(cudaMemcpyAsync(&n_st, stidx.p + n, 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PCUDA(cudaStreamSynchronize(ctx->stream));
    if (nlpos.alloc(n_nl) || (fasta && stpos.alloc(n_st))) return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
    k_compact<<<grid_for(n), BLOCK, 0, ctx->stream>>>(f_nl.p, nlidx.p, n32, nlpos.p);
    ctx->launches++;
    if (fasta) { k_compact<<<grid_for(n), BLOCK, 0, ctx->stream>>>(f_st.p, stidx.p, n32, stpos.p); ctx->launches++; }

    DevBuf<ntg_record> drecs_own;
    DevBuf<ntg_record>& drecs = keep_drecs ? *keep_drecs : drecs_own;
    uint32_t le_rec = 0;     // index of the first record whose all() contains a newline (FASTQ: always record 0)
    if (!fasta) {
        // -------------------------------------------------------------------------- FASTQ
        uint32_t n_complete = n_nl / 4, rem = n_nl % 4;
        DevBuf<uint8_t> errkind; DevBuf<uint32_t> first_err;
        if (drecs.alloc((size_t)n_complete + 1) || errkind.alloc(n_complete) || first_err.alloc(1))      // (+1: a last record without newline)
            return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
        PCUDA(cudaMemsetAsync(first_err.p, 0xFF, 4, ctx->stream));
        if (n_complete) {
            k_fastq_records<<<grid_for(n_complete), BLOCK, 0, ctx->stream>>>(dbytes.p, nlpos.p, n_complete, drecs.p, errkind.p, first_err.p);
            ctx->launches++;
        }
        PCUDA(cudaGetLastError());
        uint32_t ferr = 0xFFFFFFFFu;
        uint32_t tailnl[4] = {0, 0, 0, 0};     // [0] = last newline of the last complete record, [1..rem] = trailing newlines
        PCUDA(cudaMemcpyAsync(&ferr, first_err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (n_complete) PCUDA(cudaMemcpyAsync(&tailnl[0], nlpos.p + (4 * (size_t)n_complete - 1), 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (rem) PCUDA(cudaMemcpyAsync(&tailnl[1], nlpos.p + 4 * (size_t)n_complete, 4 * rem, cudaMemcpyDeviceToHost, ctx->stream));
        PCUDA(cudaStreamSynchronize(ctx->stream));

        uint64_t n_ok = n_complete;
        ntg_record eof_rec; bool have_eof_rec = false;
        if (ferr != 0xFFFFFFFFu) {
            // first invalid complete record ends the stream (fastq.rs:243,253,277)
            n_ok = ferr;
            uint8_t ek = 0; ntg_record bad;
            PCUDA(cudaMemcpy(&ek, errkind.p + ferr, 1, cudaMemcpyDeviceToHost));
            PCUDA(cudaMemcpy(&bad, drecs.p + ferr, sizeof(bad), cudaMemcpyDeviceToHost));
            res->error.kind = ek; res->error.record_index = ferr;
            res->error.line = 1 + 4ull * ferr + (ek == NTG_EINVALID_SEPARATOR ? 2 : 0);     // fastq.rs:246,256,281
            if (ek != NTG_EINVALID_START) host_error_id(pk, bad.start, bad.seq_b, &res->error);
        } else if (!at_eof) {
            if (consumed) *consumed = n_complete ? (uint64_t)tailnl[0] + 1 : 0;
        } else {
            // end of stream: check_end (fastq.rs:337-356)
            uint64_t start = n_complete ? (uint64_t)tailnl[0] + 1 : 0;
            uint64_t line = 1 + 4ull * n_complete;
            if (rem == 3) {
                uint64_t seq = (uint64_t)tailnl[1] + 1, sep = (uint64_t)tailnl[2] + 1, qual = (uint64_t)tailnl[3] + 1, end = n;
                auto trim = [&](uint64_t b, uint64_t e) { return (e > b && pk.at(e - 1) == '\r') ? e - 1 : e; };
                ntg_record o;
                o.start = start; o.id_b = start + 1; o.id_e = trim(start + 1, seq - 1);
                o.seq_b = seq; o.seq_e = trim(seq, sep - 1); o.qual_b = qual; o.qual_e = trim(qual, end);
                o.all_e = end; o.num_bases = o.seq_e - o.seq_b; o.line = line;
                int ek = 0;
                if (pk.at(start) != '@') ek = NTG_EINVALID_START;
                else if (pk.at(sep) != '+') ek = NTG_EINVALID_SEPARATOR;
                else if (o.seq_e - o.seq_b != o.qual_e - o.qual_b) ek = NTG_EUNEQUAL_LENGTHS;
                if (ek) {
                    res->error.kind = ek; res->error.record_index = n_complete;
                    res->error.line = line + (ek == NTG_EINVALID_SEPARATOR ? 2 : 0);
                    if (ek != NTG_EINVALID_START) host_error_id(pk, start, seq, &res->error);
                } else { eof_rec = o; have_eof_rec = true; }
            } else if (start < n || rem) {
                // leftover must consist solely of empty / "\r" lines (fastq.rs:346-350)
                bool blank = true;
                uint64_t ls = start;
                for (uint32_t i = 0; i <= rem && blank; i++) {
                    uint64_t le = (i < rem) ? (uint64_t)tailnl[1 + i] : n;
                    uint64_t len = le - ls;
                    if (len > 1 || (len == 1 && pk.at(ls) != '\r')) blank = false;
                    ls = le + 1;
                }
                if (!blank) {
                    res->error.kind = NTG_EUNEXPECTED_END; res->error.record_index = n_complete;
                    res->error.line = line + rem;                       // search_pos as line offset (fastq.rs:352-355)
                    if (rem > 0) host_error_id(pk, start, (uint64_t)tailnl[1] + 1, &res->error);
                }
            }
        }
        uint64_t total = n_ok + (have_eof_rec ? 1 : 0);
        if (priv->recs.alloc(total)) return done(ntg_set_error(ctx, NTG_ENOMEM, "pinned allocation failed"));
        if (n_ok) PCUDA(cudaMemcpy(priv->recs.p, drecs.p, n_ok * sizeof(ntg_record), cudaMemcpyDeviceToHost));
        if (have_eof_rec) {
            priv->recs.p[n_ok] = eof_rec;
            PCUDA(cudaMemcpy(drecs.p + n_ok, &eof_rec, sizeof(eof_rec), cudaMemcpyHostToDevice));      // (callers that keep the device table)
        }
        res->n_records = total; res->records = priv->recs.p;
        // FastxReader::position() after the last next(): position of the last record attempted (fastq.rs:411-415)
        {
            const uint64_t idx = (ferr != 0xFFFFFFFFu) ? ferr : n_complete;
            res->final_line = 1 + 4ull * idx;
            res->final_byte = (idx == 0) ? 0 : (idx <= n_ok ? priv->recs.p[idx - 1].all_e + 1 : (uint64_t)tailnl[0] + 1);
        }
    } else {
        // -------------------------------------------------------------------------- FASTA
        DevBuf<uint32_t> last_bad, first_le;
        if (drecs.alloc(n_st) || last_bad.alloc(1) || first_le.alloc(1)) return done(ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed"));
        PCUDA(cudaMemsetAsync(last_bad.p, 0, 4, ctx->stream));
        PCUDA(cudaMemsetAsync(first_le.p, 0xFF, 4, ctx->stream));
        k_fasta_records<<<grid_for(n_st), BLOCK, 0, ctx->stream>>>(dbytes.p, n32, stpos.p, n_st, nlidx.p, cridx.p, nlpos.p, n_nl, drecs.p, last_bad.p, first_le.p);
        ctx->launches++;
        PCUDA(cudaGetLastError());
        uint32_t bad = 0;
        PCUDA(cudaMemcpyAsync(&bad, last_bad.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PCUDA(cudaMemcpyAsync(&le_rec, first_le.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PCUDA(cudaStreamSynchronize(ctx->stream));
        uint64_t total = n_st - (bad ? 1 : 0);
        if (priv->recs.alloc(n_st)) return done(ntg_set_error(ctx, NTG_ENOMEM, "pinned allocation failed"));
        PCUDA(cudaMemcpy(priv->recs.p, drecs.p, (size_t)n_st * sizeof(ntg_record), cudaMemcpyDeviceToHost));
        res->n_records = total; res->records = priv->recs.p;
        if (!at_eof) {
            // the last record start of the window is not delivered: the record may continue behind the window
            const ntg_record& l = priv->recs.p[n_st - 1];
            res->n_records = n_st - 1;
            res->final_line = l.line; res->final_byte = l.start;
            if (consumed) *consumed = l.start;
        } else if (bad) {
            const ntg_record& b = priv->recs.p[n_st - 1];
            res->error.kind = NTG_EUNEXPECTED_END; res->error.record_index = n_st - 1;
            res->error.line = b.line;                                   // fasta.rs:348-356
            res->final_line = b.line; res->final_byte = b.start;
        } else {
            const ntg_record& l = priv->recs.p[n_st - 1];
            res->final_line = l.line; res->final_byte = l.start;
        }
    }
#undef PCUDA
#undef PTRY
    if (le_rec < res->n_records) {
        // the first newline inside that record is the one ending its header line
        const ntg_record& lr = res->records[le_rec];
        uint64_t q = lr.id_e;
        if (pk.at(q) == '\r') q++;                               // id() had a '\r' trimmed
        res->line_ending = host_line_ending(pk, lr.start, lr.all_e, q);
    }
    return done(NTG_OK);
}
