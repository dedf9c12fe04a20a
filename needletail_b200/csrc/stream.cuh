// stream.cuh — the streaming feeder of the fused tallies path (unity build).
//
//   ntg_stream  : a tally session over a byte stream of unknown length, fed in arbitrary pieces (the reference's
//                 `parse_fastx_reader<R: Read>`, src/parser/mod.rs:85-150).  Pieces are staged in a ring of pinned host
//                 buffers; full buffers go to the device as segments (SegmentFeed, fused_host.cuh) while the next one fills.
//   GzInflater  : host-side inflate in front of a session.  Multi-member gzip = flate2::MultiGzDecoder (mod.rs:98); BGZF
//                 files (block sizes in the member headers) are inflated block-parallel by a thread pool straight into
//                 the pinned staging buffer.
// Errors: a stream cannot be re-read, so a flagged launch is resolved while its segment (and its neighbours) are still
// resident: speculation misses are re-run in place, parse errors replay the launch truncated at the failing record and the
// stream ends there (iterator semantics); the error itself is classified at finish from the bytes behind the failing record.
#pragma once
#include <zlib.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "fused_host.cuh"
#include "inflate.cuh"

constexpr int NPIN = 4;                                        // pinned staging buffers (launches L-3 .. L-1 stay readable while L fills)
constexpr size_t ERRWIN_CAP = size_t(64) << 20;                // bytes kept behind a failing record for its classification

struct ntg_stream {
    ntg_ctx* ctx = nullptr;
    ntg_tally_config cfg{};
    SegmentFeed feed;
    bool opened = false, finished = false;
    int format = 0;
    uint32_t TB = 0;
    uint8_t* pin[NPIN] = {};
    size_t cap = 0, fill = 0;
    uint64_t submitted_bytes = 0;                              // stream bytes handed to launches so far
    uint64_t total_fed = 0;
    // results
    PassResult total;
    unsigned long long hist[4][16] = {};                       // tallies of the launches accumulated most recently (by launch index & 3)
    uint32_t spec_missed = 0;
    // first parse error
    bool failed = false, unsupported = false;
    uint32_t unsupported_flags = 0;
    uint64_t err_E = 0, err_recs_before = 0;
    bool fasta_end_error = false;
    std::vector<uint8_t> errwin; bool errwin_capped = false;
    bool in_submit = false;                                     // a check runs inside a submit: the staging buffer is not a launch yet
    std::string io_error;                                       // inflate failures (ParseErrorKind::Io)
    struct GzState* gz = nullptr;
    struct DevGz* devgz = nullptr;                              // device-side inflate (BGZF): the text never exists on the host
};
static void gz_free(struct GzState* g);
static void devgz_free(struct DevGz* d);

static void stream_free(ntg_stream* s) {
    if (!s) return;
    if (s->ctx) cudaSetDevice(s->ctx->device);
    if (s->opened && s->ctx && s->ctx->fused) cudaStreamSynchronize(s->ctx->stream);
    for (auto& p : s->pin) if (p) cudaFreeHost(p);
    gz_free(s->gz);
    devgz_free(s->devgz);
    delete s;
}

static int stream_create(ntg_ctx* ctx, const ntg_tally_config* cfg, ntg_stream** out) {
    NTG_TRY(check_tally_cfg(ctx, cfg));
    NTG_TRY(fused_init(ctx));
    auto* s = new ntg_stream();
    s->ctx = ctx; s->cfg = *cfg;
    s->cap = STREAM_SEG + fused::TILE;
    *out = s;
    return NTG_OK;
}
static int stream_need_staging(ntg_stream* s) {                 // the pinned ring of text staging buffers, on first use
    if (s->pin[0]) return NTG_OK;
    for (auto& p : s->pin)
        if (cudaMallocHost((void**)&p, s->cap) != cudaSuccess) { cudaGetLastError(); return ntg_set_error(s->ctx, NTG_ENOMEM, "pinned staging allocation failed"); }
    return NTG_OK;
}

// ---- resolution of a flagged launch (see the header comment) --------------------------------------------------------
static int stream_redo(ntg_stream* s, uint64_t q, uint64_t n_vis, bool final, bool spec, LaunchCtl* res) {
    ntg_ctx* ctx = s->ctx; FusedState* st = ctx->fused;
    const SegmentFeed::Rec& r = s->feed.recs[q % (NCTL - 1)];
    const uint64_t start = r.tb * (uint64_t)s->TB;
    const uint64_t te = final ? (n_vis + s->TB - 1) / s->TB : r.te;
    const uint32_t spec_saved = st->P.spec;
    st->P.spec = spec ? spec_saved : 0;
    const uint8_t* base = s->feed.st->seg[q % NSEG] + STREAM_BACK - start;
    int rc = fused_enqueue_launch(ctx, base, start > STREAM_BACK ? start - STREAM_BACK : 0, n_vis, r.tb, te, final, CTL_REDO);
    st->P.spec = spec_saved;
    NTG_TRY(rc);
    NTG_CUDA(ctx, cudaEventSynchronize(st->ev_done[CTL_REDO]));
    *res = st->h_ctl[CTL_REDO];
    return NTG_OK;
}

static void stream_collect_errwin(ntg_stream* s, uint64_t q) {
    // bytes from the failing record to the end of what has been fed: the resident launches q .. L-1, then the buffer being filled
    s->errwin.clear(); s->errwin_capped = false;
    auto append = [&](const uint8_t* p, size_t n) {
        if (s->errwin.size() + n > ERRWIN_CAP) { n = ERRWIN_CAP - s->errwin.size(); s->errwin_capped = true; }
        s->errwin.insert(s->errwin.end(), p, p + n);
    };
    for (uint64_t i = q; i < s->feed.L; i++) {
        const SegmentFeed::Rec& r = s->feed.recs[i % (NCTL - 1)];
        const uint64_t b0 = r.tb * (uint64_t)s->TB;
        const uint64_t from = s->err_E > b0 ? s->err_E - b0 : 0;
        const size_t len = r.final ? (size_t)(r.n_vis - b0) : r.len;
        if (from >= len) continue;
        if (!s->devgz) { append(s->pin[i % NPIN] + from, len - (size_t)from); continue; }
        size_t n = len - (size_t)from;                          // device-side inflate: the text only exists in the device segments
        if (s->errwin.size() + n > ERRWIN_CAP) { n = ERRWIN_CAP - s->errwin.size(); s->errwin_capped = true; }
        const size_t old = s->errwin.size();
        s->errwin.resize(old + n);
        cudaMemcpy(s->errwin.data() + old, s->feed.device_addr(i, b0 + from), n, cudaMemcpyDeviceToHost);
    }
    if (s->in_submit && !s->devgz) append(s->pin[s->feed.L % NPIN], s->fill);
    if (s->devgz && !s->finished) s->errwin_capped = true;      // (text behind the resident launches is not available: treat as cut)
}

static int stream_check(ntg_stream* s, uint64_t j, const LaunchCtl& c_in) {
    if (s->failed || s->unsupported) return NTG_OK;             // the stream has ended at an earlier launch
    LaunchCtl c = c_in;
    if (c.flags & fused::FLAG_SPEC_MISS) {                      // re-run this launch without speculation, in place
        const SegmentFeed::Rec& r = s->feed.recs[j % (NCTL - 1)];
        s->spec_missed |= c.flags;
        NTG_TRY(stream_redo(s, j, r.n_vis, r.final, false, &c));
    }
    if (c.flags == 0 || (c.flags == fused::FLAG_PARSE_ERROR && s->format == NTG_FMT_FASTA)) {
        s->total.add(c);
        std::memcpy(s->hist[j & 3], c.tallies, sizeof(c.tallies));
        if (c.flags) { s->failed = true; s->fasta_end_error = true; }      // (end-of-stream rule of the last record, fasta.rs:348-356)
        return NTG_OK;
    }
    if (c.flags != fused::FLAG_PARSE_ERROR || c.err_key == ~0ull) { s->unsupported = true; s->unsupported_flags = c.flags; return NTG_OK; }
    // FASTQ parse error: the stream ends in front of the failing record
    const uint64_t E = c.err_key >> 2;
    s->failed = true; s->err_E = E;
    uint64_t q = j;
    if (E == 0) {                                               // the first record of the stream
        if (j > 1) { s->unsupported = true; s->unsupported_flags = c.flags; return NTG_OK; }
        s->total = PassResult{}; s->err_recs_before = 0;
        stream_collect_errwin(s, 0);
        return NTG_OK;
    }
    const SegmentFeed::Rec& rj = s->feed.recs[j % (NCTL - 1)];
    if (E <= rj.tb * (uint64_t)s->TB) {                         // the record began in the previous launch: that one is replayed
        if (j == 0) { s->unsupported = true; s->unsupported_flags = c.flags; return NTG_OK; }
        q = j - 1;
        const SegmentFeed::Rec& rq = s->feed.recs[q % (NCTL - 1)];
        if (E <= rq.tb * (uint64_t)s->TB) { s->unsupported = true; s->unsupported_flags = c.flags; return NTG_OK; }   // a record longer than a segment
        for (int i = 0; i < 16; i++) s->total.tallies[i] -= s->hist[q & 3][i];
    }
    LaunchCtl t;
    NTG_TRY(stream_redo(s, q, E, true, true, &t));
    if (t.flags & fused::FLAG_SPEC_MISS) NTG_TRY(stream_redo(s, q, E, true, false, &t));
    if (t.flags) { s->unsupported = true; s->unsupported_flags = t.flags; return NTG_OK; }
    s->total.add(t);
    s->err_recs_before = s->total.tallies[0];
    stream_collect_errwin(s, q);
    return NTG_OK;
}

static int stream_open_pass(ntg_stream* s, const uint8_t* sample, size_t ns) {
    s->format = sample[0] == '>' ? NTG_FMT_FASTA : NTG_FMT_FASTQ;
    s->TB = pick_tile_bytes(sample, ns < 65536 ? ns : 65536, s->format);
    NTG_TRY(s->feed.open(s->ctx, s->format, &s->cfg, s->TB, true));
    s->opened = true;
    return NTG_OK;
}

// the staging buffer is full: launch its whole tiles (all but at least one byte: the last tile of a stream needs its length)
static int stream_submit_full(ntg_stream* s) {
    if (s->failed || s->unsupported) {                           // the tallies are final: only the error window still grows
        if (s->failed && !s->fasta_end_error && !s->errwin_capped) {
            size_t n = s->fill;
            if (s->errwin.size() + n > ERRWIN_CAP) { n = ERRWIN_CAP - s->errwin.size(); s->errwin_capped = true; }
            s->errwin.insert(s->errwin.end(), s->pin[s->feed.L % NPIN], s->pin[s->feed.L % NPIN] + n);
        }
        s->fill = 0;
        return NTG_OK;
    }
    uint8_t* cur = s->pin[s->feed.L % NPIN];
    if (!s->opened) {
        if (cur[0] != '>' && cur[0] != '@') { s->unsupported = true; s->unsupported_flags = fused::FLAG_FORMAT; s->fill = 0; return NTG_OK; }
        NTG_TRY(stream_open_pass(s, cur, s->fill));
    }
    uint64_t ntiles = (s->fill - 1) / s->TB;
    if (ntiles > s->feed.seg_tiles()) ntiles = s->feed.seg_tiles();
    const size_t len = (size_t)(ntiles * (uint64_t)s->TB), carry = s->fill - len;
    uint8_t* next = s->pin[(s->feed.L + 1) % NPIN];
    auto check = [&](uint64_t j, const LaunchCtl& c) { return stream_check(s, j, c); };
    s->in_submit = true;
    const int rc = s->feed.submit(cur, len, false, 0, check);    // (blocks until launch L-2 has been checked)
    s->in_submit = false;
    NTG_TRY(rc);
    if (s->failed || s->unsupported) { s->fill = 0; return NTG_OK; }   // (a failure found meanwhile: the error window already holds these bytes)
    std::memcpy(next, cur + len, carry);
    s->submitted_bytes += len; s->fill = carry;
    return NTG_OK;
}

static int stream_acquire(ntg_stream* s, uint8_t** ptr, size_t* avail) {
    if (s->finished) return ntg_set_error(s->ctx, NTG_EINVAL, "stream already finished");
    if (s->devgz) return ntg_set_error(s->ctx, NTG_EINVAL, "this session inflates on the device: feed it compressed bytes only");
    NTG_TRY(stream_need_staging(s));
    if (s->fill == s->cap) NTG_TRY(stream_submit_full(s));
    *ptr = s->pin[s->feed.L % NPIN] + s->fill;
    *avail = s->cap - s->fill;
    return NTG_OK;
}
static int stream_commit(ntg_stream* s, size_t n) {
    if (n > s->cap - s->fill) return ntg_set_error(s->ctx, NTG_EINVAL, "commit larger than the acquired space");
    s->fill += n; s->total_fed += n;
    if (s->fill == s->cap) NTG_TRY(stream_submit_full(s));
    return NTG_OK;
}
static int stream_feed(ntg_stream* s, const uint8_t* bytes, size_t n) {
    while (n) {
        uint8_t* p; size_t avail;
        NTG_TRY(stream_acquire(s, &p, &avail));
        const size_t take = n < avail ? n : avail;
        std::memcpy(p, bytes, take);
        NTG_TRY(stream_commit(s, take));
        bytes += take; n -= take;
    }
    return NTG_OK;
}

static int gz_finish(ntg_stream* s);
static int devgz_finish(ntg_stream* s);
static int stream_finish(ntg_stream* s, ntg_tallies* out, ntg_parse_error* err) {
    ntg_ctx* ctx = s->ctx;
    if (s->finished) return ntg_set_error(ctx, NTG_EINVAL, "stream already finished");
    if (out) std::memset(out, 0, sizeof(*out));
    if (err) std::memset(err, 0, sizeof(*err));
    struct Done { ntg_stream* s; ~Done() { s->finished = true; } } done{s};
    NTG_TRY(gz_finish(s));
    if (s->devgz) NTG_TRY(devgz_finish(s));
    if (!s->io_error.empty()) { if (err) err->kind = NTG_EIO; return ntg_set_error(ctx, NTG_OK, "%s", s->io_error.c_str()); }
    if (!s->devgz) NTG_TRY(stream_need_staging(s));
    uint8_t* cur = s->pin[s->feed.L % NPIN];
    if (!s->opened && !s->unsupported) {
        // a stream shorter than one staging buffer: the sniff rules of parse_fastx_reader (mod.rs:85-93,37-46)
        int format;
        if (sniff_format(ctx, s->fill ? cur[0] : 0, s->fill, out, err, &format)) return NTG_OK;
        NTG_TRY(stream_open_pass(s, cur, s->fill));
    }
    if (s->unsupported && s->unsupported_flags == fused::FLAG_FORMAT && !s->opened) { if (err) err->kind = NTG_EUNKNOWN_FORMAT; return NTG_OK; }
    if (err) err->format = s->format;
    auto check = [&](uint64_t j, const LaunchCtl& c) { return stream_check(s, j, c); };
    if (s->devgz) {
        // (devgz_finish has launched the last batch)
    } else if (!s->failed && !s->unsupported) {
        const uint64_t n_total = s->submitted_bytes + s->fill;
        s->in_submit = true;
        const int rc = s->feed.submit(cur, s->fill, true, n_total, check);
        s->in_submit = false;
        NTG_TRY(rc);
    } else if (s->failed && !s->fasta_end_error && !s->errwin_capped) {
        size_t n = s->fill;
        if (s->errwin.size() + n > ERRWIN_CAP) { n = ERRWIN_CAP - s->errwin.size(); s->errwin_capped = true; }
        s->errwin.insert(s->errwin.end(), cur, cur + n);
    }
    NTG_TRY(s->feed.drain(check));
    if (s->unsupported)
        return ntg_set_error(ctx, NTG_EUNSUPPORTED, "this stream needs the exact record-table path (flags %u): use ntg_tally_fastx on the whole input", s->unsupported_flags);
    if (out) { tallies_from_pass(s->total, out); out->reserved[1] = s->spec_missed; }
    if (s->failed && err) {
        if (s->fasta_end_error) { err->kind = (int32_t)s->total.fin[0]; err->line = s->total.fin[1]; err->record_index = s->total.fin[2]; }
        else {
            bool confirmed = false;
            ByteSource src{s->errwin.data(), nullptr, s->errwin.size(), s->errwin_capped};
            ntg_parse_error e2; std::memset(&e2, 0, sizeof(e2)); e2.format = s->format;
            NTG_TRY(classify_error_at(ctx, src, 0, s->err_recs_before, &e2, &confirmed));
            if (!confirmed) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "the failing record could not be classified (longer than the error window?)");
            *err = e2;
        }
    }
    return NTG_OK;
}

// ============================================================================================ gzip in front of a session
// raw-deflate payload of one complete BGZF member -> dst (exactly isize bytes)
static bool bgzf_inflate_member(const uint8_t* p, size_t csize, uint8_t* dst, uint32_t isize) {
    size_t hdr, plen;
    if (!bgzf_payload(p, csize, &hdr, &plen)) return false;
    z_stream z; std::memset(&z, 0, sizeof(z));
    if (inflateInit2(&z, -15) != Z_OK) return false;
    z.next_in = const_cast<Bytef*>(p + hdr); z.avail_in = (uInt)plen;
    z.next_out = dst; z.avail_out = isize;
    const int rc = isize ? inflate(&z, Z_FINISH) : Z_STREAM_END;
    const bool ok = rc == Z_STREAM_END && z.total_out == isize;
    inflateEnd(&z);
    return ok;
}

struct GzState {
    bool decided = false, bgzf = false, z_init = false, between_members = false, ended = false;
    z_stream z;
    std::vector<uint8_t> carry;                 // BGZF: an incomplete member from the previous piece
    uint64_t out_bytes = 0;
};
static void gz_free(GzState* g) {
    if (!g) return;
    if (g->z_init) inflateEnd(&g->z);
    delete g;
}

// sequential multi-member inflate (flate2::MultiGzDecoder, src/parser/mod.rs:98) of one more piece of the compressed stream
static int gz_feed_sequential(ntg_stream* s, const uint8_t* in, size_t n) {
    GzState* g = s->gz;
    if (!g->z_init) {
        std::memset(&g->z, 0, sizeof(g->z));
        if (inflateInit2(&g->z, 15 + 32) != Z_OK) return ntg_set_error(s->ctx, NTG_ENOMEM, "inflateInit2 failed");
        g->z_init = true;
    }
    size_t off = 0;
    while (off < n && !g->ended && s->io_error.empty()) {
        if (g->between_members) {                               // another member, zero padding, or garbage
            if (in[off] == 0x1f) { inflateReset(&g->z); g->between_members = false; }
            else if (in[off] == 0) { g->ended = true; break; }
            else { s->io_error = "invalid gzip header after a member"; break; }
        }
        uint8_t* p; size_t avail;
        NTG_TRY(stream_acquire(s, &p, &avail));
        const size_t in_take = n - off < (size_t(1) << 30) ? n - off : (size_t(1) << 30);
        g->z.next_in = const_cast<Bytef*>(in + off); g->z.avail_in = (uInt)in_take;
        g->z.next_out = p; g->z.avail_out = (uInt)(avail < (size_t(1) << 30) ? avail : (size_t(1) << 30));
        const uInt out0 = g->z.avail_out;
        const int rc = inflate(&g->z, Z_NO_FLUSH);
        off += in_take - g->z.avail_in;
        g->out_bytes += out0 - g->z.avail_out;
        NTG_TRY(stream_commit(s, out0 - g->z.avail_out));
        if (rc == Z_STREAM_END) g->between_members = true;
        else if (rc != Z_OK && rc != Z_BUF_ERROR) s->io_error = std::string("inflate: ") + (g->z.msg ? g->z.msg : "data error");
    }
    return NTG_OK;
}

// BGZF: batches of complete members sized to the acquired staging space, inflated by `threads` workers in place
static int gz_feed_bgzf(ntg_stream* s, const uint8_t* in, size_t n, int threads) {
    GzState* g = s->gz;
    struct Job { const uint8_t* p; size_t csize; uint8_t* dst; uint32_t isize; };
    std::vector<Job> jobs;
    size_t off = 0;
    if (!g->carry.empty()) {
        // complete the member that the previous piece left unfinished
        for (;;) {
            const long ms = bgzf_member_size(g->carry.data(), g->carry.size());
            if (ms == 0) { s->io_error = "not a BGZF member where one was expected"; return NTG_OK; }
            const size_t need = ms < 0 ? g->carry.size() + 64 : (size_t)ms;
            if (g->carry.size() >= need && ms > 0) break;
            const size_t take = need - g->carry.size() < n - off ? need - g->carry.size() : n - off;
            if (take == 0) return NTG_OK;                       // still incomplete: wait for the next piece
            g->carry.insert(g->carry.end(), in + off, in + off + take);
            off += take;
        }
        const size_t csize = g->carry.size();
        const uint32_t isize = bgzf_isize(g->carry.data(), csize);
        for (;;) {
            uint8_t* p; size_t avail;
            NTG_TRY(stream_acquire(s, &p, &avail));
            if (isize > avail) { if (avail == s->cap) return ntg_set_error(s->ctx, NTG_EUNSUPPORTED, "BGZF member larger than the staging buffer"); NTG_TRY(stream_submit_full(s)); continue; }
            if (!bgzf_inflate_member(g->carry.data(), csize, p, isize)) { s->io_error = "inflate: corrupt BGZF member"; return NTG_OK; }
            g->out_bytes += isize;
            NTG_TRY(stream_commit(s, isize));
            break;
        }
        g->carry.clear();
    }
    while (off < n && s->io_error.empty()) {
        uint8_t* p; size_t avail;
        NTG_TRY(stream_acquire(s, &p, &avail));
        jobs.clear();
        size_t used = 0;
        bool partial = false;
        while (off < n) {
            const long ms = bgzf_member_size(in + off, n - off);
            if (ms == 0) { s->io_error = "not a BGZF member where one was expected"; break; }
            if (ms < 0 || (size_t)ms > n - off) { partial = true; break; }
            const uint32_t is = bgzf_isize(in + off, (size_t)ms);
            if (used + is > avail) break;
            jobs.push_back(Job{in + off, (size_t)ms, p + used, is});
            used += is; off += (size_t)ms;
        }
        if (!s->io_error.empty()) break;
        if (jobs.empty()) {
            if (partial) { g->carry.assign(in + off, in + n); off = n; break; }
            if (avail == s->cap) return ntg_set_error(s->ctx, NTG_EUNSUPPORTED, "BGZF member larger than the staging buffer");
            NTG_TRY(stream_submit_full(s));                     // the rest of the buffer is too small for the next member
            continue;
        }
        std::atomic<size_t> next{0}; std::atomic<bool> bad{false};
        auto work = [&]() {
            for (size_t i; (i = next.fetch_add(1)) < jobs.size();)
                if (!bgzf_inflate_member(jobs[i].p, jobs[i].csize, jobs[i].dst, jobs[i].isize)) bad = true;
        };
        std::vector<std::thread> pool;
        const int nt = (int)(jobs.size() < (size_t)threads ? jobs.size() : (size_t)threads);
        for (int t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
        if (bad) { s->io_error = "inflate: corrupt BGZF member"; break; }
        g->out_bytes += used;
        NTG_TRY(stream_commit(s, used));
        if (partial) { g->carry.assign(in + off, in + n); off = n; }
    }
    return NTG_OK;
}


// ============================================================================================ device-side inflate (BGZF)
// The host only walks the member headers.  The compressed bytes of every piece go to the device as they are fed (one H2D per
// run of members, straight from the caller's buffer: pin it for full PCIe speed), the member table follows per batch of
// ~DEVGZ_BATCH text bytes, and gzdev::k_inflate writes the text into the next device segment behind the bytes left over from
// the previous launch (a launch covers whole tiles); then the fused kernel runs.  A batch is large because one THREAD inflates
// one member: 2 GiB of text are 32 768 members, about the number of threads the decoder keeps resident.
constexpr size_t DEVGZ_BATCH = size_t(2) << 30;
struct DevGz {
    size_t comp_cap = 0; uint32_t mem_cap = 0;
    uint8_t* d_comp[2] = {}; gzdev::Member* d_mem[2] = {}; gzdev::Member* h_mem[2] = {};     // h_mem pinned
    uint32_t* d_err = nullptr;
    cudaStream_t gz_stream = nullptr;                            // H2D of compressed bytes (the caller's buffer is free again when a feed returns)
    cudaEvent_t ev_free[2] = {}, ev_h2d = nullptr;
    bool busy[2] = {false, false};
    int cur = 0;
    size_t comp_fill = 0; uint32_t n_mem = 0;
    uint64_t text_fill = 0;                                      // text bytes the staged members will produce
    size_t carry = 0;                                            // text bytes left over from the previous launch
    std::vector<uint8_t> partial;                                // an incomplete member from the previous piece
    std::vector<uint8_t> sample;                                 // the first text bytes (inflated on the host): format sniff + tile size
};
static void devgz_free(DevGz* d) {
    if (!d) return;
    if (d->gz_stream) { cudaStreamSynchronize(d->gz_stream); cudaStreamDestroy(d->gz_stream); }
    for (int i = 0; i < 2; i++) { cudaFreeHost(d->h_mem[i]); cudaFree(d->d_comp[i]); cudaFree(d->d_mem[i]); if (d->ev_free[i]) cudaEventDestroy(d->ev_free[i]); }
    if (d->ev_h2d) cudaEventDestroy(d->ev_h2d);
    cudaFree(d->d_err);
    delete d;
}
static int devgz_init(ntg_stream* s) {
    ntg_ctx* ctx = s->ctx;
    auto* d = new DevGz();
    s->devgz = d;
    d->comp_cap = DEVGZ_BATCH + (size_t(16) << 20);              // (a member is at most 64 KiB compressed, like its text; headers travel too)
    d->mem_cap = (uint32_t)(DEVGZ_BATCH >> 12);                  // a batch also closes after this many members (small members)
    for (int i = 0; i < 2; i++) {
        if (cudaMallocHost((void**)&d->h_mem[i], d->mem_cap * sizeof(gzdev::Member)) != cudaSuccess ||
            cudaMalloc((void**)&d->d_comp[i], d->comp_cap) != cudaSuccess || cudaMalloc((void**)&d->d_mem[i], d->mem_cap * sizeof(gzdev::Member)) != cudaSuccess) {
            cudaGetLastError();
            return ntg_set_error(ctx, NTG_ENOMEM, "device-inflate staging allocation failed");
        }
        NTG_CUDA(ctx, cudaEventCreateWithFlags(&d->ev_free[i], cudaEventDisableTiming));
    }
    NTG_CUDA(ctx, cudaEventCreateWithFlags(&d->ev_h2d, cudaEventDisableTiming));
    NTG_CUDA(ctx, cudaStreamCreateWithFlags(&d->gz_stream, cudaStreamNonBlocking));
    NTG_CUDA(ctx, cudaMalloc((void**)&d->d_err, sizeof(uint32_t)));
    NTG_CUDA(ctx, cudaMemsetAsync(d->d_err, 0, sizeof(uint32_t), ctx->copy_stream));
    return inflate_init(ctx);
}
// launch the staged members: member table to the device, inflate into the next segment, fused kernel over its whole tiles
static int devgz_flush(ntg_stream* s, bool final) {
    ntg_ctx* ctx = s->ctx; DevGz* d = s->devgz;
    if (s->failed || s->unsupported || !s->io_error.empty()) { d->comp_fill = 0; d->n_mem = 0; d->text_fill = 0; return NTG_OK; }
    if (!s->opened) {
        if (d->sample.empty()) return NTG_OK;                    // no text yet
        if (d->sample[0] != '>' && d->sample[0] != '@') { s->unsupported = true; s->unsupported_flags = fused::FLAG_FORMAT; return NTG_OK; }
        s->format = d->sample[0] == '>' ? NTG_FMT_FASTA : NTG_FMT_FASTQ;
        s->TB = pick_tile_bytes(d->sample.data(), d->sample.size() < 65536 ? d->sample.size() : 65536, s->format);
        NTG_TRY(s->feed.open(ctx, s->format, &s->cfg, s->TB, true, DEVGZ_BATCH + (size_t(2) << 20)));
        s->opened = true;
    }
    const uint64_t have = d->carry + d->text_fill;
    if (have == 0 && !final) return NTG_OK;
    const uint64_t ntiles = final ? (have + s->TB - 1) / s->TB : (have - 1) / s->TB;
    const size_t len = final ? (size_t)have : (size_t)(ntiles * (uint64_t)s->TB);
    if (!final && ntiles == 0) return NTG_OK;                    // (less than a tile so far: keep staging)
    const int cur = d->cur;
    const uint32_t n_mem = d->n_mem; const size_t carry = d->carry;
    NTG_CUDA(ctx, cudaEventRecord(d->ev_h2d, d->gz_stream));    // every compressed byte of this batch has been enqueued
    std::function<int(uint8_t*, const uint8_t*, size_t)> produce = [&](uint8_t* dst, const uint8_t* prev, size_t prev_len) -> int {
        if (carry) NTG_CUDA(ctx, cudaMemcpyAsync(dst, prev + prev_len, carry, cudaMemcpyDeviceToDevice, ctx->copy_stream));
        if (n_mem) {
            NTG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, d->ev_h2d, 0));
            NTG_CUDA(ctx, cudaMemcpyAsync(d->d_mem[cur], d->h_mem[cur], n_mem * sizeof(gzdev::Member), cudaMemcpyHostToDevice, ctx->copy_stream));
            NTG_TRY(inflate_enqueue(ctx, ctx->copy_stream, d->d_comp[cur], d->d_mem[cur], n_mem, dst, d->d_err));
        }
        NTG_CUDA(ctx, cudaEventRecord(d->ev_free[cur], ctx->copy_stream));
        return NTG_OK;
    };
    auto check = [&](uint64_t j, const LaunchCtl& c) { return stream_check(s, j, c); };
    const uint64_t n_total = s->submitted_bytes + have;
    s->in_submit = true;
    const int rc = s->feed.submit(nullptr, len, final, n_total, check, &produce);
    s->in_submit = false;
    NTG_TRY(rc);
    d->busy[cur] = true;
    s->submitted_bytes += len;
    d->carry = (size_t)(have - len);
    d->comp_fill = 0; d->n_mem = 0; d->text_fill = 0;
    d->cur ^= 1;
    // the other staging pair: its compressed bytes and table may be overwritten once the inflate that read them is done
    if (d->busy[d->cur]) { NTG_CUDA(ctx, cudaStreamWaitEvent(d->gz_stream, d->ev_free[d->cur], 0)); NTG_CUDA(ctx, cudaEventSynchronize(d->ev_free[d->cur])); d->busy[d->cur] = false; }
    return NTG_OK;
}
// a run of complete members [p, p + bytes) of the caller's buffer: table entries + one H2D copy
struct DevGzRun { const uint8_t* p = nullptr; size_t bytes = 0; };
static int devgz_send_run(ntg_stream* s, DevGzRun& run) {
    DevGz* d = s->devgz;
    if (!run.bytes) return NTG_OK;
    NTG_CUDA(s->ctx, cudaMemcpyAsync(d->d_comp[d->cur] + d->comp_fill, run.p, run.bytes, cudaMemcpyHostToDevice, d->gz_stream));
    d->comp_fill += run.bytes;
    run = DevGzRun{};
    return NTG_OK;
}
// one complete member (csize bytes at p): joins the current run (or starts one)
static int devgz_member(ntg_stream* s, const uint8_t* p, size_t csize, DevGzRun& run) {
    DevGz* d = s->devgz;
    size_t off, plen;
    if (!bgzf_payload(p, csize, &off, &plen)) { s->io_error = "truncated BGZF member"; return NTG_OK; }
    const uint32_t isize = bgzf_isize(p, csize);
    if (isize > 65536) { s->io_error = "BGZF member larger than 64 KiB"; return NTG_OK; }
    if (d->sample.size() < 65536 && isize) {                     // host inflate of the first members only: sniff + tile size
        const size_t old = d->sample.size();
        d->sample.resize(old + isize);
        if (!bgzf_inflate_member(p, csize, d->sample.data() + old, isize)) { s->io_error = "inflate: corrupt BGZF member"; return NTG_OK; }
    }
    if (d->comp_fill + run.bytes + csize + 8 > d->comp_cap || d->n_mem == d->mem_cap || d->carry + d->text_fill + isize > DEVGZ_BATCH + (size_t(1) << 20)) {
        NTG_TRY(devgz_send_run(s, run));
        NTG_TRY(devgz_flush(s, false));
    }
    if (run.bytes && run.p + run.bytes != p) NTG_TRY(devgz_send_run(s, run));
    if (!run.bytes) run.p = p;
    if (isize) {
        d->h_mem[d->cur][d->n_mem++] = gzdev::Member{d->comp_fill + run.bytes + off, d->carry + d->text_fill, (uint32_t)plen, isize};
        d->text_fill += isize; s->total_fed += isize;
    }
    run.bytes += csize;
    if (d->carry + d->text_fill >= DEVGZ_BATCH) { NTG_TRY(devgz_send_run(s, run)); NTG_TRY(devgz_flush(s, false)); }
    return NTG_OK;
}
static int devgz_feed(ntg_stream* s, const uint8_t* in, size_t n) {
    DevGz* d = s->devgz;
    size_t off = 0;
    DevGzRun run;
    while (off < n && s->io_error.empty()) {
        if (!d->partial.empty()) {
            // complete the member that the previous piece left unfinished
            const long ms = bgzf_member_size(d->partial.data(), d->partial.size());
            if (ms == 0) { s->io_error = "not a BGZF member where one was expected"; break; }
            const size_t need = ms < 0 ? d->partial.size() + 64 : (size_t)ms;
            if (d->partial.size() < need) {
                const size_t take = need - d->partial.size() < n - off ? need - d->partial.size() : n - off;
                d->partial.insert(d->partial.end(), in + off, in + off + take);
                off += take;
                continue;
            }
            if (ms < 0) continue;
            DevGzRun one;
            NTG_TRY(devgz_member(s, d->partial.data(), (size_t)ms, one));
            NTG_TRY(devgz_send_run(s, one));
            NTG_CUDA(s->ctx, cudaStreamSynchronize(d->gz_stream));      // (the copy reads `partial`)
            d->partial.clear();
            continue;
        }
        const long ms = bgzf_member_size(in + off, n - off);
        if (ms == 0) {
            if (in[off] == 0) { off = n; break; }                // zero padding behind the last member
            s->io_error = "not a BGZF member where one was expected"; break;
        }
        if (ms < 0 || (size_t)ms > n - off) { d->partial.assign(in + off, in + n); off = n; break; }
        NTG_TRY(devgz_member(s, in + off, (size_t)ms, run));
        off += (size_t)ms;
    }
    NTG_TRY(devgz_send_run(s, run));
    NTG_CUDA(s->ctx, cudaStreamSynchronize(d->gz_stream));             // the caller's buffer is free again
    return NTG_OK;
}
static int devgz_finish(ntg_stream* s) {
    ntg_ctx* ctx = s->ctx; DevGz* d = s->devgz;
    if (!d->partial.empty() && s->io_error.empty()) s->io_error = "gzip stream ends inside a member";
    if (!s->io_error.empty()) return NTG_OK;
    if (!s->opened && d->sample.size() < 2) {                    // the whole text is shorter than two bytes: host-side sniff rules below
        NTG_TRY(stream_need_staging(s));
        std::memcpy(s->pin[0], d->sample.data(), d->sample.size());
        s->fill = d->sample.size();
        devgz_free(d); s->devgz = nullptr;
        return NTG_OK;
    }
    NTG_TRY(devgz_flush(s, true));
    uint32_t e = 0;
    NTG_CUDA(ctx, cudaMemcpyAsync(&e, d->d_err, sizeof(e), cudaMemcpyDeviceToHost, ctx->copy_stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    if (e) s->io_error = "inflate: corrupt BGZF member (device), code " + std::to_string(e & 31u);
    return NTG_OK;
}

// One more piece of a gzip stream.  threads > 1 and a BGZF first member: block-parallel inflate; else sequential.
// threads == 0 and BGZF: the members are inflated on the device (NTG_GZ_DEVICE).
static int stream_feed_gz(ntg_stream* s, const uint8_t* in, size_t n, int threads) {
    if (!s->gz) s->gz = new GzState();
    GzState* g = s->gz;
    if (!n || !s->io_error.empty()) return NTG_OK;
    if (!g->decided) {
        // (a first piece too short to show the extra field is taken as plain gzip)
        const bool is_bgzf = bgzf_member_size(in, n) > 0;
        g->bgzf = threads > 1 && is_bgzf;
        g->decided = true;
        if (threads == 0 && is_bgzf && s->total_fed == 0 && !s->opened) NTG_TRY(devgz_init(s));
    }
    if (s->devgz) return devgz_feed(s, in, n);
    return g->bgzf ? gz_feed_bgzf(s, in, n, threads) : gz_feed_sequential(s, in, n);
}
static int gz_finish(ntg_stream* s) {
    GzState* g = s->gz;
    if (!g || !s->io_error.empty()) return NTG_OK;
    if (g->bgzf ? !g->carry.empty() : (g->z_init && !g->between_members && !g->ended)) s->io_error = "gzip stream ends inside a member";
    return NTG_OK;
}
