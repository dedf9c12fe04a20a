// inflate.cuh — DEFLATE (RFC 1951) on the device for block-compressed gzip (BGZF, SAM spec 4.1).
//
// The reference decompresses on the host in front of the parser (flate2::MultiGzDecoder, src/parser/mod.rs:96-108); for inputs
// whose members are independent (BGZF: at most 64 KiB each, sizes in the member headers) the members can be inflated in
// parallel where the text is consumed: the compressed bytes cross PCIe (2-4x fewer than the text) and k_inflate writes the
// text straight into the device segment the fused kernel reads.  The host only walks the member headers.
//
// One THREAD per member: a member is a serial bit stream, 64 KiB of output at most; a 2 GiB batch holds ~33 000 of them.
// Canonical Huffman decoding by code length (count-per-length in registers, symbols-by-code-order in shared memory,
// interleaved across the CTA's threads so that neighbouring lanes use neighbouring banks); length / distance base values are
// arithmetic, not tables.  Every malformed stream (bad block type, over-subscribed code, distance before the start, output
// overrun, input overrun, length mismatch with ISIZE) sets the member's error word; CRC-32 is not recomputed on the device.
// Part of the unity build (ntgpu.cu).
#pragma once
#include "common.cuh"

namespace gzdev {
constexpr int THREADS = 128;
constexpr int MAXL = 288, MAXD = 30;

struct Member {
    uint64_t in_off;       // offset of the raw DEFLATE payload inside the compressed buffer
    uint64_t out_off;      // offset of the member's text inside the output buffer
    uint32_t in_len;       // payload bytes (without the gzip header and the CRC32 / ISIZE trailer)
    uint32_t out_len;      // ISIZE
};

struct Bits {
    const uint8_t* p; const uint8_t* end;
    uint64_t buf; uint32_t cnt; uint32_t err;
    // bytes up to the next 4-byte boundary, then one aligned 32-bit load per refill (the tail of the payload by bytes again)
    __device__ __forceinline__ void refill() {
        while (cnt <= 56 && p < end && ((uintptr_t)p & 3u)) { buf |= (uint64_t)(*p++) << cnt; cnt += 8; }
        if (cnt <= 32 && p + 4 <= end) { buf |= (uint64_t)(*reinterpret_cast<const uint32_t*>(p)) << cnt; p += 4; cnt += 32; }
        else if (p + 4 > end) while (cnt <= 56 && p < end) { buf |= (uint64_t)(*p++) << cnt; cnt += 8; }
    }
    __device__ __forceinline__ uint32_t take(uint32_t n) {          // n <= 16
        if (cnt < n) { refill(); if (cnt < n) { err = 1; cnt = 0; buf = 0; return 0; } }
        const uint32_t v = (uint32_t)buf & ((1u << n) - 1u);
        buf >>= n; cnt -= n;
        return v;
    }
};

// counts per code length 1..15, two per 32-bit word (registers after unrolling); index 0 unused
struct Counts { uint32_t w[8]; };
__device__ __forceinline__ uint32_t count_of(const Counts& c, int len) { return (c.w[len >> 1] >> (16 * (len & 1))) & 0xFFFFu; }

// symbols ordered by code (canonical order) from the code lengths; returns false for an over-subscribed set
__device__ __forceinline__ bool build(const uint8_t* lengths, int n, Counts& c, uint16_t* sym, int stride) {
    uint16_t cnt[16], offs[16];
#pragma unroll
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int s = 0; s < n; s++) cnt[lengths[s]]++;
    int left = 1;
#pragma unroll
    for (int len = 1; len <= 15; len++) { left <<= 1; left -= cnt[len]; if (left < 0) return false; }
    offs[1] = 0;
#pragma unroll
    for (int len = 1; len < 15; len++) offs[len + 1] = offs[len] + cnt[len];
    for (int s = 0; s < n; s++) if (lengths[s]) sym[(offs[lengths[s]]++) * stride] = (uint16_t)s;
#pragma unroll
    for (int i = 0; i < 8; i++) c.w[i] = (uint32_t)cnt[2 * i] | ((uint32_t)cnt[2 * i + 1] << 16);
    c.w[0] &= 0xFFFF0000u;                                           // length 0 = unused symbols
    return true;
}
// one symbol: walk the code lengths (RFC 1951 3.2.2: codes of one length are consecutive, shorter codes come first)
__device__ __forceinline__ int decode(Bits& b, const Counts& c, const uint16_t* sym, int stride) {
    if (b.cnt < 15) b.refill();
    uint32_t bits = (uint32_t)b.buf;
    int code = 0, first = 0, index = 0;
#pragma unroll
    for (int len = 1; len <= 15; len++) {
        code |= (int)(bits & 1u); bits >>= 1;
        const int count = (int)count_of(c, len);
        if (code - count < first) {
            if (b.cnt < (uint32_t)len) { b.err = 1; return -1; }
            b.buf >>= len; b.cnt -= len;
            return sym[(index + (code - first)) * stride];
        }
        index += count; first += count; first <<= 1; code <<= 1;
    }
    b.err = 1;
    return -1;
}

__global__ void __launch_bounds__(THREADS) k_inflate(const uint8_t* __restrict__ comp, const Member* __restrict__ members, uint32_t n_members,
                                                    uint8_t* __restrict__ out, uint32_t* __restrict__ errors) {
    extern __shared__ uint16_t s_tables[];                            // [MAXL + MAXD][THREADS], interleaved by thread
    uint16_t* s_lsym = s_tables;
    uint16_t* s_dsym = s_tables + MAXL * THREADS;
    const uint32_t g = blockIdx.x * THREADS + threadIdx.x;
    if (g >= n_members) return;
    const Member m = members[g];
    uint16_t* lsym = s_lsym + threadIdx.x;
    uint16_t* dsym = s_dsym + threadIdx.x;
    Bits b{comp + m.in_off, comp + m.in_off + m.in_len, 0, 0, 0};
    uint8_t* o = out + m.out_off;
    uint32_t pos = 0;
    uint32_t bad = 0;
    uint8_t lengths[MAXL + MAXD + 2];
    Counts lc, dc;
    // Control flow note: no `break` out of a divergent branch.  A branch that can leave the loop has its reconvergence point
    // behind the loop, so lanes that once took different sides (literal / match) would never rejoin and the warp would run
    // its 32 members one after the other (measured: 73 MB/s).  Every exit goes through a flag tested at the loop head.
    bool more = true;
    while (more && !bad) {
        const uint32_t last = b.take(1), type = b.take(2);
        if (b.err) bad = 1;
        else if (type == 0) {
            // stored block: to the byte boundary, LEN, ~LEN, LEN raw bytes
            const uint32_t drop = b.cnt & 7u;
            b.buf >>= drop; b.cnt -= drop;
            const uint32_t len = b.take(16), nlen = b.take(16);
            if (b.err || (len ^ 0xFFFFu) != nlen || pos + len > m.out_len) bad = 2;
            for (uint32_t i = 0; i < len && !bad; i++) {
                uint32_t v = 0;
                if (b.cnt >= 8) { v = (uint32_t)b.buf & 0xFFu; b.buf >>= 8; b.cnt -= 8; }
                else if (b.p < b.end) v = *b.p++;
                else bad = 3;
                if (!bad) o[pos++] = (uint8_t)v;
            }
        } else if (type == 3) bad = 17;
        else {
            int nlen = 288, ndist = 30;
            if (type == 1) {
                // fixed code (RFC 1951 3.2.6)
                for (int s = 0; s < 144; s++) lengths[s] = 8;
                for (int s = 144; s < 256; s++) lengths[s] = 9;
                for (int s = 256; s < 280; s++) lengths[s] = 7;
                for (int s = 280; s < 288; s++) lengths[s] = 8;
                for (int s = 0; s < 30; s++) lengths[288 + s] = 5;
            } else {
                nlen = (int)b.take(5) + 257; ndist = (int)b.take(5) + 1;
                const int ncode = (int)b.take(4) + 4;
                if (b.err || nlen > 286 || ndist > 30) { bad = 4; nlen = 257; ndist = 1; }
                // code-length code: 19 symbols in the order of 3.2.7; its table uses the distance slot until the real one is built
                for (int i = 0; i < 19; i++) lengths[i] = 0;
                for (int i = 0; i < ncode && !bad; i++) {
                    // 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
                    const int sym_i = i < 3 ? 16 + i : (i == 3 ? 0 : ((i & 1) ? 7 - ((i - 5) >> 1) : 8 + ((i - 4) >> 1)));
                    lengths[sym_i] = (uint8_t)b.take(3);
                }
                Counts cc;
                if (!bad && (b.err || !build(lengths, 19, cc, dsym, THREADS))) bad = 5;
                int idx = 0;
                while (idx < nlen + ndist && !bad) {
                    const int s = decode(b, cc, dsym, THREADS);
                    if (s < 0) bad = 6;
                    else if (s < 16) lengths[idx++] = (uint8_t)s;
                    else {
                        int rep, val = 0;
                        if (s == 16) { if (idx == 0) bad = 7; else val = lengths[idx - 1]; rep = 3 + (int)b.take(2); }
                        else if (s == 17) rep = 3 + (int)b.take(3);
                        else rep = 11 + (int)b.take(7);
                        if (idx + rep > nlen + ndist) bad = 8;
                        for (; rep > 0 && !bad; rep--) lengths[idx++] = (uint8_t)val;
                    }
                }
                if (!bad && b.err) bad = 9;
                if (!bad && lengths[256] == 0) bad = 10;                   // no end-of-block code
            }
            if (!bad) {
                // move the distance lengths out of the way of build() (it reads lengths[0..n))
                uint8_t dl[MAXD];
                for (int s = 0; s < ndist; s++) dl[s] = lengths[nlen + s];
                if (!build(lengths, nlen, lc, lsym, THREADS) || !build(dl, ndist, dc, dsym, THREADS)) bad = 11;
            }
            bool in_block = !bad;
            while (in_block) {
                int s = decode(b, lc, lsym, THREADS);
                if (s < 0) { bad = 12; in_block = false; }
                else if (s < 256) {
                    if (pos >= m.out_len) { bad = 13; in_block = false; }
                    else o[pos++] = (uint8_t)s;
                } else if (s == 256) in_block = false;
                else {
                    s -= 257;
                    uint32_t len = 0, dist = 0;
                    if (s >= 29) bad = 14;
                    else {
                        if (s < 8) len = 3 + s;
                        else if (s == 28) len = 258;
                        else { const int e = (s - 4) >> 2; len = 3 + ((4 + (s & 3)) << e) + b.take(e); }
                        const int ds = decode(b, dc, dsym, THREADS);
                        if (ds < 0 || ds >= 30) bad = 15;
                        else {
                            if (ds < 4) dist = 1 + ds;
                            else { const int e = (ds - 2) >> 1; dist = 1 + ((2 + (ds & 1)) << e) + b.take(e); }
                            if (b.err || dist > pos || pos + len > m.out_len) bad = 16;
                        }
                    }
                    if (bad) in_block = false;
                    else {
                        // The source bytes were written by this thread a moment ago and live in L2 (stores do not allocate in
                        // L1): a byte-by-byte copy pays one L2 round trip per byte.  Chunks of up to 8 bytes, never longer than
                        // `dist` (so a chunk's sources are older than its destinations): all loads of a chunk are in flight together.
                        const uint8_t* src = o + pos - dist;
                        for (uint32_t i = 0; i < len;) {
                            uint32_t c = len - i; c = c < 8u ? c : 8u; c = c < dist ? c : dist;
                            uint8_t t[8];
#pragma unroll
                            for (uint32_t j = 0; j < 8; j++) if (j < c) t[j] = src[i + j];
#pragma unroll
                            for (uint32_t j = 0; j < 8; j++) if (j < c) o[pos + i + j] = t[j];
                            i += c;
                        }
                        pos += len;
                    }
                }
            }
        }
        if (last) more = false;
    }
    if (!bad && pos != m.out_len) bad = 18;
    if (bad) atomicMax(errors, (g << 5) | bad | 0x80000000u);
}
constexpr size_t SMEM = (size_t)(MAXL + MAXD) * THREADS * sizeof(uint16_t);
}  // namespace gzdev

// ---- host side: BGZF member walk + one inflate launch -------------------------------------------------------------------
// Header of the BGZF member at p (n bytes available): 0 = not BGZF, -1 = need more bytes, else the member's compressed size.
static long bgzf_member_size(const uint8_t* p, size_t n) {
    if (n < 12) return -1;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
    const size_t xlen = p[10] | (p[11] << 8);
    if (12 + xlen > n) return -1;
    for (size_t o = 12; o + 4 <= 12 + xlen;) {
        const size_t slen = p[o + 2] | (p[o + 3] << 8);
        if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) {
            const size_t bs = (size_t)(p[o + 4] | (p[o + 5] << 8)) + 1;
            return bs < 12 + xlen + 8 ? 0 : (long)bs;
        }
        o += 4 + slen;
    }
    return 0;
}
static uint32_t bgzf_isize(const uint8_t* p, size_t csize) {
    return (uint32_t)p[csize - 4] | ((uint32_t)p[csize - 3] << 8) | ((uint32_t)p[csize - 2] << 16) | ((uint32_t)p[csize - 1] << 24);
}
// payload (raw DEFLATE) of a complete member: offset of its first byte and its length.  Optional gzip header fields
// (FNAME, FCOMMENT, FHCRC) are skipped like zlib does.
static bool bgzf_payload(const uint8_t* p, size_t csize, size_t* off, size_t* len) {
    size_t o = 12 + (p[10] | (p[11] << 8));
    if (p[3] & 8) { while (o < csize && p[o]) o++; o++; }
    if (p[3] & 16) { while (o < csize && p[o]) o++; o++; }
    if (p[3] & 2) o += 2;
    if (o + 8 > csize) return false;
    *off = o; *len = csize - 8 - o;
    return true;
}
static int inflate_init(ntg_ctx* ctx) {
    NTG_CUDA(ctx, cudaFuncSetAttribute(gzdev::k_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gzdev::SMEM));
    return NTG_OK;
}
// enqueue: members[0..n) of the compressed device buffer -> text at d_out (offsets inside the table)
static int inflate_enqueue(ntg_ctx* ctx, cudaStream_t stream, const uint8_t* d_comp, const gzdev::Member* d_members, uint32_t n, uint8_t* d_out,
                           uint32_t* d_err) {
    if (!n) return NTG_OK;
    gzdev::k_inflate<<<(n + gzdev::THREADS - 1) / gzdev::THREADS, gzdev::THREADS, gzdev::SMEM, stream>>>(d_comp, d_members, n, d_out, d_err);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    return NTG_OK;
}
