// synth.cuh — device generator of the deterministic synthetic FASTQ / FASTA inputs
// (spec: DESIGN.md "Synthetic inputs"; bytes identical to oracle/synth.hpp).  Unity build.
#pragma once
#include "common.cuh"

namespace synth {
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t rec, uint64_t w) {
    return mix64(seed + 0x9E3779B97F4A7C15ull * ((rec << 26) | (stream << 24) | w));
}
// one thread per (record, 32-byte-ish unit): units 0.. cover header / bases / separator / quals
// Simple and coalesced enough for an untimed generator: thread per output byte group of 8.
__global__ void __launch_bounds__(256) k_gen(uint8_t* __restrict__ out, uint64_t seed, uint64_t rec0, uint64_t nrec,
                                            uint32_t L, uint32_t n_thresh, int fastq) {
    const uint64_t rec_bytes = fastq ? 2ull * L + 16 : (uint64_t)L + 12;
    const uint64_t total = nrec * rec_bytes;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = g / rec_bytes, o = g - r * rec_bytes, rec = rec0 + r;
        const uint32_t hdr = fastq ? 12 : 11, nd = fastq ? 9 : 8;
        uint8_t b;
        if (o < hdr) {
            if (o == 0) b = fastq ? '@' : '>';
            else if (o == 1) b = 'r';
            else if (o == hdr - 1) b = '\n';
            else { uint64_t v = rec; for (uint32_t i = 0; i < nd - 1 - (o - 2); i++) v /= 10; b = (uint8_t)('0' + v % 10); }
        } else if (o < hdr + L) {
            const uint64_t j = o - hdr;
            b = (uint8_t)"ACGT"[(rnd(seed, 0, rec, j >> 5) >> (2 * (j & 31))) & 3];
            if (n_thresh && ((rnd(seed, 1, rec, j >> 2) >> (16 * (j & 3))) & 0xFFFF) < n_thresh) b = 'N';
        } else if (o == hdr + L) b = '\n';
        else if (o == hdr + L + 1) b = '+';
        else if (o == hdr + L + 2) b = '\n';
        else if (o < hdr + 2ull * L + 3) {
            const uint64_t j = o - (hdr + L + 3);
            const uint32_t by = (uint32_t)((rnd(seed, 2, rec, j >> 3) >> (8 * (j & 7))) & 0xFF);
            b = (uint8_t)('!' + ((by * 42u) >> 8));
        } else b = '\n';
        out[g] = b;
    }
}
}  // namespace synth

static int run_synth(ntg_ctx* ctx, uint64_t dptr, uint64_t seed, uint64_t rec0, uint64_t nrec, uint32_t L, uint32_t n_thresh, int fastq) {
    if (!dptr) return ntg_set_error(ctx, NTG_EINVAL, "null device pointer");
    if (L == 0 || L > (1u << 24) * 32u - 1) return ntg_set_error(ctx, NTG_EINVAL, "read_len out of range");
    if (rec0 + nrec > (fastq ? 1000000000ull : 100000000ull)) return ntg_set_error(ctx, NTG_EINVAL, "record index exceeds the header digits");
    if (nrec == 0) return NTG_OK;
    unsigned grid = (unsigned)ctx->sm_count * 32;
    synth::k_gen<<<grid, 256, 0, ctx->stream>>>((uint8_t*)dptr, seed, rec0, nrec, L, n_thresh, fastq);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    return NTG_OK;
}
