// scan.cuh — device-wide exclusive prefix sum of byte flags (reduce / scan partials / downsweep).
// Used by the materialising paths (record tables, compaction); the fused tallies kernel has its
// own single-pass decoupled look-back (fused.cuh).
#pragma once
#include "common.cuh"

namespace scan_detail {
constexpr int BLOCK = 256;
constexpr int ITEMS = 16;                 // bytes per thread (one uint4 load)
constexpr int TILE = BLOCK * ITEMS;       // 4096 flags per block

__device__ __forceinline__ uint32_t sum16(const uint8_t* p, size_t base, size_t n) {
    uint32_t s = 0;
    if (base + ITEMS <= n && ((reinterpret_cast<uintptr_t>(p + base) & 15) == 0)) {
        uint4 v = *reinterpret_cast<const uint4*>(p + base);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) s += (w[i] & 0xFF) + ((w[i] >> 8) & 0xFF) + ((w[i] >> 16) & 0xFF) + (w[i] >> 24);
    } else {
        for (int i = 0; i < ITEMS; i++) if (base + i < n) s += p[base + i];
    }
    return s;
}
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* smem /*BLOCK/32+1*/) {
    uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t x = lane < BLOCK / 32 ? smem[lane] : 0, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= d) xi += t; }
        if (lane < BLOCK / 32) smem[lane] = xi - x;
        if (lane == BLOCK / 32 - 1) smem[BLOCK / 32] = xi;
    }
    __syncthreads();
    uint32_t r = inc - v + smem[w];
    if (total) *total = smem[BLOCK / 32];
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(BLOCK) k_reduce(const uint8_t* flags, size_t n, uint32_t* partial) {
    __shared__ uint32_t sm[BLOCK / 32 + 1];
    size_t base = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * ITEMS;
    uint32_t s = base < n ? sum16(flags, base, n) : 0, tot;
    block_exclusive_scan(s, &tot, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(BLOCK) k_scan_partials(uint32_t* partial, size_t nb) {   // single block
    __shared__ uint32_t sm[BLOCK / 32 + 1];
    uint32_t carry = 0;
    for (size_t base = 0; base < nb; base += BLOCK) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < nb ? partial[i] : 0, tot;
        uint32_t e = block_exclusive_scan(v, &tot, sm);
        if (i < nb) partial[i] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0) partial[nb] = carry;
}
__global__ void __launch_bounds__(BLOCK) k_downsweep(const uint8_t* flags, size_t n, const uint32_t* partial, uint32_t* out, size_t nb) {
    __shared__ uint32_t sm[BLOCK / 32 + 1];
    size_t base = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * ITEMS;
    uint32_t s = base < n ? sum16(flags, base, n) : 0;
    uint32_t e = block_exclusive_scan(s, nullptr, sm) + partial[blockIdx.x];
    if (base < n) {
        for (int i = 0; i < ITEMS && base + i < n; i++) { out[base + i] = e; e += flags[base + i]; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = partial[nb];
}
}  // namespace scan_detail

static size_t scan_tmp_count(size_t n) { return (n + scan_detail::TILE - 1) / scan_detail::TILE + 2; }

static int exclusive_scan_u8(ntg_ctx* ctx, const uint8_t* flags, uint32_t* out, size_t n, uint32_t* tmp) {
    using namespace scan_detail;
    size_t nb = (n + TILE - 1) / TILE;
    if (nb == 0) { NTG_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(uint32_t), ctx->stream)); return NTG_OK; }
    k_reduce<<<(unsigned)nb, BLOCK, 0, ctx->stream>>>(flags, n, tmp);
    k_scan_partials<<<1, BLOCK, 0, ctx->stream>>>(tmp, nb);
    k_downsweep<<<(unsigned)nb, BLOCK, 0, ctx->stream>>>(flags, n, tmp, out, nb);
    ctx->launches += 3;
    NTG_CUDA(ctx, cudaGetLastError());
    return NTG_OK;
}
