// common.cuh — shared host/device helpers for libntgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ntgpu.h"

// ----------------------------------------------------------------------------- context
// Grow-only device scratch of the materialising paths (record scanner): a context keeps its buffers between calls instead
// of paying cudaMalloc / cudaFree (which synchronise the device) for ~6 buffers per window.  Calls on one context are serial.
struct ScratchPool {
    enum Slot { BYTES, COUNTS, NLPOS, STPOS, STORD, STCR, RECS, ERRKIND, MISC,                                  // record scanner
                SQ_SEQS, SQ_OFFS, SQ_RC, SQ_FLAGS, SQ_IDX, SQ_TMP, SQ_OUT, SQ_CHG, SQ_LO, SQ_HI, SQ_OOFFS, SLOTS };  // Sequence batch calls
    void* dev[SLOTS] = {};
    size_t cap[SLOTS] = {};
    void* get(int slot, size_t bytes) {                  // nullptr: out of device memory
        if (dev[slot] && bytes <= cap[slot]) return dev[slot];
        if (dev[slot]) { cudaFree(dev[slot]); dev[slot] = nullptr; cap[slot] = 0; }
        size_t want = bytes + bytes / 8 + 256;            // headroom: windows of a stream differ a little in their counts
        if (cudaMalloc(&dev[slot], want) != cudaSuccess) {
            cudaGetLastError();
            want = bytes ? bytes : 1;
            if (cudaMalloc(&dev[slot], want) != cudaSuccess) { cudaGetLastError(); dev[slot] = nullptr; return nullptr; }
        }
        cap[slot] = want;
        return dev[slot];
    }
    void release() { for (int i = 0; i < SLOTS; i++) { if (dev[i]) cudaFree(dev[i]); dev[i] = nullptr; cap[i] = 0; } }
};
// Pinned host buffers of record tables: a freed ntg_records hands its buffer back (pinning costs ~0.1 ms per MiB).  Shared by the
// context and the tables it returned, so a table may outlive its context.
struct PinPool {
    struct Ent { void* p; size_t cap; };
    static constexpr size_t MAX_SPARE = 8;
    std::mutex mu;
    bool closed = false;
    std::vector<Ent> spare;
    void* take(size_t bytes, size_t* cap_out) {
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = spare.size();
            for (size_t i = 0; i < spare.size(); i++)
                if (spare[i].cap >= bytes && (best == spare.size() || spare[i].cap < spare[best].cap)) best = i;
            if (best < spare.size()) { Ent e = spare[best]; spare.erase(spare.begin() + best); *cap_out = e.cap; return e.p; }
        }
        void* p = nullptr;
        const size_t want = bytes + bytes / 8 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        *cap_out = want;
        return p;
    }
    void give(void* p, size_t cap) {
        std::lock_guard<std::mutex> g(mu);
        if (closed) { cudaFreeHost(p); return; }
        spare.push_back(Ent{p, cap});
        if (spare.size() > MAX_SPARE) {                   // drop the smallest one
            size_t w = 0;
            for (size_t i = 1; i < spare.size(); i++) if (spare[i].cap < spare[w].cap) w = i;
            cudaFreeHost(spare[w].p);
            spare.erase(spare.begin() + w);
        }
    }
    void trim() {                                          // give the spare buffers back, keep serving
        std::lock_guard<std::mutex> g(mu);
        for (auto& e : spare) cudaFreeHost(e.p);
        spare.clear();
    }
    void close() {
        std::lock_guard<std::mutex> g(mu);
        closed = true;
        for (auto& e : spare) cudaFreeHost(e.p);
        spare.clear();
    }
};

struct ntg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;        // compute stream: every kernel is launched here
    cudaStream_t copy_stream = nullptr;   // H2D feed for the end-to-end path
    cudaEvent_t events[64] = {};
    std::string last_error;
    uint64_t launches = 0;
    // fused path state (fused.cu)
    struct FusedState* fused = nullptr;
    // NCCL (nccl_dyn.cpp)
    void* nccl_comm = nullptr;
    int nccl_ranks = 1, nccl_rank = 0;
    void* nccl_buf = nullptr;             // device staging for the tallies all-reduce
    ScratchPool scratch;
    std::shared_ptr<PinPool> pinpool = std::make_shared<PinPool>();
};

int ntg_set_error(ntg_ctx* ctx, int status, const char* fmt, ...);

#define NTG_CUDA(ctx, expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            return ntg_set_error((ctx), NTG_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,  \
                                 cudaGetErrorString(_e));                                        \
    } while (0)

#define NTG_TRY(expr)                  \
    do {                               \
        int _s = (expr);               \
        if (_s != NTG_OK) return _s;   \
    } while (0)

// RAII device buffer (freed on scope exit; allocation failures surface as NTG_ENOMEM)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    bool owned = true;
    ~DevBuf() { if (p && owned) cudaFree(p); }
    cudaError_t alloc(size_t count) {
        n = count;
        return cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
    }
    // a slot of the context's grow-only scratch instead of an allocation of its own (nothing is freed here)
    cudaError_t alloc_pooled(ScratchPool& pool, int slot, size_t count) {
        n = count; owned = false;
        p = static_cast<T*>(pool.get(slot, (count ? count : 1) * sizeof(T)));
        return p ? cudaSuccess : cudaErrorMemoryAllocation;
    }
};
// RAII pinned host buffer
template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t n = 0;
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    std::shared_ptr<PinPool> pool;        // set: the buffer came from (and goes back to) the context's pinned pool
    size_t cap = 0;
    ~PinBuf() { if (p) { if (pool) pool->give(p, cap); else cudaFreeHost(p); } }
    cudaError_t alloc(size_t count) {
        n = count;
        return cudaMallocHost((void**)&p, (count ? count : 1) * sizeof(T));
    }
    cudaError_t alloc_pooled(const std::shared_ptr<PinPool>& from, size_t count) {
        n = count; pool = from;
        p = static_cast<T*>(pool->take((count ? count : 1) * sizeof(T), &cap));
        return p ? cudaSuccess : cudaErrorMemoryAllocation;
    }
    T* release() { T* q = p; p = nullptr; return q; }
};

// ----------------------------------------------------------------------------- byte classes
// Built once on the host (luts.cuh) and copied to device memory at ntg_create.
//  c_norm[iupac][b] : output byte of sequence::normalize (src/sequence.rs:19-62); 0 = deleted
//  c_comp[b]        : sequence::complement (src/sequence.rs:67-105)
//  c_code[b]        : 0..3 = A/C/G/T (case-insensitive, src/bitkmer.rs:5-18 == kmer.rs:6-8 good set), 4 = not a good base
//  c_ncls[b]        : class of b *after* normalize: 0..3 = ACGT, 4 = kept but not ACGT (N, -, IUPAC), 5 = deleted
// (global memory + __ldg: per-lane divergent indices would serialise in __constant__ memory)
// libntgpu is a unity build (ntgpu.cu includes every *.cuh once), hence `static`.
static __device__ uint8_t c_norm[2][256];
static __device__ uint8_t c_comp[256];
static __device__ uint8_t c_code[256];
static __device__ uint8_t c_ncls[256];
static int ntg_upload_luts(ntg_ctx* ctx);

// ----------------------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

// 2 * k-bit mask with k <= 32 (k == 32 -> all ones), mirrors Rust's wrapping pow in bitkmer.rs:31
__host__ __device__ __forceinline__ uint64_t mask2k(uint32_t k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }

// bitkmer::reverse_complement (src/bitkmer.rs:112-132) via brev + adjacent-bit swap
__device__ __forceinline__ uint64_t bit_rc(uint64_t x, uint32_t k) {
    uint64_t r = __brevll(x);                                               // reverses single bits
    r = ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);  // restore order inside each 2-bit group
    r = ~r;
    uint32_t sh = 2 * (32 - k);
    return sh >= 64 ? 0ull : (r >> sh);
}
// bitkmer::minimizer (src/bitkmer.rs:146-162), RC at width k
__device__ __forceinline__ uint64_t bit_minimizer_slow(uint64_t kmer, uint32_t k, uint32_t m) {
    uint64_t lowest = ~0ull, bm = mask2k(m);
    for (uint32_t i = 0; i + m <= k; i++) {
        uint64_t cur = kmer & bm;
        if (cur < lowest) lowest = cur;
        uint64_t r = bit_rc(cur, k);
        if (r < lowest) lowest = r;
        kmer >>= 2;
    }
    return lowest;
}

// ----------------------------------------------------------------------------- device-wide scan
// exclusive prefix sum of n uint8 flags into uint32 (out has n+1 entries; out[n] = total).
// tmp must hold scan_tmp_count(n) uint32.  Three kernels (reduce / scan partials / downsweep).
static size_t scan_tmp_count(size_t n);
static int exclusive_scan_u8(ntg_ctx* ctx, const uint8_t* flags, uint32_t* out, size_t n, uint32_t* tmp);
