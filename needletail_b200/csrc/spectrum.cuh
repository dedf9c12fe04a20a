// spectrum.cuh — k-mer spectrum: how often every distinct canonical k-mer occurs (unity build).
//
// The consumer every real user of `canonical_kmers` ends in (the README loop counts one k-mer: src/lib.rs:31-35).
// The reference has no counter, so the definition is the multiset of items of
//     for rec in reader { rec.normalize(false).canonical_kmers(k, &rc) }            (bit form: bit_kmers(k, true), k <= 32)
// keyed by the 2-bit pack of the canonical k-mer, and the oracle's multiset is what the tests compare with.
//   k <= 14 : dense histogram, 4^k u32 counters (1 GiB at k = 14); multi-GPU: one ncclAllReduce over the whole histogram
//   k <= 32 : open-addressing hash table (u64 keys, u32 counts, linear probing, MurmurHash3 finaliser); multi-GPU: entries are
//             exchanged by owner rank (hash of the key) with grouped ncclSend / ncclRecv, so every rank ends with the job-wide
//             counts of the keys it owns — a reduce-scatter by k-mer hash over NVLink
// Counting runs inside the fused pass (the generic walker counts what it tallies).  A validating tally pass goes first: an
// input with a parse error contributes the records in front of the failing one, like the iterator.
#pragma once
#include "fused_host.cuh"

struct ntg_spectrum {
    ntg_ctx* ctx = nullptr;
    uint32_t k = 0;
    bool dense = false;
    uint64_t capacity = 0;                  // slots (hash) / bins (dense)
    uint32_t* d_dense = nullptr;
    unsigned long long* d_keys = nullptr;
    uint32_t* d_counts = nullptr;
    uint32_t* d_overflow = nullptr;
    uint64_t n_added = 0;                   // k-mers counted so far
};

namespace spectrum {
constexpr int BLOCK = 256;
__global__ void __launch_bounds__(BLOCK) k_lookup(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t mask,
                                                  uint64_t key, uint32_t* out) {
    uint64_t h = fused::fmix64(key) & mask;
    for (uint32_t probe = 0; probe < fused::SPECTRUM_MAX_PROBES; probe++) {
        const unsigned long long cur = keys[h];
        if (cur == key) { *out = counts[h]; return; }
        if (cur == ~0ull) break;
        h = (h + 1) & mask;
    }
    *out = 0;
}
// compact the occupied entries: out_keys / out_counts in arbitrary order, *cursor = number of distinct k-mers
__global__ void __launch_bounds__(BLOCK) k_export(const uint32_t* __restrict__ dense, const unsigned long long* __restrict__ keys,
                                                  const uint32_t* __restrict__ counts, uint64_t n_slots, unsigned long long* out_keys,
                                                  uint32_t* out_counts, uint64_t cap, unsigned long long* cursor) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * BLOCK) {
        const uint32_t c = dense ? dense[i] : counts[i];
        const bool occ = dense ? c != 0 : keys[i] != ~0ull;
        if (!occ) continue;
        const unsigned long long o = atomicAdd(cursor, 1ull);
        if (o < cap) { out_keys[o] = dense ? i : keys[i]; out_counts[o] = c; }
    }
}
// count-of-counts: hist[min(c, n_bins - 1)]++ over the occupied entries
__global__ void __launch_bounds__(BLOCK) k_count_hist(const uint32_t* __restrict__ dense, const unsigned long long* __restrict__ keys,
                                                      const uint32_t* __restrict__ counts, uint64_t n_slots, unsigned long long* hist, uint32_t n_bins) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * BLOCK) {
        const uint32_t c = dense ? dense[i] : counts[i];
        const bool occ = dense ? c != 0 : keys[i] != ~0ull;
        if (occ) atomicAdd(&hist[c < n_bins - 1 ? c : n_bins - 1], 1ull);
    }
}
// owner rank of a key and the per-owner partition of the occupied entries (hash tables, multi-GPU)
__device__ __forceinline__ uint32_t owner_of(uint64_t key, uint32_t world) { return (uint32_t)((fused::fmix64(key ^ 0x9E3779B97F4A7C15ull) >> 32) % world); }
__global__ void __launch_bounds__(BLOCK) k_owner_counts(const unsigned long long* __restrict__ keys, uint64_t n_slots, uint32_t world, unsigned long long* per_owner) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * BLOCK)
        if (keys[i] != ~0ull) atomicAdd(&per_owner[owner_of(keys[i], world)], 1ull);
}
__global__ void __launch_bounds__(BLOCK) k_owner_scatter(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n_slots,
                                                         uint32_t world, unsigned long long* cursors /* start offsets, advanced */, unsigned long long* out_keys,
                                                         uint32_t* out_counts) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n_slots; i += (uint64_t)gridDim.x * BLOCK)
        if (keys[i] != ~0ull) {
            const unsigned long long o = atomicAdd(&cursors[owner_of(keys[i], world)], 1ull);
            out_keys[o] = keys[i]; out_counts[o] = counts[i];
        }
}
__global__ void __launch_bounds__(BLOCK) k_insert(const unsigned long long* __restrict__ in_keys, const uint32_t* __restrict__ in_counts, uint64_t n,
                                                  unsigned long long* keys, uint32_t* counts, uint64_t mask, uint32_t* overflow) {
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (uint64_t)gridDim.x * BLOCK)
        fused::spectrum_add(keys, counts, mask, overflow, in_keys[i], in_counts[i]);
}
static inline unsigned grid_for(uint64_t n, int sms) { uint64_t g = (n + BLOCK - 1) / BLOCK; uint64_t cap = (uint64_t)sms * 16; return (unsigned)(g < cap ? (g ? g : 1) : cap); }
}  // namespace spectrum

static void spectrum_free(ntg_spectrum* sp) {
    if (!sp) return;
    if (sp->ctx) cudaSetDevice(sp->ctx->device);
    cudaFree(sp->d_dense); cudaFree(sp->d_keys); cudaFree(sp->d_counts); cudaFree(sp->d_overflow);
    delete sp;
}
static int spectrum_clear(ntg_spectrum* sp) {
    ntg_ctx* ctx = sp->ctx;
    if (sp->dense) NTG_CUDA(ctx, cudaMemsetAsync(sp->d_dense, 0, sp->capacity * sizeof(uint32_t), ctx->stream));
    else {
        NTG_CUDA(ctx, cudaMemsetAsync(sp->d_keys, 0xFF, sp->capacity * sizeof(unsigned long long), ctx->stream));
        NTG_CUDA(ctx, cudaMemsetAsync(sp->d_counts, 0, sp->capacity * sizeof(uint32_t), ctx->stream));
    }
    NTG_CUDA(ctx, cudaMemsetAsync(sp->d_overflow, 0, sizeof(uint32_t), ctx->stream));
    sp->n_added = 0;
    return NTG_OK;
}
static int spectrum_create(ntg_ctx* ctx, uint32_t k, uint64_t capacity, ntg_spectrum** out) {
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output");
    *out = nullptr;
    if (k == 0 || k > 32) return ntg_set_error(ctx, NTG_EINVAL, "spectrum: 1 <= k <= 32 (the bit form of a k-mer is a u64, src/bitkmer.rs:2-3)");
    NTG_TRY(fused_init(ctx));
    auto* sp = new ntg_spectrum();
    sp->ctx = ctx; sp->k = k;
    sp->dense = k <= 14;
    if (sp->dense) sp->capacity = uint64_t(1) << (2 * k);
    else {
        uint64_t c = 1024;
        while (c < capacity) c <<= 1;
        sp->capacity = c;
    }
    cudaError_t e = cudaMalloc((void**)&sp->d_overflow, 2 * sizeof(uint32_t));      // [0] dropped k-mers, [1] scratch of ntg_spectrum_count
    if (!e) e = sp->dense ? cudaMalloc((void**)&sp->d_dense, sp->capacity * sizeof(uint32_t)) : cudaMalloc((void**)&sp->d_keys, sp->capacity * sizeof(unsigned long long));
    if (!e && !sp->dense) e = cudaMalloc((void**)&sp->d_counts, sp->capacity * sizeof(uint32_t));
    if (e) { cudaGetLastError(); spectrum_free(sp); return ntg_set_error(ctx, NTG_ENOMEM, "spectrum table of %llu slots does not fit", (unsigned long long)sp->capacity); }
    int st = spectrum_clear(sp);
    if (st != NTG_OK) { spectrum_free(sp); return st; }
    *out = sp;
    return NTG_OK;
}

// Count the canonical k-mers of one FASTX input.  `run(n_eff, spec, &result)` makes one pass over its first n_eff bytes.
template <typename Run>
static int spectrum_add_input(ntg_spectrum* sp, uint64_t n, int format, Run&& run, ntg_tallies* tallies, ntg_parse_error* err) {
    ntg_ctx* ctx = sp->ctx; FusedState* st = ctx->fused;
    if (err) { std::memset(err, 0, sizeof(*err)); err->format = format; }
    // (1) validating pass with the fast kernels: clean, or the start E of the first failing record
    PassResult v;
    NTG_TRY(run(n, true, &v));
    if (v.flags & fused::FLAG_SPEC_MISS) NTG_TRY(run(n, false, &v));
    uint64_t n_eff = n;
    if (v.flags == fused::FLAG_PARSE_ERROR && format == NTG_FMT_FASTQ && v.err_key != ~0ull) {
        n_eff = v.err_key >> 2;
        if (err) err->kind = NTG_EUNEXPECTED_END;                 // (the precise kind / line comes from ntg_tally_fastx; here: "stopped early")
    } else if (v.flags == fused::FLAG_PARSE_ERROR && format == NTG_FMT_FASTA) {
        if (err) { err->kind = (int32_t)v.fin[0]; err->line = v.fin[1]; err->record_index = v.fin[2]; }
    } else if (v.flags) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "spectrum: this input needs the exact record-table path (flags %u)", v.flags);
    // (2) counting pass over the valid prefix: the generic walker adds every canonical k-mer it tallies
    PassResult c;
    if (n_eff >= 2) {
        st->sp_dense = sp->d_dense; st->sp_keys = sp->d_keys; st->sp_counts = sp->d_counts; st->sp_mask = sp->capacity - 1; st->sp_overflow = sp->d_overflow;
        int rc = run(n_eff, true, &c);
        if (rc == NTG_OK && (c.flags & fused::FLAG_SPEC_MISS)) rc = NTG_EUNSUPPORTED;      // (would count twice; the validating pass has ruled it out)
        st->sp_dense = nullptr; st->sp_keys = nullptr; st->sp_counts = nullptr; st->sp_mask = 0; st->sp_overflow = nullptr;
        NTG_TRY(rc);
        if (c.flags) return ntg_set_error(ctx, NTG_EUNSUPPORTED, "spectrum: counting pass flagged %u", c.flags);
    }
    uint32_t ovf = 0;
    NTG_CUDA(ctx, cudaMemcpyAsync(&ovf, sp->d_overflow, sizeof(ovf), cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ovf) return ntg_set_error(ctx, NTG_ENOMEM, "spectrum: hash table of %llu slots is full (%u k-mers dropped): create it larger", (unsigned long long)sp->capacity, ovf);
    sp->n_added += c.tallies[2];
    if (tallies) tallies_from_pass(c, tallies);
    return NTG_OK;
}

// ---- multi-GPU --------------------------------------------------------------------------------------------------------
namespace nccldyn {
typedef int (*SendRecv_t)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*Group_t)();
static SendRecv_t Send = nullptr, Recv = nullptr;
static Group_t GroupStart = nullptr, GroupEnd = nullptr;
constexpr int kUint32 = 3;
static bool load_p2p() {
    if (Send) return true;
    if (!load()) return false;
    Send = (SendRecv_t)dlsym(handle, "ncclSend"); Recv = (SendRecv_t)dlsym(handle, "ncclRecv");
    GroupStart = (Group_t)dlsym(handle, "ncclGroupStart"); GroupEnd = (Group_t)dlsym(handle, "ncclGroupEnd");
    return Send && Recv && GroupStart && GroupEnd;
}
}  // namespace nccldyn

static int spectrum_reduce(ntg_spectrum* sp) {
    ntg_ctx* ctx = sp->ctx;
    if (!ctx->nccl_comm) return ntg_set_error(ctx, NTG_EINVAL, "ntg_comm_init has not been called");
    const int world = ctx->nccl_ranks, rank = ctx->nccl_rank;
    if (sp->dense) {
        // the whole histogram in one all-reduce: every rank ends with the job-wide counts (up to 1 GiB over NVLink at k = 14)
        int r = nccldyn::AllReduce(sp->d_dense, sp->d_dense, sp->capacity, nccldyn::kUint32, nccldyn::kSum, ctx->nccl_comm, ctx->stream);
        if (r != 0) return ntg_set_error(ctx, NTG_ENCCL, "ncclAllReduce: %s", nccldyn::GetErrorString(r));
        NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return NTG_OK;
    }
    if (!nccldyn::load_p2p()) return ntg_set_error(ctx, NTG_ENCCL, "ncclSend / ncclRecv not available");
    using namespace spectrum;
    // (1) entries per owner, on every rank: one all-reduce of a world x world matrix (row = sender)
    DevBuf<unsigned long long> d_mat, d_cur;
    if (d_mat.alloc((size_t)world * world) || d_cur.alloc(world)) return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    NTG_CUDA(ctx, cudaMemsetAsync(d_mat.p, 0, (size_t)world * world * 8, ctx->stream));
    k_owner_counts<<<grid_for(sp->capacity, ctx->sm_count), BLOCK, 0, ctx->stream>>>(sp->d_keys, sp->capacity, world, d_mat.p + (size_t)rank * world);
    ctx->launches++;
    int r = nccldyn::AllReduce(d_mat.p, d_mat.p, (size_t)world * world, nccldyn::kUint64, nccldyn::kSum, ctx->nccl_comm, ctx->stream);
    if (r != 0) return ntg_set_error(ctx, NTG_ENCCL, "ncclAllReduce: %s", nccldyn::GetErrorString(r));
    std::vector<unsigned long long> mat((size_t)world * world);
    NTG_CUDA(ctx, cudaMemcpyAsync(mat.data(), d_mat.p, mat.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // (2) partition my entries by owner
    std::vector<unsigned long long> send_off(world + 1, 0), recv_off(world + 1, 0);
    for (int d = 0; d < world; d++) send_off[d + 1] = send_off[d] + mat[(size_t)rank * world + d];
    for (int s = 0; s < world; s++) recv_off[s + 1] = recv_off[s] + mat[(size_t)s * world + rank];
    const uint64_t n_send = send_off[world], n_recv = recv_off[world];
    DevBuf<unsigned long long> sk, rk; DevBuf<uint32_t> sc, rcn;
    if (sk.alloc(n_send) || sc.alloc(n_send) || rk.alloc(n_recv) || rcn.alloc(n_recv)) return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    NTG_CUDA(ctx, cudaMemcpyAsync(d_cur.p, send_off.data(), world * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_owner_scatter<<<grid_for(sp->capacity, ctx->sm_count), BLOCK, 0, ctx->stream>>>(sp->d_keys, sp->d_counts, sp->capacity, world, d_cur.p, sk.p, sc.p);
    ctx->launches++;
    // (3) exchange: every pair of ranks, keys and counts
    nccldyn::GroupStart();
    for (int peer = 0; peer < world && r == 0; peer++) {
        const uint64_t ns = send_off[peer + 1] - send_off[peer], nr = recv_off[peer + 1] - recv_off[peer];
        if (ns) { r = nccldyn::Send(sk.p + send_off[peer], ns, nccldyn::kUint64, peer, ctx->nccl_comm, ctx->stream); if (!r) r = nccldyn::Send(sc.p + send_off[peer], ns, nccldyn::kUint32, peer, ctx->nccl_comm, ctx->stream); }
        if (nr && !r) { r = nccldyn::Recv(rk.p + recv_off[peer], nr, nccldyn::kUint64, peer, ctx->nccl_comm, ctx->stream); if (!r) r = nccldyn::Recv(rcn.p + recv_off[peer], nr, nccldyn::kUint32, peer, ctx->nccl_comm, ctx->stream); }
    }
    const int r2 = nccldyn::GroupEnd();
    if (r != 0 || r2 != 0) return ntg_set_error(ctx, NTG_ENCCL, "ncclSend/ncclRecv: %s", nccldyn::GetErrorString(r ? r : r2));
    // (4) my table becomes the job-wide counts of the keys I own
    const uint64_t added = sp->n_added;
    NTG_TRY(spectrum_clear(sp));
    sp->n_added = added;
    if (n_recv) {
        k_insert<<<grid_for(n_recv, ctx->sm_count), BLOCK, 0, ctx->stream>>>(rk.p, rcn.p, n_recv, sp->d_keys, sp->d_counts, sp->capacity - 1, sp->d_overflow);
        ctx->launches++;
    }
    uint32_t ovf = 0;
    NTG_CUDA(ctx, cudaMemcpyAsync(&ovf, sp->d_overflow, sizeof(ovf), cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ovf) return ntg_set_error(ctx, NTG_ENOMEM, "spectrum: hash table too small for the keys this rank owns");
    return NTG_OK;
}
