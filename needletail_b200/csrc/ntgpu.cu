// ntgpu.cu — libntgpu's single translation unit (unity build) and the C ABI of include/ntgpu.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
//        (see needletail_b200/build.py).  No torch, no CPU compute fallback: without a CUDA device
//        ntg_create fails and every entry point needs a context.
#include <cstdarg>
#include <cstdlib>
#include <dlfcn.h>

#include "common.cuh"
#include "luts.cuh"
#include "scan.cuh"
#include "seqops.cuh"
#include "parse.cuh"
#include "fused.cuh"
#include "fused_host.cuh"
#include "synth.cuh"

int ntg_set_error(ntg_ctx* ctx, int status, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return status;
}

// ------------------------------------------------------------------------------ NCCL (lazy dlopen)
namespace nccldyn {
typedef struct { char internal[NTG_NCCL_ID_BYTES]; } UniqueId;
typedef int (*GetUniqueId_t)(UniqueId*);
typedef int (*CommInitRank_t)(void**, int, UniqueId, int);
typedef int (*AllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*CommDestroy_t)(void*);
typedef const char* (*GetErrorString_t)(int);
static void* handle = nullptr;
static GetUniqueId_t GetUniqueId = nullptr;
static CommInitRank_t CommInitRank = nullptr;
static AllReduce_t AllReduce = nullptr;
static CommDestroy_t CommDestroy = nullptr;
static GetErrorString_t GetErrorString = nullptr;
constexpr int kUint64 = 5, kSum = 0;     // ncclUint64, ncclSum (nccl.h)
static bool load() {
    if (handle) return true;
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) return false;
    GetUniqueId = (GetUniqueId_t)dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (CommInitRank_t)dlsym(handle, "ncclCommInitRank");
    AllReduce = (AllReduce_t)dlsym(handle, "ncclAllReduce");
    CommDestroy = (CommDestroy_t)dlsym(handle, "ncclCommDestroy");
    GetErrorString = (GetErrorString_t)dlsym(handle, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy && GetErrorString;
}
}  // namespace nccldyn

extern "C" {

int ntg_abi_version(void) { return NTG_ABI_VERSION; }

int ntg_device_count(int* count) {
    if (!count) return NTG_EINVAL;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { *count = 0; cudaGetLastError(); return NTG_ECUDA; }
    *count = c;
    return NTG_OK;
}

int ntg_create(int device, ntg_ctx** out) {
    if (!out) return NTG_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return NTG_ECUDA; }
    if (device < 0 || device >= count) return NTG_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return NTG_ECUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NTG_ECUDA;
    if (prop.major != 10) return NTG_EUNSUPPORTED;      // sm_100a code only
    auto* ctx = new ntg_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& e : ctx->events) ok = ok && cudaEventCreate(&e) == cudaSuccess;
    if (!ok || ntg_upload_luts(ctx) != NTG_OK) { ntg_destroy(ctx); return NTG_ECUDA; }
    *out = ctx;
    return NTG_OK;
}

void ntg_destroy(ntg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ntg_comm_destroy(ctx);
    fused_destroy(ctx);
    for (auto& e : ctx->events) if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

const char* ntg_last_error(const ntg_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
uint64_t ntg_launch_count(const ntg_ctx* ctx) { return ctx ? ctx->launches : 0; }

#define CTX_ENTER(ctx)                                                        \
    if (!(ctx)) return NTG_EINVAL;                                            \
    if (cudaSetDevice((ctx)->device) != cudaSuccess) return ntg_set_error((ctx), NTG_ECUDA, "cudaSetDevice failed")

int ntg_device_info(ntg_ctx* ctx, int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor) {
    CTX_ENTER(ctx);
    cudaDeviceProp prop;
    NTG_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return NTG_OK;
}
int ntg_sync(ntg_ctx* ctx) {
    CTX_ENTER(ctx);
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return NTG_OK;
}
int ntg_alloc_pinned(size_t bytes, void** out) {
    if (!out) return NTG_EINVAL;
    return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? NTG_OK : NTG_ENOMEM;
}
int ntg_free_pinned(void* p) { return cudaFreeHost(p) == cudaSuccess ? NTG_OK : NTG_ECUDA; }
int ntg_device_alloc(ntg_ctx* ctx, size_t bytes, uint64_t* dptr) {
    CTX_ENTER(ctx);
    if (!dptr) return ntg_set_error(ctx, NTG_EINVAL, "null pointer");
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return ntg_set_error(ctx, NTG_ENOMEM, "cudaMalloc(%zu) failed", bytes); }
    *dptr = (uint64_t)(uintptr_t)p;
    return NTG_OK;
}
int ntg_device_free(ntg_ctx* ctx, uint64_t dptr) {
    CTX_ENTER(ctx);
    NTG_CUDA(ctx, cudaFree((void*)(uintptr_t)dptr));
    return NTG_OK;
}
int ntg_memcpy_h2d(ntg_ctx* ctx, uint64_t dptr, const void* host, size_t bytes) {
    CTX_ENTER(ctx);
    NTG_CUDA(ctx, cudaMemcpyAsync((void*)(uintptr_t)dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NTG_OK;
}
int ntg_memcpy_d2h(ntg_ctx* ctx, void* host, uint64_t dptr, size_t bytes) {
    CTX_ENTER(ctx);
    NTG_CUDA(ctx, cudaMemcpyAsync(host, (const void*)(uintptr_t)dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NTG_OK;
}
int ntg_event_record(ntg_ctx* ctx, int slot) {
    CTX_ENTER(ctx);
    if (slot < 0 || slot >= 64) return ntg_set_error(ctx, NTG_EINVAL, "event slot out of range");
    NTG_CUDA(ctx, cudaEventRecord(ctx->events[slot], ctx->stream));
    return NTG_OK;
}
int ntg_event_elapsed_ms(ntg_ctx* ctx, int a, int b, float* ms) {
    CTX_ENTER(ctx);
    if (a < 0 || a >= 64 || b < 0 || b >= 64 || !ms) return ntg_set_error(ctx, NTG_EINVAL, "bad event arguments");
    NTG_CUDA(ctx, cudaEventSynchronize(ctx->events[b]));
    NTG_CUDA(ctx, cudaEventElapsedTime(ms, ctx->events[a], ctx->events[b]));
    return NTG_OK;
}

// ---- (1) record scanner ------------------------------------------------------------------------
int ntg_parse_fastx(ntg_ctx* ctx, const uint8_t* bytes, size_t n, ntg_records** out) {
    CTX_ENTER(ctx);
    return run_parse_device(ctx, bytes, nullptr, n, out, nullptr);
}
void ntg_records_free(ntg_records* r) {
    if (!r) return;
    delete static_cast<RecordsPriv*>(r->_priv);
    delete r;
}

// ---- (2) Sequence trait, batch form ------------------------------------------------------------
int ntg_normalize(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, int allow_iupac,
                  uint8_t* out, uint64_t* out_offs, uint8_t* changed) {
    CTX_ENTER(ctx);
    return run_xform(ctx, seqs, offs, n, allow_iupac ? 1 : 0, out, out_offs, changed);
}
int ntg_strip_returns(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, uint8_t* out, uint64_t* out_offs, uint8_t* changed) {
    CTX_ENTER(ctx);
    return run_xform(ctx, seqs, offs, n, 2, out, out_offs, changed);
}
int ntg_reverse_complement(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, uint8_t* out) {
    CTX_ENTER(ctx);
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output pointer");
    BatchOnDevice b;
    NTG_TRY(upload_batch(ctx, seqs, offs, n, b));
    if (!b.total) return NTG_OK;
    DevBuf<uint8_t> dout;
    if (dout.alloc(b.total)) return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    seqops::k_revcomp<<<seqops::grid_for(b.total), seqops::BLOCK, 0, ctx->stream>>>(b.seqs.p, b.offs.p, b.nseq, b.total, dout.p);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    NTG_CUDA(ctx, cudaMemcpyAsync(out, dout.p, b.total, cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NTG_OK;
}
int ntg_quality_mask(ntg_ctx* ctx, const uint8_t* seqs, const uint8_t* quals, const uint64_t* offs, size_t n, uint8_t score, uint8_t* out) {
    CTX_ENTER(ctx);
    if (!out || !quals) return ntg_set_error(ctx, NTG_EINVAL, "null pointer");
    BatchOnDevice b;
    NTG_TRY(upload_batch(ctx, seqs, offs, n, b));
    if (!b.total) return NTG_OK;
    DevBuf<uint8_t> dq, dout;
    if (dq.alloc(b.total) || dout.alloc(b.total)) return ntg_set_error(ctx, NTG_ENOMEM, "device allocation failed");
    NTG_CUDA(ctx, cudaMemcpyAsync(dq.p, quals, b.total, cudaMemcpyHostToDevice, ctx->stream));
    seqops::k_qmask<<<seqops::grid_for(b.total), seqops::BLOCK, 0, ctx->stream>>>(b.seqs.p, dq.p, b.total, score, dout.p);
    ctx->launches++;
    NTG_CUDA(ctx, cudaGetLastError());
    NTG_CUDA(ctx, cudaMemcpyAsync(out, dout.p, b.total, cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NTG_OK;
}
void ntg_items_free(ntg_items* it) {
    if (!it) return;
    delete static_cast<ItemsPriv*>(it->_priv);
    delete it;
}
int ntg_canonical_kmers(ntg_ctx* ctx, const uint8_t* seqs, const uint8_t* rc, const uint64_t* offs, size_t n, uint32_t k, ntg_items** out) {
    CTX_ENTER(ctx);
    return run_kmers(ctx, seqs, rc, offs, n, k, 0, 0, out);
}
int ntg_bit_kmers(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, uint32_t k, int canonical, ntg_items** out) {
    CTX_ENTER(ctx);
    return run_kmers(ctx, seqs, nullptr, offs, n, k, 0, canonical ? 2 : 1, out);
}
int ntg_bit_minimizers(ntg_ctx* ctx, const uint8_t* seqs, const uint64_t* offs, size_t n, uint32_t k, uint32_t m, ntg_items** out) {
    CTX_ENTER(ctx);
    return run_kmers(ctx, seqs, nullptr, offs, n, k, m, 3, out);
}
int ntg_bitkmer_reverse_complement(ntg_ctx* ctx, const uint64_t* in, size_t n, uint32_t k, uint64_t* out) {
    CTX_ENTER(ctx);
    return run_bitkmer_elem(ctx, in, n, k, 0, 0, out, nullptr);
}
int ntg_bitkmer_canonical(ntg_ctx* ctx, const uint64_t* in, size_t n, uint32_t k, uint64_t* out, uint8_t* was_rc) {
    CTX_ENTER(ctx);
    return run_bitkmer_elem(ctx, in, n, k, 0, 1, out, was_rc);
}
int ntg_bitkmer_minimizer(ntg_ctx* ctx, const uint64_t* in, size_t n, uint32_t k, uint32_t m, uint64_t* out) {
    CTX_ENTER(ctx);
    return run_bitkmer_elem(ctx, in, n, k, m, 2, out, nullptr);
}

// ---- (3) fused hot path ------------------------------------------------------------------------
int ntg_tally_fastx_device_enqueue(ntg_ctx* ctx, uint64_t dptr, size_t n, const ntg_tally_config* cfg) {
    CTX_ENTER(ctx);
    NTG_TRY(check_tally_cfg(ctx, cfg));
    NTG_TRY(fused_init(ctx));
    FusedState* st = ctx->fused;
    if (n < 2 || !dptr) return ntg_set_error(ctx, NTG_EINVAL, "enqueue needs >= 2 device-resident bytes (use ntg_tally_fastx_device for the sniff rules)");
    // sniff + tile sizing from the first 64 KiB (one small D2H; the stream is otherwise untouched)
    static thread_local std::vector<uint8_t> sample;
    const size_t ns = n < 65536 ? n : 65536;
    sample.resize(ns);
    NTG_CUDA(ctx, cudaMemcpyAsync(sample.data(), (const void*)(uintptr_t)dptr, ns, cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint8_t b0 = sample[0];
    int format = b0 == '>' ? NTG_FMT_FASTA : (b0 == '@' ? NTG_FMT_FASTQ : NTG_FMT_NONE);
    if (format == NTG_FMT_NONE) return ntg_set_error(ctx, NTG_EUNKNOWN_FORMAT, "first byte is neither '>' nor '@'");
    NTG_TRY(fused_begin(ctx, (const uint8_t*)(uintptr_t)dptr, n, format, cfg, pick_tile_bytes(sample.data(), ns, format)));
    st->host_bytes = nullptr;
    NTG_CUDA(ctx, cudaEventRecord(st->ev_k0, ctx->stream));
    int s = fused_launch(ctx, 0, st->P.num_tiles, 0);
    if (s == NTG_OK) { cudaEventRecord(st->ev_k1, ctx->stream); s = fused_finish_enqueue(ctx); }
    if (s != NTG_OK) st->pending = false;
    return s;
}
int ntg_tally_fastx_device_collect(ntg_ctx* ctx, ntg_tallies* out, ntg_parse_error* err, float* fused_kernel_ms) {
    CTX_ENTER(ctx);
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output");
    return fused_collect(ctx, out, err, fused_kernel_ms);
}
int ntg_tally_fastx_device(ntg_ctx* ctx, uint64_t dptr, size_t n, const ntg_tally_config* cfg, ntg_tallies* out, ntg_parse_error* err) {
    CTX_ENTER(ctx);
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output");
    NTG_TRY(check_tally_cfg(ctx, cfg));
    uint8_t b0 = 0;
    if (n >= 1) {
        if (!dptr) return ntg_set_error(ctx, NTG_EINVAL, "null device pointer");
        NTG_CUDA(ctx, cudaMemcpy(&b0, (const void*)(uintptr_t)dptr, 1, cudaMemcpyDeviceToHost));
    }
    int format;
    if (sniff_format(ctx, b0, n, out, err, &format)) return NTG_OK;
    NTG_TRY(ntg_tally_fastx_device_enqueue(ctx, dptr, n, cfg));
    return fused_collect(ctx, out, err, nullptr);
}

// Host bytes: the device copy is fed in chunks on the copy stream; the fused kernel is launched over
// each chunk's tiles as soon as the chunk has landed (look-back state carries across launches).
int ntg_tally_fastx(ntg_ctx* ctx, const uint8_t* bytes, size_t n, const ntg_tally_config* cfg, ntg_tallies* out, ntg_parse_error* err) {
    CTX_ENTER(ctx);
    if (!out) return ntg_set_error(ctx, NTG_EINVAL, "null output");
    NTG_TRY(check_tally_cfg(ctx, cfg));
    if (n && !bytes) return ntg_set_error(ctx, NTG_EINVAL, "null input");
    int format;
    if (sniff_format(ctx, n ? bytes[0] : 0, n, out, err, &format)) return NTG_OK;
    NTG_TRY(fused_init(ctx));
    FusedState* st = ctx->fused;
    if (st->feed_cap < n) {
        cudaFree(st->feed_buf); st->feed_buf = nullptr; st->feed_cap = 0;
        size_t cap = (n + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);
        if (cudaMalloc((void**)&st->feed_buf, cap) != cudaSuccess) { cudaGetLastError(); return ntg_set_error(ctx, NTG_ENOMEM, "cudaMalloc(%zu) failed", cap); }
        st->feed_cap = cap;
    }
    const uint32_t tile_bytes = pick_tile_bytes(bytes, n < 65536 ? n : 65536, format);
    NTG_TRY(fused_begin(ctx, st->feed_buf, n, format, cfg, tile_bytes));
    st->host_bytes = bytes;
    // chunk size: a multiple of the tile, at most FUSED_MAX_LAUNCHES chunks
    const uint64_t num_tiles = st->P.num_tiles;
    const uint32_t tbytes = st->P.tile_bytes;             // (the warp-specialised kernel has its own tile size)
    uint64_t tiles_per_chunk = ((size_t(256) << 20) + tbytes - 1) / tbytes;
    if ((num_tiles + tiles_per_chunk - 1) / tiles_per_chunk > FUSED_MAX_LAUNCHES)
        tiles_per_chunk = (num_tiles + FUSED_MAX_LAUNCHES - 1) / FUSED_MAX_LAUNCHES;
    // the copy stream must not overwrite feed_buf while an earlier call's kernels still read it, and the
    // control block reset (compute stream) must precede the first launch: both are stream-ordered here.
    cudaEvent_t ev_ready = st->ev_ready;
    int s = NTG_OK;
    cudaError_t e = cudaEventRecord(ev_ready, ctx->stream);
    if (!e) e = cudaStreamWaitEvent(ctx->copy_stream, ev_ready, 0);
    if (!e) e = cudaEventRecord(st->ev_k0, ctx->stream);
    int li = 0;
    for (uint64_t tb = 0; tb < num_tiles && !e && s == NTG_OK; tb += tiles_per_chunk, li++) {
        const uint64_t te = tb + tiles_per_chunk < num_tiles ? tb + tiles_per_chunk : num_tiles;
        const size_t b0 = tb * (uint64_t)tbytes, b1 = te * (uint64_t)tbytes < n ? te * (uint64_t)tbytes : n;
        e = cudaMemcpyAsync(st->feed_buf + b0, bytes + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (!e) e = cudaEventRecord(st->ev_chunk[li], ctx->copy_stream);
        if (!e) e = cudaStreamWaitEvent(ctx->stream, st->ev_chunk[li], 0);
        if (!e) s = fused_launch(ctx, tb, te, li);
    }
    if (!e) e = cudaEventRecord(st->ev_k1, ctx->stream);
    if (e) { st->pending = false; return ntg_set_error(ctx, NTG_ECUDA, "feed: %s", cudaGetErrorString(e)); }
    if (s == NTG_OK) s = fused_finish_enqueue(ctx);
    if (s != NTG_OK) { st->pending = false; return s; }
    return fused_collect(ctx, out, err, nullptr);
}

// ---- (4) synthetic inputs ----------------------------------------------------------------------
int ntg_synth_fastq_device(ntg_ctx* ctx, uint64_t dptr, uint64_t seed, uint64_t rec0, uint64_t nrec, uint32_t read_len, uint32_t n_thresh) {
    CTX_ENTER(ctx);
    return run_synth(ctx, dptr, seed, rec0, nrec, read_len, n_thresh, 1);
}
int ntg_synth_fasta_device(ntg_ctx* ctx, uint64_t dptr, uint64_t seed, uint64_t rec0, uint64_t nrec, uint32_t read_len, uint32_t n_thresh) {
    CTX_ENTER(ctx);
    return run_synth(ctx, dptr, seed, rec0, nrec, read_len, n_thresh, 0);
}

// ---- (5) multi-GPU -----------------------------------------------------------------------------
int ntg_comm_unique_id(uint8_t id[NTG_NCCL_ID_BYTES]) {
    if (!id) return NTG_EINVAL;
    if (!nccldyn::load()) return NTG_ENCCL;
    nccldyn::UniqueId u;
    if (nccldyn::GetUniqueId(&u) != 0) return NTG_ENCCL;
    std::memcpy(id, u.internal, NTG_NCCL_ID_BYTES);
    return NTG_OK;
}
int ntg_comm_init(ntg_ctx* ctx, int n_ranks, int rank, const uint8_t id[NTG_NCCL_ID_BYTES]) {
    CTX_ENTER(ctx);
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return ntg_set_error(ctx, NTG_EINVAL, "bad communicator arguments");
    if (!nccldyn::load()) return ntg_set_error(ctx, NTG_ENCCL, "libnccl.so.2 not loadable: %s", dlerror());
    if (ctx->nccl_comm) return ntg_set_error(ctx, NTG_EINVAL, "communicator already initialised");
    nccldyn::UniqueId u;
    std::memcpy(u.internal, id, NTG_NCCL_ID_BYTES);
    int r = nccldyn::CommInitRank(&ctx->nccl_comm, n_ranks, u, rank);
    if (r != 0) { ctx->nccl_comm = nullptr; return ntg_set_error(ctx, NTG_ENCCL, "ncclCommInitRank: %s", nccldyn::GetErrorString(r)); }
    NTG_CUDA(ctx, cudaMalloc(&ctx->nccl_buf, sizeof(ntg_tallies)));
    return NTG_OK;
}
int ntg_comm_allreduce_tallies(ntg_ctx* ctx, ntg_tallies* inout) {
    CTX_ENTER(ctx);
    if (!inout) return ntg_set_error(ctx, NTG_EINVAL, "null pointer");
    if (!ctx->nccl_comm) return ntg_set_error(ctx, NTG_EINVAL, "ntg_comm_init has not been called");
    NTG_CUDA(ctx, cudaMemcpyAsync(ctx->nccl_buf, inout, sizeof(ntg_tallies), cudaMemcpyHostToDevice, ctx->stream));
    int r = nccldyn::AllReduce(ctx->nccl_buf, ctx->nccl_buf, sizeof(ntg_tallies) / 8, nccldyn::kUint64, nccldyn::kSum, ctx->nccl_comm, ctx->stream);
    if (r != 0) return ntg_set_error(ctx, NTG_ENCCL, "ncclAllReduce: %s", nccldyn::GetErrorString(r));
    NTG_CUDA(ctx, cudaMemcpyAsync(inout, ctx->nccl_buf, sizeof(ntg_tallies), cudaMemcpyDeviceToHost, ctx->stream));
    NTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NTG_OK;
}
int ntg_comm_destroy(ntg_ctx* ctx) {
    if (!ctx) return NTG_EINVAL;
    if (ctx->nccl_comm && nccldyn::CommDestroy) { nccldyn::CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    if (ctx->nccl_buf) { cudaFree(ctx->nccl_buf); ctx->nccl_buf = nullptr; }
    return NTG_OK;
}

}  // extern "C"
