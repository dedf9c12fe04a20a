// pipes.cu — issue-rate microbenchmark for the SASS ops the walker is made of (which pipe, how many cycles per warp instruction).
// One CTA of 512 threads per SM (4 warps per SMSP), 8 independent chains per thread.  Prints cycles per warp-instruction per SMSP
// for each op alone and for pairs (a pair that costs the SUM of its parts shares a pipe; the MAX means different pipes).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 512
#define OPS(X) X(SEL) X(FSEL) X(LOP3) X(SHF) X(IADD3) X(IMAD) X(IMADW) X(PRMT) X(VIMNMX) X(VIMNMX3) X(POPC) X(FADD) X(FFMA) X(SHL) X(DSETPSEL) X(ISETPSEL) X(DADD) X(LEA) X(IMADHI) X(BREV) X(PADD) X(FSETPSEL) X(NONE)
enum Op {
#define E(n) n,
OPS(E)
#undef E
NOPS };
static const char* names[] = {
#define E(n) #n,
OPS(E)
#undef E
};
template <int OP> __device__ __forceinline__ void one(uint32_t& d, uint32_t b, uint32_t c, unsigned long long& w, double& dd, double db) {
    if (OP == SEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(d) : "r"(b), "r"(c));
    if (OP == FSEL) asm volatile("{.reg .pred p; .reg .f32 x, y; setp.ne.u32 p, %2, 0; mov.b32 x, %0; mov.b32 y, %1; selp.f32 x, x, y, p; mov.b32 %0, x;}" : "+r"(d) : "r"(b), "r"(c));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d) : "r"(b), "r"(c));
    if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(d) : "r"(b));
    if (OP == IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(d) : "r"(b), "r"(c));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(b), "r"(c));
    if (OP == IMADW) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(b), "r"(c));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(d) : "r"(b), "r"(c));
    if (OP == VIMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(d) : "r"(b));
    if (OP == VIMNMX3) asm volatile("{.reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(d) : "r"(b), "r"(c));
    if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(d));
    if (OP == FADD) asm volatile("{.reg .f32 x, y; mov.b32 x, %0; mov.b32 y, %1; add.f32 x, x, y; mov.b32 %0, x;}" : "+r"(d) : "r"(b));
    if (OP == FFMA) asm volatile("{.reg .f32 x, y, z; mov.b32 x, %0; mov.b32 y, %1; mov.b32 z, %2; fma.rn.f32 x, x, y, z; mov.b32 %0, x;}" : "+r"(d) : "r"(b), "r"(c));
    if (OP == SHL) asm volatile("shl.b32 %0, %0, 2;" : "+r"(d));
    if (OP == DSETPSEL) asm volatile("{.reg .pred p; .reg .f64 x; mov.b64 x, {%0, %1}; setp.lt.f64 p, x, %2; selp.b32 %0, %0, %1, p;}" : "+r"(d) : "r"(b), "d"(db));
    if (OP == ISETPSEL) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %2; selp.b32 %0, %0, %1, p;}" : "+r"(d) : "r"(b), "r"(c));
    if (OP == FSETPSEL) asm volatile("{.reg .pred p; .reg .f32 x, y; mov.b32 x, %0; mov.b32 y, %2; setp.lt.f32 p, x, y; selp.b32 %0, %0, %1, p;}" : "+r"(d) : "r"(b), "r"(c));
    if (OP == DADD) asm volatile("add.f64 %0, %0, %1;" : "+d"(dd) : "d"(db));
    if (OP == LEA) asm volatile("{.reg .u32 t; shl.b32 t, %0, 2; add.u32 %0, t, %1;}" : "+r"(d) : "r"(b));
    if (OP == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(b), "r"(c));
    if (OP == BREV) asm volatile("brev.b32 %0, %0;" : "+r"(d));
    if (OP == PADD) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; @p add.u32 %0, %0, 1;}" : "+r"(d) : "r"(c));
}
template <int A, int B> __global__ void __launch_bounds__(512, 1) k(uint32_t* out, uint32_t b, uint32_t c, double db, long long* cyc) {
    uint32_t d[8]; unsigned long long w[8]; double dd[8];
    uint32_t e[8]; unsigned long long w2[8]; double dd2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { d[i] = threadIdx.x * 77u + i; w[i] = d[i]; dd[i] = (double)d[i]; e[i] = d[i] ^ 0x55u; w2[i] = e[i]; dd2[i] = (double)e[i]; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { one<A>(d[i], b, c, w[i], dd[i], db); one<B>(e[i], b, c, w2[i], dd2[i], db); }
    }
    const long long t1 = clock64();
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i] + (uint32_t)w[i] + (uint32_t)dd[i] + e[i] + (uint32_t)w2[i] + (uint32_t)dd2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int A, int B> double run(uint32_t* out, long long* cyc, int sms) {
    k<A, B><<<sms, 512>>>(out, 3u, 1u, 1.5, cyc);
    k<A, B><<<sms, 512>>>(out, 3u, 1u, 1.5, cyc);
    cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < sms; i++) s += (double)h[i];
    // per SMSP: 4 warps x ITER x 8 chain-steps (each step = one A-group + one B-group)
    return s / sms / (4.0 * ITER * 8);
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc; cudaMalloc(&out, sms * 512 * 4); cudaMalloc(&cyc, 1024 * 8);
    printf("cycles per (A-group + B-group) per warp per SMSP; groups: IADD3/VIMNMX3/LEA = the fused form if ptxas fuses; *SETPSEL = setp + sel\n");
#define S(A) printf("%-10s alone %.2f\n", names[A], run<A, NONE>(out, cyc, sms));
    OPS(S)
#define P(A, B) printf("%-10s + %-10s %.2f\n", names[A], names[B], run<A, B>(out, cyc, sms));
    P(SEL, FSEL) P(SEL, LOP3) P(SEL, IMAD) P(FSEL, IMAD) P(FSEL, LOP3) P(FSEL, FADD) P(LOP3, IMAD) P(LOP3, SHF) P(SHF, IMAD) P(LOP3, FADD) P(LOP3, FFMA)
    P(IMAD, FFMA) P(IMADW, LOP3) P(IMADW, IMAD) P(DSETPSEL, LOP3) P(DSETPSEL, IMAD) P(DADD, LOP3) P(DADD, IMAD) P(DADD, DSETPSEL) P(PRMT, LOP3) P(PRMT, IMAD)
    P(VIMNMX, LOP3) P(VIMNMX, IMAD) P(POPC, LOP3) P(POPC, IMAD) P(SHL, LOP3) P(SHL, IMAD) P(LEA, LOP3) P(LEA, IMAD) P(BREV, LOP3) P(BREV, POPC) P(IMADHI, IMAD) P(IMADHI, LOP3)
    P(FSETPSEL, LOP3) P(FSETPSEL, IMAD) P(ISETPSEL, IMAD) P(PADD, IMAD) P(PADD, LOP3)
    return 0;
}
