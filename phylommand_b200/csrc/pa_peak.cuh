// pa_peak.cuh -- INT32 issue-rate micro-benchmarks (measurement support).
//
// The DP is bound by the integer pipes, and MEASURED_PEAKS.json only holds HBM
// and bf16 figures, so the roofline denominator is measured here: 8 independent
// chains per thread of one instruction class, every SM full.  The SASS of each
// instantiation is checked with cuobjdump (profiles/r01_peak_sass.txt).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pa {

constexpr int PEAK_CHAINS = 8;
constexpr int PEAK_UNROLL = 8;

template <int OP>
__device__ __forceinline__ void peak_step(int (&v)[PEAK_CHAINS], const int b, const int c) {
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) {
        const int w = v[(k + 1) & (PEAK_CHAINS - 1)];
        if (OP == 0) v[k] = v[k] + b + w;                                             // IADD3
        else if (OP == 1) v[k] = __vimax3_s32(v[k], w, b) ^ 0;                        // VIMNMX3
        else if (OP == 2) v[k] = __viaddmax_s32(v[k], b, w);                          // VIADDMNMX
        else if (OP == 3) v[k] = v[k] * b + w;                                        // IMAD
        else if (OP == 4) { unsigned d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v[k]), "r"(w), "r"(b)); v[k] = (int)d; }
        else if (OP == 5) v[k] = (v[k] >= w) ? b : c;                                 // ISETP + SEL
        else if (OP == 6) { if (k & 1) v[k] = v[k] * b + w; else v[k] = v[k] + b + w; }   // IMAD / IADD3 mix
        else if (OP == 7) v[k] = (int)__vimax3_s16x2((unsigned)v[k], (unsigned)w, (unsigned)b);
        else if (OP == 8) v[k] = (int)__viaddmax_s16x2((unsigned)v[k], (unsigned)b, (unsigned)w);
        else if (OP == 9) v[k] = (int)__vadd2((unsigned)v[k], (unsigned)w);
        else if (OP == 10) v[k] = (v[k] & b) ^ w;                                     // LOP3
        else if (OP == 11) v[k] = __shfl_up_sync(0xffffffffu, v[k], 1);               // SHFL
        else if (OP == 12) v[k] = max(v[k], w + 0) ;                                  // VIMNMX
        else if (OP == 13) { if (k & 1) v[k] = v[k] * b + w; else v[k] = __viaddmax_s32(v[k], b, w); }  // IMAD / VIADDMNMX mix
        // one register source + immediates: is the 64 lanes/clk/SM of the 3-register forms a pipe or an operand limit?
        else if (OP == 14) v[k] = v[k] + 3;                                           // VIADD imm
        else if (OP == 15) v[k] = v[k] * 3 + 7;                                       // IMAD imm
        else if (OP == 16) { if (k & 1) v[k] = v[k] * 3 + 7; else v[k] = v[k] + 3; }  // VIADD imm / IMAD imm mix
        else if (OP == 17) { if (k & 1) v[k] = v[k] * 3 + 7; else v[k] = max(v[k], w); }   // VIMNMX / IMAD imm mix
        else if (OP == 18) { if (k & 1) v[k] = v[k] * 3 + 7; else v[k] = __viaddmax_s32(v[k], 5, w); }  // VIADDMNMX / IMAD imm mix
        else if (OP == 19) { if ((k & 3) == 3) v[k] = v[k] * 3 + 7; else v[k] = __viaddmax_s32(v[k], 5, w); }  // 3:1
        else if (OP == 20) { unsigned d; asm volatile("prmt.b32 %0, %1, %2, 0x3210;" : "=r"(d) : "r"(v[k]), "r"(w)); v[k] = (int)d ^ 1; } // PRMT + LOP3
        else if (OP == 21) { if (k & 1) v[k] = v[k] * 3 + 7; else v[k] = (int)__viaddmax_s16x2((unsigned)v[k], 0x00050005u, (unsigned)w); }
    }
}

template <int OP>
__global__ void __launch_bounds__(256) peak_kernel(int *out, const int iters, const int b, const int c, long long *clk) {
    int v[PEAK_CHAINS];
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) v[k] = (int)threadIdx.x * (k + 3) + b;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < PEAK_UNROLL; ++u) peak_step<OP>(v, b + (OP == 4 ? 0 : 0), c);
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int k = 0; k < PEAK_CHAINS; ++k) s ^= v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    // SM 0's cycle counter from its first CTA start to its last CTA end (warps are not scheduled fairly)
    if (smid == 0 && threadIdx.x == 0) { atomicMin((unsigned long long *)&clk[0], (unsigned long long)t0); atomicMax((unsigned long long *)&clk[1], (unsigned long long)t1); }
}

template <int OP>
inline cudaError_t run_peak_op(int n_sm, cudaStream_t st, double *gops, double *mhz) {
    const int threads = 256, blocks = n_sm * 8, iters = 4096;
    int *out = nullptr;
    long long *clk = nullptr;
    cudaError_t e = cudaMalloc(&out, (size_t)threads * blocks * sizeof(int));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&clk, 2 * sizeof(long long));
    if (e != cudaSuccess) { cudaFree(out); return e; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int b = (OP == 4) ? 0x3210 : 3, c = 5;
    float best = 1e30f;
    long long hclk = 0;
    for (int rep = 0; rep < 4; ++rep) {
        const long long init[2] = {0x7fffffffffffffffll, 0};
        cudaMemcpyAsync(clk, init, sizeof init, cudaMemcpyHostToDevice, st);
        cudaEventRecord(e0, st);
        peak_kernel<OP><<<blocks, threads, 0, st>>>(out, iters, b, c, clk);
        cudaEventRecord(e1, st);
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) { best = ms; long long h2[2]; cudaMemcpy(h2, clk, sizeof h2, cudaMemcpyDeviceToHost); hclk = h2[1] - h2[0]; }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out); cudaFree(clk);
    if (e != cudaSuccess) return e;
    const double per_thread = (double)iters * PEAK_UNROLL * PEAK_CHAINS * ((OP == 5 || OP == 20) ? 2.0 : 1.0);
    *gops = per_thread * threads * blocks / (best * 1e-3) / 1e9;
    // block 0's cycle count over (roughly) the whole kernel: with 8 CTAs per SM all resident it spans the launch
    *mhz = (double)hclk / (best * 1e-3) / 1e6;
    return cudaSuccess;
}

inline cudaError_t run_peak(int which, int n_sm, cudaStream_t st, double *gops, double *mhz) {
    switch (which) {
    case 0: return run_peak_op<0>(n_sm, st, gops, mhz);
    case 1: return run_peak_op<1>(n_sm, st, gops, mhz);
    case 2: return run_peak_op<2>(n_sm, st, gops, mhz);
    case 3: return run_peak_op<3>(n_sm, st, gops, mhz);
    case 4: return run_peak_op<4>(n_sm, st, gops, mhz);
    case 5: return run_peak_op<5>(n_sm, st, gops, mhz);
    case 6: return run_peak_op<6>(n_sm, st, gops, mhz);
    case 7: return run_peak_op<7>(n_sm, st, gops, mhz);
    case 8: return run_peak_op<8>(n_sm, st, gops, mhz);
    case 9: return run_peak_op<9>(n_sm, st, gops, mhz);
    case 10: return run_peak_op<10>(n_sm, st, gops, mhz);
    case 11: return run_peak_op<11>(n_sm, st, gops, mhz);
    case 12: return run_peak_op<12>(n_sm, st, gops, mhz);
    case 13: return run_peak_op<13>(n_sm, st, gops, mhz);
    case 14: return run_peak_op<14>(n_sm, st, gops, mhz);
    case 15: return run_peak_op<15>(n_sm, st, gops, mhz);
    case 16: return run_peak_op<16>(n_sm, st, gops, mhz);
    case 17: return run_peak_op<17>(n_sm, st, gops, mhz);
    case 18: return run_peak_op<18>(n_sm, st, gops, mhz);
    case 19: return run_peak_op<19>(n_sm, st, gops, mhz);
    case 20: return run_peak_op<20>(n_sm, st, gops, mhz);
    case 21: return run_peak_op<21>(n_sm, st, gops, mhz);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace pa
