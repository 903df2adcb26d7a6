// Microbenchmark: how fast does ONE warp run a sequential float sum fed from shared memory?
// (the inner loop of nj_sums).  Prints cycles per row for a few shapes.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void chain(const float *in, float *out, long long *cyc, int rows, int reps) {
    extern __shared__ float tile[];
    for (int k = threadIdx.x; k < rows * 32; k += blockDim.x) tile[k] = in[k % 4096];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *col = tile + lane;
    float s = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {            // LDS + dependent FADD per row, one column per lane
#pragma unroll 16
            for (int a = 0; a < rows; ++a) s = __fadd_rn(s, col[a * 32]);
        } else if (MODE == 1) {     // register-only dependent chain
#pragma unroll 16
            for (int a = 0; a < rows; ++a) s = __fadd_rn(s, 1.0001f);
        } else {                    // LDS.128 + four independent chains
            const float4 *c4 = reinterpret_cast<const float4 *>(tile) + lane;
#pragma unroll 16
            for (int a = 0; a < rows / 4; ++a) {
                const float4 x = c4[a * 32];
                s = __fadd_rn(s, x.x); s1 = __fadd_rn(s1, x.y); s2 = __fadd_rn(s2, x.z); s3 = __fadd_rn(s3, x.w);
            }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + s1 + s2 + s3;
    if (lane == 0 && warp == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float *in, *out; long long *cyc, h;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    cudaMemset(in, 0, 4096 * 4);
    const int rows = 1024, reps = 64;
    cudaFuncSetAttribute(chain<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 128);
    cudaFuncSetAttribute(chain<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 128);
    cudaFuncSetAttribute(chain<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, rows * 128);
    for (int threads : {32, 64, 128, 256, 512}) {
        chain<0><<<148, threads, rows * 128>>>(in, out, cyc, rows, reps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads/CTA %4d  LDS+FADD      %6.2f cycles/row\n", threads, (double)h / (rows * reps));
        chain<1><<<148, threads, rows * 128>>>(in, out, cyc, rows, reps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads/CTA %4d  FADD only     %6.2f cycles/row\n", threads, (double)h / (rows * reps));
        chain<2><<<148, threads, rows * 128>>>(in, out, cyc, rows, reps); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads/CTA %4d  LDS.128+4FADD %6.2f cycles/row-of-4-cols (x%d rows)\n", threads, (double)h / (rows / 4 * reps), rows / 4);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
