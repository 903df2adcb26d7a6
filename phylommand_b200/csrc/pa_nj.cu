// pa_nj.cu -- neighbour joining on the GPU with the arithmetic of phylommand's
// njtree::build_nj_tree (reference src/nj_tree.cpp:32-205): the consumer of the
// distance matrix pairalign -m prints (`... | treeator -n`, README.md:113).
//
// The reference keeps ragged vector<float> rows, erases two cells per row and
// two rows per join and inserts the new node at the FRONT; every join costs
// O(r^2) on one core.  What is observable is
//   * S_p: the float sum of taxon p's distances in ascending order of the others (:39-47),
//   * the FIRST strictly smallest (r-2)*d - S_p - S_q over pairs in row-major order, from 100000 (:53-74),
//   * the branch lengths (:92-94) and the new distances (d_ki + d_kj - d_ij)/2 in float (:150),
//   * the order of the taxa after the join: new node first, the others as before (:176-178).
// Here the distances live in a dense r x r float matrix kept PHYSICALLY in that order (two
// buffers, ping-pong), so every access is coalesced, and a whole tree is built without a host
// round trip: two launches per join, everything the host needs is recorded on the device.
//
//   nj_argmin  all pairs in parallel; key = (order-preserving float bits, row-major rank) reduced with
//              a 64-bit atomicMin: the smallest value wins, ties go to the first pair visited.
//   nj_join    one CTA per block of W columns of the NEXT matrix: builds those columns from the old
//              matrix (new node = row/column 0), writes them, and adds them up row by row into the next
//              S -- tiles are loaded by all 256 threads (memory parallelism), the adds of one column are
//              done by one thread in row order (the reference's summation order; float addition is not
//              associative, a tree reduction would change the bits).
// __fmul_rn/__fsub_rn/__fadd_rn/__fdiv_rn keep the compiler from contracting into FMAs: the bits
// must be the ones the reference's scalar float code produces.
// HBM-bound: per join 4 B x (r^2/2 [argmin] + r^2 read + r^2 write [join]); the add chain (4 cycles per
// row per column) stays hidden behind the loads of the other CTAs on the SM.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pairalign_b200.h"

void pa_internal_set_error(const char *msg);   // pa_capi.cu: message behind pa_last_error()

namespace {

constexpr int NJ_THREADS = 256;
constexpr int NJ_TILE = 2048;                 // floats per tile: (2048 / W) rows x W columns
constexpr float NJ_START = 100000.0f;         // "float M=100000" (src/nj_tree.cpp:53)

struct JoinRec { uint32_t left, right; float length, s_i, s_j; int r; };

__host__ __device__ __forceinline__ unsigned int ordered_bits(float v) {
    v = v + 0.0f;                              // -0 -> +0: they compare equal in the reference
#ifdef __CUDA_ARCH__
    const unsigned int u = __float_as_uint(v);
#else
    unsigned int u; memcpy(&u, &v, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__host__ __device__ __forceinline__ unsigned long long nj_sentinel() {
    return ((unsigned long long)ordered_bits(NJ_START) << 32) | 0xffffffffull;
}

// upper triangle (row-major) -> dense symmetric matrix with a zero diagonal
__global__ void nj_expand(const float *tri, float *M, const int ld, const int n) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (b >= n) return;
    float v = 0.0f;
    if (a != b) {
        const size_t lo = a < b ? a : b, hi = a < b ? b : a;
        v = tri[lo * (size_t)n - lo * (lo + 1) / 2 + (hi - lo - 1)];
    }
    M[(size_t)a * ld + b] = v;
}

// First strict minimum of (r-2)*d(p,q) - S[p] - S[q] over p < q in row-major order, below 100000.
__global__ void __launch_bounds__(NJ_THREADS) nj_argmin(const float *__restrict__ M, const int ld, const int r,
                                                        const float *__restrict__ S, unsigned long long *key) {
    const float fr2 = (float)(r - 2);
    unsigned long long best = nj_sentinel();
    for (int p = blockIdx.x; p < r - 1; p += gridDim.x) {
        const float sp = S[p];
        const float *row = M + (size_t)p * ld;
        const unsigned int base = (unsigned int)p * (unsigned int)r;
#pragma unroll 4
        for (int q = p + 1 + threadIdx.x; q < r; q += NJ_THREADS) {
            const float v = __fsub_rn(__fsub_rn(__fmul_rn(fr2, row[q]), sp), S[q]);
            if (v < NJ_START) {                       // NaN never wins, like 'value < M'
                const unsigned long long k = ((unsigned long long)ordered_bits(v) << 32) | (base + (unsigned int)q);
                best = k < best ? k : best;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    __shared__ unsigned long long sm[NJ_THREADS / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < NJ_THREADS / 32; ++w) best = sm[w] < best ? sm[w] : best;
        if (best != nj_sentinel()) atomicMin(key, best);
    }
}

// Columns [b0, b0+W) of the next round's matrix and their sums.  IDENT: first round, nothing joined
// yet -- only the sums of the matrix as uploaded.
template <int W, bool IDENT>
__global__ void __launch_bounds__(NJ_THREADS) nj_join(const float *__restrict__ M, float *__restrict__ M2, const int ld,
                                                      const int r, const float *__restrict__ S, float *__restrict__ S2,
                                                      const uint32_t *__restrict__ node, uint32_t *__restrict__ node2,
                                                      const unsigned long long *key, JoinRec *rec, const uint32_t new_id) {
    constexpr int R = NJ_THREADS / W;          // row lanes
    constexpr int TR = NJ_TILE / W;            // rows per tile
    constexpr int U = TR / R;                  // rows per thread per tile (= 8)
    __shared__ float tile[2][NJ_TILE];
    const int c = threadIdx.x % W, rl = threadIdx.x / W;
    const int b = blockIdx.x * W + c;
    int i = 0, jp = 1;                          // nothing below 100000: the reference joins its initial i = 0, j = 0
    float length = 0.0f;
    const int r2 = IDENT ? r : r - 1;
    if (!IDENT) {
        const unsigned long long k = *key;
        if (k != nj_sentinel()) { const unsigned int rank = (unsigned int)k; i = rank / (unsigned int)r; jp = rank % (unsigned int)r; }
        length = M[(size_t)i * ld + jp];
    }
    auto old_of = [&](int k) { int o = k - 1; if (o >= i) ++o; if (o >= jp) ++o; return o; };   // k >= 1
    const bool col_ok = b < r2;
    const int ob = IDENT ? b : (b > 0 ? old_of(b) : 0);
    const float *row_i = M + (size_t)i * ld, *row_j = M + (size_t)jp * ld;
    if (!IDENT) {
        if (rl == 0 && col_ok) node2[b] = b == 0 ? new_id : node[ob];
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            JoinRec jr; jr.left = node[i]; jr.right = node[jp]; jr.length = length; jr.s_i = S[i]; jr.s_j = S[jp]; jr.r = r;
            *rec = jr;
        }
    }
    auto value = [&](int a) -> float {          // element (a, b) of the next matrix
        if (IDENT) return M[(size_t)a * ld + b];
        if (a == 0) return b == 0 ? 0.0f : __fdiv_rn(__fsub_rn(__fadd_rn(row_i[ob], row_j[ob]), length), 2.0f);
        const int oa = old_of(a);
        if (b == 0) return __fdiv_rn(__fsub_rn(__fadd_rn(row_i[oa], row_j[oa]), length), 2.0f);
        return M[(size_t)oa * ld + ob];
    };
    const int n_tiles = (r2 + TR - 1) / TR;
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const int a = rl + u * R; v[u] = (col_ok && a < r2) ? value(a) : 0.0f; }
    float s = 0.0f;
    for (int t = 0; t < n_tiles; ++t) {
        float *buf = tile[t & 1];
        const int a0 = t * TR;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int a = a0 + rl + u * R;
            buf[(rl + u * R) * W + c] = v[u];
            if (!IDENT && col_ok && a < r2) M2[(size_t)a * ld + b] = v[u];
        }
        __syncthreads();
        if (t + 1 < n_tiles) {
#pragma unroll
            for (int u = 0; u < U; ++u) { const int a = a0 + TR + rl + u * R; v[u] = (col_ok && a < r2) ? value(a) : 0.0f; }
        }
        if (threadIdx.x < W) {
            // rows in ascending order; rows past the end hold +0 and the diagonal is +0: neither changes the sum
            const int rows = r2 - a0 < TR ? r2 - a0 : TR;
            for (int a = 0; a < rows; ++a) s = __fadd_rn(s, buf[a * W + c]);
        }
    }
    if (threadIdx.x < W && col_ok) S2[b] = s;
}

int g_nj_cols = 0;      // PAIRALIGN_NJ_COLS=8|16|32 forces the columns per CTA (tests, tuning)

template <bool IDENT>
void launch_join(cudaStream_t st, const float *M, float *M2, int ld, int r, const float *S, float *S2, const uint32_t *node,
                 uint32_t *node2, const unsigned long long *key, JoinRec *rec, uint32_t new_id) {
    const int r2 = IDENT ? r : r - 1;
    int w = r2 >= 148 * 4 * 32 ? 32 : r2 >= 148 * 2 * 16 ? 16 : 8;
    if (g_nj_cols == 8 || g_nj_cols == 16 || g_nj_cols == 32) w = g_nj_cols;
    if (w == 32)
        nj_join<32, IDENT><<<(r2 + 31) / 32, NJ_THREADS, 0, st>>>(M, M2, ld, r, S, S2, node, node2, key, rec, new_id);
    else if (w == 16)
        nj_join<16, IDENT><<<(r2 + 15) / 16, NJ_THREADS, 0, st>>>(M, M2, ld, r, S, S2, node, node2, key, rec, new_id);
    else
        nj_join<8, IDENT><<<(r2 + 7) / 8, NJ_THREADS, 0, st>>>(M, M2, ld, r, S, S2, node, node2, key, rec, new_id);
}

__global__ void nj_init(unsigned long long *key, uint32_t *node, const int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { key[k] = nj_sentinel(); node[k] = (uint32_t)k; }
}

thread_local uint64_t g_nj_launches = 0, g_nj_bytes = 0;

}  // namespace

extern "C" {

int pa_nj_last_stats(uint64_t *launches, uint64_t *bytes) {
    if (launches) *launches = g_nj_launches;
    if (bytes) *bytes = g_nj_bytes;
    return PA_OK;
}

int pa_nj_build(const float *dist, uint32_t n, pa_nj_join *joins, uint32_t *root_left, uint32_t *root_right,
                double *root_right_len, double *kernel_ms) {
    auto fail = [&](int code, const std::string &msg) {
        pa_internal_set_error(msg.c_str());
        return code;
    };
    if (!dist || n < 2 || n > PA_NJ_MAX_TAXA || !root_left || !root_right || !root_right_len || (n > 2 && !joins))
        return fail(PA_EINVAL, "pa_nj_build needs 2..65535 taxa and non-NULL buffers");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PA_ENODEVICE, "no CUDA device available; neighbour joining has no CPU fallback");
    const char *cols = getenv("PAIRALIGN_NJ_COLS");
    g_nj_cols = cols ? atoi(cols) : 0;
    const int ld = (int)((n + 31u) & ~31u);
    const size_t n_tri = (size_t)n * (n - 1) / 2, mat_bytes = (size_t)n * ld * sizeof(float);
    float *M[2] = {nullptr, nullptr}, *S[2] = {nullptr, nullptr}, *tri = nullptr;
    uint32_t *node[2] = {nullptr, nullptr};
    unsigned long long *key = nullptr;
    JoinRec *rec = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaMalloc(&M[0], mat_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&M[1], mat_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&tri, n_tri * sizeof(float));
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaMalloc(&S[k], n * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&node[k], n * sizeof(uint32_t));
    }
    if (e == cudaSuccess) e = cudaMalloc(&key, n * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&rec, n * sizeof(JoinRec));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tri, dist, n_tri * sizeof(float), cudaMemcpyHostToDevice, st);
    uint64_t launches = 0, bytes = 0;
    int cur = 0;
    if (e == cudaSuccess) {
        nj_expand<<<dim3((n + 255) / 256, n), 256, 0, st>>>(tri, M[0], ld, (int)n);
        nj_init<<<(n + 255) / 256, 256, 0, st>>>(key, node[0], (int)n);
        e = cudaEventRecord(e0, st);
        // sums of the matrix as read (src/nj_tree.cpp:39-47, first pass)
        launch_join<true>(st, M[0], nullptr, ld, (int)n, nullptr, S[0], nullptr, nullptr, nullptr, nullptr, 0);
        ++launches; bytes += (uint64_t)n * n * 4;
        uint32_t next_id = n;
        for (int r = (int)n, round = 0; r > 2 && e == cudaSuccess; --r, ++round) {
            const int nb = r - 1 < 148 * 8 ? r - 1 : 148 * 8;
            nj_argmin<<<nb, NJ_THREADS, 0, st>>>(M[cur], ld, r, S[cur], key + round);
            launch_join<false>(st, M[cur], M[cur ^ 1], ld, r, S[cur], S[cur ^ 1], node[cur], node[cur ^ 1], key + round,
                               rec + round, next_id++);
            launches += 2;
            bytes += (uint64_t)r * (r - 1) / 2 * 4 + 2ull * (r - 1) * (r - 1) * 4;
            cur ^= 1;
            if ((round & 1023) == 1023) e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(e1, st);
    std::vector<JoinRec> h_rec(n > 2 ? n - 2 : 0);
    uint32_t h_node[2] = {0, 1};
    float last = 0.0f;
    if (e == cudaSuccess && n > 2)
        e = cudaMemcpyAsync(h_rec.data(), rec, h_rec.size() * sizeof(JoinRec), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_node, node[cur], sizeof h_node, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&last, M[cur] + 1, sizeof(float), cudaMemcpyDeviceToHost, st);   // d(0,1)
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(M[0]); cudaFree(M[1]); cudaFree(tri); cudaFree(S[0]); cudaFree(S[1]); cudaFree(node[0]); cudaFree(node[1]);
    cudaFree(key); cudaFree(rec);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (e != cudaSuccess) return fail(PA_ECUDA, std::string("neighbour joining failed: ") + cudaGetErrorString(e));
    for (size_t k = 0; k < h_rec.size(); ++k) {
        // branch lengths with the reference's expressions (src/nj_tree.cpp:92-94): float arithmetic, stored as double
        const JoinRec &jr = h_rec[k];
        const float length = jr.length;
        const double left_len = (length / 2) + (jr.s_i - jr.s_j) / (2 * (jr.r - 2));
        const double right_len = length - left_len;
        joins[k].left = jr.left; joins[k].right = jr.right;
        joins[k].left_len = left_len; joins[k].right_len = right_len;
    }
    *root_left = h_node[0];
    *root_right = h_node[1];
    *root_right_len = last;
    if (kernel_ms) *kernel_ms = ms;
    g_nj_launches = launches; g_nj_bytes = bytes;
    return PA_OK;
}

}  // extern "C"
