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
//
// Layout.  One dense symmetric float matrix in HBM, never rewritten: K slots are kept free in FRONT of the taxa
// and every new node takes the free slot next to the occupied ones, so physical order == the reference's order.
// The two joined taxa die in place: their rows and columns are zeroed (adding +0 never changes a float sum that
// started at +0) and their S becomes -inf (which makes every Q value with them +inf or NaN, never below 100000).
// When the free slots run out (K ~ 0.67 sqrt(r) joins: balances the dead slots carried along against the copy)
// the live rows and columns are compacted into the second buffer.  All accesses are aligned 128-bit.
//
// Three launches per join, no host round trip (everything the host needs is recorded on the device):
//   nj_argmin  all live pairs in parallel; key = (order-preserving float bits, row-major rank) reduced with a
//              64-bit atomicMin: the smallest value wins, ties go to the first pair in the reference's visiting order.
//   nj_sums    one warp per block of W columns: the rows of its columns arrive through a ring of shared-memory tiles
//              filled by bulk copies (cp.async.bulk, completion on mbarriers, 7 tiles in flight) and are added up row
//              by row -- the adds of one column are done by ONE thread in row order (the reference's summation order;
//              float addition is not associative, a tree reduction would change the bits); a lane runs one column.
//   nj_update  the join: new node's row and column, zeroes for the two joined taxa, the record for the host.
// __fmul_rn/__fsub_rn/__fadd_rn/__fdiv_rn keep the compiler from contracting into FMAs: the bits must be the ones
// the reference's scalar float code produces.  HBM-bound: per join 4 B x (r^2/2 [argmin] + r^2 [sums]).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pairalign_b200.h"

void pa_internal_set_error(const char *msg);   // pa_capi.cu: message behind pa_last_error()

namespace {

constexpr int NJ_THREADS = 256;               // argmin
constexpr int NJ_TILE = 2048;                 // floats per tile: (2048 / W) rows x W columns
constexpr float NJ_START = 100000.0f;         // "float M=100000" (src/nj_tree.cpp:53)
constexpr int NJ_ROW_PAD = 256;               // allocated rows are a multiple of the tallest tile

struct JoinRec { uint32_t left, right; float length, s_i, s_j; int r, pi, pj; };

// Programmatic dependent launch: the three kernels of a join are launched with programmaticStreamSerialization, let their
// successor's CTAs become resident at once (launch_dependents) and only then wait for their predecessor's memory
// (griddepcontrol.wait).  The launch latency of kernel k+1 then overlaps kernel k instead of following it -- below a
// few thousand taxa the joins are launch-bound, not bandwidth-bound.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

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

__device__ __forceinline__ float neg_inf() { return __uint_as_float(0xff800000u); }

// upper triangle (row-major) -> dense symmetric matrix with a zero diagonal, placed at slot `first`
__global__ void nj_expand(const float *tri, float *M, const int ld, const int n, const int first) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (b >= n) return;
    float v = 0.0f;
    if (a != b) {
        const size_t lo = a < b ? a : b, hi = a < b ? b : a;
        v = tri[lo * (size_t)n - lo * (lo + 1) / 2 + (hi - lo - 1)];
    }
    M[(size_t)(first + a) * ld + first + b] = v;
}

// S = -inf, nobody alive, no node
__global__ void nj_fill(float *S0, float *S1, uint32_t *node, uint8_t *alive, const int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) { S0[k] = neg_inf(); S1[k] = neg_inf(); node[k] = 0; alive[k] = 0; }
}

__global__ void nj_init(unsigned long long *key, const int n_keys, uint32_t *node, uint8_t *alive, const int first, const int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_keys) key[k] = nj_sentinel();
    if (k < n) { node[first + k] = (uint32_t)k; alive[first + k] = 1; }
}

// First strict minimum of (r-2)*d(p,q) - S[p] - S[q] over live p < q in row-major order, below 100000.
// Slots first..P-1; dead slots have S = -inf (their values come out +inf or NaN) and whole dead rows are skipped.
// A thread meets its pairs in ascending rank, so "strictly smaller replaces" keeps the first of equal values, like
// the reference's scan; across threads the 64-bit key (order-preserving value bits, rank) does the same.
// The last block to finish turns "nothing below 100000" into the pair the reference joins then: its initial
// i = 0, j = 0, i.e. the first two taxa in order.
__global__ void __launch_bounds__(NJ_THREADS) nj_argmin(const float *__restrict__ M, const int ld, const int first, const int P,
                                                        const int r, const float *__restrict__ S, const uint8_t *__restrict__ alive,
                                                        unsigned long long *key, unsigned int *ticket) {
    pdl_enter();
    const float fr2 = (float)(r - 2);
    float bv = NJ_START;
    unsigned int brank = 0xffffffffu;
    const float4 *S4 = reinterpret_cast<const float4 *>(S);
    const int full4 = P >> 2;                   // chunks [.., full4) lie entirely below P
    for (int p = first + blockIdx.x; p < P - 1; p += gridDim.x) {
        const float sp = S[p];
        if (sp == neg_inf()) continue;
        const float4 *row = reinterpret_cast<const float4 *>(M + (size_t)p * ld);
        const unsigned int base = (unsigned int)p << 16;
        const int s4 = (p + 1) >> 2;
        for (int q4 = s4 + threadIdx.x; q4 * 4 < P; q4 += NJ_THREADS) {
            const float4 d = row[q4], sq = S4[q4];
            const float v0 = __fsub_rn(__fsub_rn(__fmul_rn(fr2, d.x), sp), sq.x);
            const float v1 = __fsub_rn(__fsub_rn(__fmul_rn(fr2, d.y), sp), sq.y);
            const float v2 = __fsub_rn(__fsub_rn(__fmul_rn(fr2, d.z), sp), sq.z);
            const float v3 = __fsub_rn(__fsub_rn(__fmul_rn(fr2, d.w), sp), sq.w);
            const unsigned int rk = base + (unsigned int)(q4 * 4);
            if (q4 > s4 && q4 < full4) {        // NaN never wins, like 'value < M'
                if (v0 < bv) { bv = v0; brank = rk; }
                if (v1 < bv) { bv = v1; brank = rk + 1; }
                if (v2 < bv) { bv = v2; brank = rk + 2; }
                if (v3 < bv) { bv = v3; brank = rk + 3; }
            } else {
                const int q = q4 * 4;
                if (q > p && q < P && v0 < bv) { bv = v0; brank = rk; }
                if (q + 1 > p && q + 1 < P && v1 < bv) { bv = v1; brank = rk + 1; }
                if (q + 2 > p && q + 2 < P && v2 < bv) { bv = v2; brank = rk + 2; }
                if (q + 3 > p && q + 3 < P && v3 < bv) { bv = v3; brank = rk + 3; }
            }
        }
    }
    unsigned long long best = brank == 0xffffffffu ? nj_sentinel() : (((unsigned long long)ordered_bits(bv) << 32) | brank);
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
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0;
            if (*(volatile unsigned long long *)key == nj_sentinel()) {
                int pair[2] = {first, first + 1}, found = 0;
                for (int p = first; p < P && found < 2; ++p) if (alive[p]) pair[found++] = p;
                *key = ((unsigned long long)ordered_bits(NJ_START) << 32) | ((unsigned int)pair[0] << 16) | (unsigned int)pair[1];
            }
        }
    }
}

// ---- bulk-copy ring --------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one 2-D box (W columns x TR rows) of the matrix -> shared memory; completes on the mbarrier
__device__ __forceinline__ void tma_tile_g2s(void *dst, const CUtensorMap *map, int col, int row, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned int parity) {
    unsigned int ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

constexpr int NJ_W = 32;                       // columns per warp: one column per lane
constexpr int NJ_TR = NJ_TILE / NJ_W;          // rows per tile (64)
constexpr int NJ_STAGES = 6;                   // 8 KB tiles: 5 of them (40 KB) in flight per warp, 4 warps per SM;
                                               // 3, 12 and 24 stages measured the same (the add chain paces the warp)

// The join itself (src/nj_tree.cpp:79-176), one thread per slot: the new node's row and column
// (d_ki + d_kj - d_ij) / 2 in the free slot in front, zeroes in the rows and columns of the two joined taxa, and
// the record the host turns into branch lengths.  Afterwards the matrix is simply the next round's matrix.
__global__ void __launch_bounds__(256) nj_update(float *M, const int ld, const int first, const int P, const int r,
                                                 const float *__restrict__ S, uint32_t *node, uint8_t *alive,
                                                 const unsigned long long *key, JoinRec *rec, const uint32_t new_id) {
    pdl_enter();
    const unsigned int rank = (unsigned int)*key;
    const int i = (int)(rank >> 16), jp = (int)(rank & 0xffffu), nw = first - 1;
    const size_t ldz = (size_t)ld;
    const float length = M[(size_t)i * ldz + jp];          // nobody writes (i, jp): the threads of i and jp skip it
    const int k = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (k < P && k != i && k != jp) {
        float x = 0.0f;
        if (alive[k]) x = __fdiv_rn(__fsub_rn(__fadd_rn(M[(size_t)i * ldz + k], M[(size_t)jp * ldz + k]), length), 2.0f);
        M[(size_t)nw * ldz + k] = x; M[(size_t)k * ldz + nw] = x;
        M[(size_t)i * ldz + k] = 0.0f; M[(size_t)k * ldz + i] = 0.0f;
        M[(size_t)jp * ldz + k] = 0.0f; M[(size_t)k * ldz + jp] = 0.0f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        JoinRec jr;
        jr.left = node[i]; jr.right = node[jp]; jr.length = length; jr.s_i = S[i]; jr.s_j = S[jp];
        jr.r = r; jr.pi = i; jr.pj = jp;
        *rec = jr;
        node[nw] = new_id;
        alive[i] = 0; alive[jp] = 0; alive[nw] = 1;        // the threads above test i and jp explicitly; nw < first
    }
}

struct SumArgs {
    const float *M;
    int ld, P, lo;                  // slots lo..P-1 are in use
    float *S2;                      // the sums
    const uint8_t *alive;
};

// S[b] = ((d(lo,b) + d(lo+1,b)) + ...) down column b in row order: dead rows are zero, dead columns get -inf.
// One WARP per CTA owns 32 columns, one per lane.  Tiles of 64 rows x 32 columns arrive as TMA boxes (one request
// per 8 KB tile, issued by lane 0) that complete on the stage's mbarrier: no CTA-wide barrier, no per-element copy
// instructions.  Per row a warp issues one shared load and one FADD; the loads are those of the NEXT tile (into
// registers), so the only latency on the critical path is the dependent FADD: 4 cycles per row.
__global__ void __launch_bounds__(32) nj_sums(const SumArgs g, const __grid_constant__ CUtensorMap tmap) {
    constexpr int W = NJ_W, TR = NJ_TR;
    extern __shared__ __align__(128) float tile[];         // [NJ_STAGES][TR][W]
    __shared__ __align__(8) uint64_t full[NJ_STAGES];
    pdl_enter();
    const int P = g.P, lane = threadIdx.x;
    const int b0 = (g.lo / W) * W + blockIdx.x * W;
    const int a_start = (g.lo / TR) * TR;
    const int n_tiles = (P - a_start + TR - 1) / TR;
    const int b = b0 + lane;
    const bool dead_b = b >= P || !g.alive[b];
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < NJ_STAGES; ++st) mbar_init(&full[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int t) {
        if (t < n_tiles && lane == 0) {
            const int st = t % NJ_STAGES;
            mbar_expect_tx(&full[st], NJ_TILE * 4);
            tma_tile_g2s(tile + st * NJ_TILE, &tmap, b0, a_start + t * TR, &full[st]);
        }
    };
    for (int t = 0; t < NJ_STAGES - 1; ++t) issue(t);
    // The adds of a column are one dependent chain (4 cycles each); a tile's 64 values are therefore fetched into
    // registers while the PREVIOUS tile is being added, so that no shared-memory latency sits between two adds
    // (16 loads, then 16 adds, per trip of the old loop: 7.2 cycles per row; now the chain alone).
    float v[TR], w[TR];
    mbar_wait(&full[0], 0u);
    {
        const float *col = tile + lane;
#pragma unroll
        for (int a = 0; a < TR; ++a) v[a] = col[a * W];
    }
    float s = 0.0f;
    for (int t = 0; t < n_tiles; ++t) {
        __syncwarp();                           // every lane has tile t in registers: its slot's predecessor may be refilled
        issue(t + NJ_STAGES - 1);               // into the slot of tile t - 1
        if (t + 1 < n_tiles) {
            const int st = (t + 1) % NJ_STAGES;
            mbar_wait(&full[st], (unsigned int)((t + 1) / NJ_STAGES) & 1u);
            const float *col = tile + st * NJ_TILE + lane;
#pragma unroll
            for (int a = 0; a < TR; ++a) { s = __fadd_rn(s, v[a]); w[a] = col[a * W]; }
#pragma unroll
            for (int a = 0; a < TR; ++a) v[a] = w[a];
        } else {
#pragma unroll
            for (int a = 0; a < TR; ++a) s = __fadd_rn(s, v[a]);
        }
    }
    g.S2[b] = dead_b ? neg_inf() : s;           // columns past P are dead: -inf
}

constexpr int sums_smem() { return NJ_STAGES * NJ_TILE * (int)sizeof(float); }

cudaError_t configure_sums() {                  // > 48 KB of dynamic shared memory needs the opt-in (per device)
    return cudaFuncSetAttribute(nj_sums, cudaFuncAttributeMaxDynamicSharedMemorySize, sums_smem());
}

// tensor map of one matrix buffer: boxes of 32 columns x 64 rows
bool make_tile_map(float *base, int rows, int ld, CUtensorMap *out) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
    auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t estr[2] = {1, 1}, box[2] = {NJ_W, NJ_TR};
    return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void launch_sums(cudaStream_t st, const SumArgs &a, const CUtensorMap &map) {
    const int span = a.P - (a.lo / NJ_W) * NJ_W;
    const int blocks = (span + NJ_W - 1) / NJ_W;
    launch_pdl(nj_sums, dim3(blocks), dim3(32), sums_smem(), st, a, map);
}

// ---- compaction: live slots of [first, P) -> slots [k2, k2 + r) of the other buffer, order kept ------------------
__global__ void __launch_bounds__(1024) nj_scan(const uint8_t *alive, const int first, const int P, int *oldidx) {
    __shared__ int warp_tot[32];
    __shared__ int running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int base = first; base < P; base += 1024) {
        const int p = base + (int)threadIdx.x;
        const bool live = p < P && alive[p];
        const unsigned int m = __ballot_sync(0xffffffffu, live);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_tot[w] = __popc(m);
        __syncthreads();
        int off = running;
        for (int k = 0; k < w; ++k) off += warp_tot[k];
        if (live) oldidx[off + __popc(m & ((1u << lane) - 1u))] = p;
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int k = 0; k < 32; ++k) tot += warp_tot[k]; running += tot; }
        __syncthreads();
    }
}

__global__ void nj_compact(const float *__restrict__ M, float *__restrict__ M2, const int ld, const int *__restrict__ oldidx,
                           const int r, const int k2, const float *__restrict__ S, float *__restrict__ S2,
                           const uint32_t *__restrict__ node, uint32_t *__restrict__ node2, uint8_t *__restrict__ alive2) {
    const int a = blockIdx.y, b = blockIdx.x * blockDim.x + threadIdx.x;
    const int oa = oldidx[a];
    if (b < r) M2[(size_t)(k2 + a) * ld + k2 + b] = M[(size_t)oa * ld + oldidx[b]];
    if (b == 0) { S2[k2 + a] = S[oa]; node2[k2 + a] = node[oa]; alive2[k2 + a] = 1; }
}

// joins before the free slots in front run out (see the header comment)
int epoch_len(int r) {
    const int k = (int)lround(0.67 * sqrt((double)r));
    return k < 4 ? 4 : k;
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
        return fail(PA_EINVAL, "pa_nj_build needs 2..65000 taxa and non-NULL buffers");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(PA_ENODEVICE, "no CUDA device available; neighbour joining has no CPU fallback");
    const int k0 = epoch_len((int)n);
    const int p_max = (int)n + k0;                                        // < 65536: ranks are p << 16 | q
    const int ld = (p_max + 127) & ~127;                                  // the widest column block is 128
    const int rows = (p_max + NJ_ROW_PAD - 1) / NJ_ROW_PAD * NJ_ROW_PAD;  // tiles read (zero) rows past the end
    const size_t n_tri = (size_t)n * (n - 1) / 2, mat_bytes = (size_t)rows * ld * sizeof(float);
    float *M[2] = {nullptr, nullptr}, *S[2] = {nullptr, nullptr}, *tri = nullptr;
    uint32_t *node[2] = {nullptr, nullptr};
    uint8_t *alive[2] = {nullptr, nullptr};
    int *oldidx = nullptr;
    unsigned int *ticket = nullptr;
    unsigned long long *key = nullptr;
    JoinRec *rec = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = configure_sums();
    if (e == cudaSuccess) e = cudaMalloc(&M[0], mat_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&M[1], mat_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&tri, n_tri * sizeof(float));
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaMalloc(&S[k], ld * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc(&node[k], ld * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMalloc(&alive[k], ld);
    }
    if (e == cudaSuccess) e = cudaMalloc(&oldidx, ld * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&ticket, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&key, n * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&rec, n * sizeof(JoinRec));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tri, dist, n_tri * sizeof(float), cudaMemcpyHostToDevice, st);
    CUtensorMap maps[2];
    if (e == cudaSuccess && !(make_tile_map(M[0], rows, ld, &maps[0]) && make_tile_map(M[1], rows, ld, &maps[1]))) {
        cudaFree(M[0]); cudaFree(M[1]); cudaFree(tri);
        return fail(PA_ECUDA, "neighbour joining failed: cuTensorMapEncodeTiled is not available");
    }
    uint64_t launches = 0, bytes = 0;
    int mcur = 0, scur = 0, ncur = 0;            // matrix buffer, sums buffer, node/alive buffer in use
    int first = k0, P = k0 + (int)n, r = (int)n;
    const int fill_blocks = (ld + 255) / 256;
    // live slots -> the other buffers, k2 free slots in front
    auto compact = [&](int k2) {
        cudaMemsetAsync(M[mcur ^ 1], 0, mat_bytes, st);
        nj_fill<<<fill_blocks, 256, 0, st>>>(S[scur ^ 1], S[scur ^ 1], node[ncur ^ 1], alive[ncur ^ 1], ld);
        nj_scan<<<1, 1024, 0, st>>>(alive[ncur], first, P, oldidx);
        nj_compact<<<dim3((r + 255) / 256, r), 256, 0, st>>>(M[mcur], M[mcur ^ 1], ld, oldidx, r, k2, S[scur], S[scur ^ 1],
                                                             node[ncur], node[ncur ^ 1], alive[ncur ^ 1]);
        launches += 3;
        mcur ^= 1; scur ^= 1; ncur ^= 1;
        first = k2; P = k2 + r;
    };
    if (e == cudaSuccess) {
        cudaMemsetAsync(M[0], 0, mat_bytes, st);
        cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st);
        nj_fill<<<fill_blocks, 256, 0, st>>>(S[0], S[1], node[0], alive[0], ld);
        nj_expand<<<dim3((n + 255) / 256, n), 256, 0, st>>>(tri, M[0], ld, (int)n, first);
        nj_init<<<(n + 255) / 256, 256, 0, st>>>(key, (int)n, node[0], alive[0], first, (int)n);
        e = cudaEventRecord(e0, st);
        // sums of the matrix as read (src/nj_tree.cpp:39-47, first pass)
        SumArgs a{};
        a.M = M[mcur]; a.ld = ld; a.P = P; a.lo = first; a.S2 = S[scur]; a.alive = alive[ncur];
        launch_sums(st, a, maps[mcur]);
        ++launches; bytes += (uint64_t)n * n * 4;
        uint32_t next_id = n;
        for (int round = 0; r > 2 && e == cudaSuccess; ++round) {
            if (first == 0) compact(epoch_len(r));
            const int span = P - first;
            launch_pdl(nj_argmin, dim3(span - 1 < 148 * 8 ? span - 1 : 148 * 8), dim3(NJ_THREADS), 0, st, (const float *)M[mcur], ld, first, P, r,
                       (const float *)S[scur], (const uint8_t *)alive[ncur], key + round, ticket);
            launch_pdl(nj_update, dim3((span + 255) / 256), dim3(256), 0, st, M[mcur], ld, first, P, r, (const float *)S[scur], node[ncur],
                       alive[ncur], (const unsigned long long *)(key + round), rec + round, next_id++);
            --first; --r;
            a.M = M[mcur]; a.P = P; a.lo = first; a.S2 = S[scur ^ 1]; a.alive = alive[ncur];
            launch_sums(st, a, maps[mcur]);
            scur ^= 1;
            launches += 3;
            bytes += (uint64_t)(r + 1) * r / 2 * 4 + (uint64_t)r * r * 4;
            if ((round & 1023) == 1023) e = cudaGetLastError();
        }
        if (e == cudaSuccess && n > 2) compact(0);   // the two that are left: slots 0 and 1
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(e1, st);
    std::vector<JoinRec> h_rec(n > 2 ? n - 2 : 0);
    uint32_t h_node[2] = {0, 1};
    float last = 0.0f;
    if (e == cudaSuccess && n > 2)
        e = cudaMemcpyAsync(h_rec.data(), rec, h_rec.size() * sizeof(JoinRec), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_node, node[ncur] + first, sizeof h_node, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)   // d(0,1)
        e = cudaMemcpyAsync(&last, M[mcur] + (size_t)first * ld + first + 1, sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(M[0]); cudaFree(M[1]); cudaFree(tri); cudaFree(S[0]); cudaFree(S[1]); cudaFree(node[0]); cudaFree(node[1]);
    cudaFree(alive[0]); cudaFree(alive[1]); cudaFree(oldidx); cudaFree(ticket); cudaFree(key); cudaFree(rec);
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
