// pa_dp32.cuh -- the int32 A/C/G/T path: pairs too long for the s16x2 kernel and
// explicit pair lists (one pair per warp), and the CTA-per-pair kernel for long
// pairs.  Same recurrence, layout trick and byte tables as pa_dp.cuh; the unit
// of work here is a BLOCK: one pass of a warp over all n rows for 32*K columns.
//
//   warp kernel : one warp walks the blocks of its pair one after the other, the
//                 right edge of a block goes through a scratch row in global
//                 memory (L2 resident) to the next block.
//   CTA kernel  : the NW warps of a CTA work on consecutive blocks of the SAME
//                 pair at the same time, block b+1 following block b by two
//                 chunks of 32 rows.  The edge between neighbouring warps is a
//                 ring of rows in shared memory guarded by two counters
//                 (produced / consumed); only the wrap-around edge (last warp
//                 -> first warp's next block) goes through global memory,
//                 because it has to hold a whole column of n rows.
//
// Two rows per step, ping-ponging two register sets, exactly as the s16x2
// kernel does (pa_dp.cuh: duo_row).
#pragma once

#include "pa_dp.cuh"

namespace pa {

constexpr int CTA_WARPS = 8;          // warps (= concurrent blocks) per long pair
constexpr int RING_ROWS = 256;        // rows per shared-memory edge ring
constexpr int CHUNK_ROWS = 32;        // hand-over granularity between warps
constexpr int XSTAGE_WORDS = 4096;    // 65536 2-bit codes: any accepted sequence fits

// tab[z*4 + x] = (score table lo, hi, increment table lo, hi) for row code x; z = 1: row 0 (+GO)
__device__ __forceinline__ void build_tab32(int4 *tab, const Scoring sc, const uint32_t y0, const int lane) {
    if (lane < 8) {
        const uint32_t xi = lane & 3u, z = lane >> 2;
        const int adj = z ? sc.go : 0;
        const uint32_t Mb = (uint32_t)(sc.match + adj) & 0xffu, Xb = (uint32_t)(sc.mismatch + adj) & 0xffu;
        const uint32_t Mz = (uint32_t)(sc.match + sc.go) & 0xffu, Xz = (uint32_t)(sc.mismatch + sc.go) & 0xffu;
        const uint32_t sh = xi * 8u;
        int4 e;
        e.x = (int)((Xb * 0x01010101u) ^ ((Mb ^ Xb) << sh));    // bytes 0-3: score against base code c
        e.y = (int)((xi == y0) ? Mz : Xz);                      // byte 4: column 0 (+GO), byte 5: 0 = pad
        e.z = (int)(0x01010101u ^ (1u << sh));                  // bytes 0-3: mismatch flag
        e.w = (int)(0x00010000u | (xi != y0 ? 1u : 0u));        // byte 4: column-0 flag, 5: 0, 6: 1
        tab[z * 4 + xi] = e;
    }
}

template <int K>
__device__ __forceinline__ void row32(const int (&Hs)[K], int (&Hd)[K], int (&Gy)[K],
                                      const uint32_t (&Cs)[K], uint32_t (&Cd)[K],
                                      const uint32_t (&selS)[K], const uint32_t (&selI)[K],
                                      const uint32_t Rlo, const uint32_t Rhi, const uint32_t Mlo, const uint32_t Mhi,
                                      const int go, const int ge, int hdiag, int Gl, uint32_t cd, uint32_t cl,
                                      int &Hout, int &Gxout, uint32_t &cout) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = (int)prmt(Rlo, Rhi, selS[k]);
        const uint32_t inc = prmt(Mlo, Mhi, selI[k]);
        const int Gu = Gy[k];
        const uint32_t cu = Cs[k];
        const int h = __vimax3_s32(hdiag, Gu, Gl) + s;
        const int o = hdiag + go;
        const int gy = __viaddmax_s32(Gu, ge, o);
        const int gx = __viaddmax_s32(Gl, ge, o);
        const bool pD = (h >= gy) && (h >= gx);
        const bool pU = (gy >= gx);
        const uint32_t cdi = cd + inc;
        const uint32_t c = pD ? cdi : (pU ? cu : cl);
        hdiag = Hs[k]; cd = cu;
        Hd[k] = h; Gy[k] = gy; Cd[k] = c;
        Gl = gx; cl = c;
    }
    Hout = Hd[K - 1]; Gxout = Gl; cout = cl;
}

// Edge policy of the warp kernel: scratch rows in global memory, no waiting.
struct GlobalEdge {
    const int4 *feed;        // rows of the left edge (or the single virtual row)
    int fmul;                // 0: every row reads feed[0] (block 0), 1: row r reads feed[r]
    int4 *sink;              // where lane 31 puts the right edge (nullptr: last block)
    __device__ __forceinline__ int4 load(int row) const { return __ldcg(&feed[fmul * row]); }
    __device__ __forceinline__ void store(int row, const int4 v) const { __stcg(&sink[row], v); }
    __device__ __forceinline__ bool has_sink() const { return sink != nullptr; }
    __device__ __forceinline__ void acquire(int, int) const {}
    __device__ __forceinline__ void reserve(int, int) const {}
    __device__ __forceinline__ void release(int, int) const {}
    __device__ __forceinline__ void consumed(int, int) const {}
};

// Edge policy of the CTA kernel.  Rows are counted per channel since the start of the
// pair (every block has n rows), so ring slots and the two counters never reset.
template <int RR>
struct RingEdgeT {
    // input side
    const int4 *in_ring;     // shared ring of the channel from the previous warp (nullptr: block 0 or wrap channel)
    const int4 *in_global;   // wrap channel / virtual row
    int in_mul;              // 0 for the virtual row
    volatile int *in_prod;   // rows the producer has published on the input channel (nullptr: nothing to wait for)
    volatile int *in_cons;   // rows this warp has consumed (for the producer's overrun check; nullptr for global)
    int in_base;             // channel row count at the start of this block
    // output side
    int4 *out_ring;
    int4 *out_global;
    volatile int *out_prod;
    volatile int *out_cons;  // nullptr when the output is global (no capacity limit)
    int out_base;

    __device__ __forceinline__ int4 load(int row) const {
        if (in_ring) return in_ring[(in_base + row) & (RR - 1)];
        return __ldcg(&in_global[in_mul * row]);
    }
    __device__ __forceinline__ void store(int row, const int4 v) const {
        if (out_ring) out_ring[(out_base + row) & (RR - 1)] = v;
        else __stcg(&out_global[row], v);
    }
    __device__ __forceinline__ bool has_sink() const { return out_ring != nullptr || out_global != nullptr; }
    // consumer: rows < row_end of this block must be published before they are read
    __device__ __forceinline__ void acquire(int row_end, int lane) const {
        if (in_prod) {
            if (lane == 0) while (*in_prod < in_base + row_end) { }
            __syncwarp();
            __threadfence_block();
        }
    }
    // producer: before writing rows < row_end, the slots they reuse must have been consumed
    __device__ __forceinline__ void reserve(int row_end, int lane) const {
        if (out_cons) {
            if (lane == 0) while (*out_cons < out_base + row_end - RR) { }
            __syncwarp();
        }
    }
    // producer: rows < row_end are written; consumer side: rows < row_end are no longer needed
    __device__ __forceinline__ void release(int row_end, int lane) const {
        __syncwarp();
        if (lane == 31 && out_prod) { __threadfence_block(); *out_prod = out_base + row_end; }
    }
    __device__ __forceinline__ void consumed(int row_end, int lane) const {
        if (lane == 0 && in_cons) *in_cons = in_base + row_end;
    }
};

using RingEdge = RingEdgeT<RING_ROWS>;

struct Best32 {
    int rowBest, rowJ; uint32_t rowC;       // last row so far (columns ascending, strict >)
    int colBest, colI; uint32_t colC;       // last column (rows ascending, strict >); valid in lane 31
};

// One block: columns [j_base, j_base + 32*K) of the pair (j_base may be negative: pad), all n rows.
template <int K, class Edge>
__device__ __forceinline__ void block32(const uint32_t *xs, const int n, const uint32_t *ys, const int j_base,
                                        const bool first_block, const bool last_block, const Scoring sc,
                                        const int4 *tab, const Edge &edge, const int lane, Best32 &best) {
    const int Hinit = -sc.go;
    const int j0 = j_base + lane * K;
    int HX[K], HY[K], Gy[K];
    uint32_t CX[K], CY[K], selS[K], selI[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = j0 + k;
        HX[k] = Hinit; HY[k] = Hinit; Gy[k] = 0; CX[k] = 0; CY[k] = 0;
        if (j < 0)       { selS[k] = 0xddd5u; selI[k] = 0x5555u; }
        else if (j == 0) { selS[k] = 0xccc4u; selI[k] = 0x5654u; }
        else { const uint32_t c = fetch2(ys, j); selS[k] = 0x8880u + 0x1111u * c; selI[k] = 0x5650u | c; }
    }
    int hprev = Hinit;
    uint32_t cprev = 0;
    int HoA = Hinit, GoA = 0, HoB = Hinit, GoB = 0;
    uint32_t coA = 0, coB = 0;
    const int n_steps = ((n + 1) >> 1) + 31;
    const int x_last_word = (n - 1) >> 4;
    constexpr int CHUNK_STEPS = CHUNK_ROWS / 2;
    (void)first_block;

    edge.acquire(min(n, 2 * CHUNK_ROWS), lane);          // two chunks ahead: the one-step prefetch never outruns it
    int4 fA = edge.load(0), fB = edge.load(n > 1 ? 1 : 0);
    uint32_t xw = xs[0];

    for (int t = 0; t < n_steps; ++t) {
        const int iA = 2 * (t - lane);
        if ((t & (CHUNK_STEPS - 1)) == 0 && t > 0) {
            edge.consumed(2 * t, lane);                                   // rows < 2t are in registers or done with
            edge.acquire(min(n, 2 * t + 2 * CHUNK_ROWS), lane);
        }
        // lane 31 starts a new chunk of output rows: the ring slots they reuse must be free
        if (t >= 31 && ((t - 31) & (CHUNK_STEPS - 1)) == 0 && edge.has_sink())
            edge.reserve(min(n, 2 * (t - 31) + CHUNK_ROWS), lane);
        int hinA = __shfl_up_sync(FULL_MASK, HoA, 1), ginA = __shfl_up_sync(FULL_MASK, GoA, 1);
        uint32_t cinA = __shfl_up_sync(FULL_MASK, coA, 1);
        int hinB = __shfl_up_sync(FULL_MASK, HoB, 1), ginB = __shfl_up_sync(FULL_MASK, GoB, 1);
        uint32_t cinB = __shfl_up_sync(FULL_MASK, coB, 1);
        if (lane == 0) {
            hinA = fA.x; ginA = fA.y; cinA = (uint32_t)fA.z;
            hinB = fB.x; ginB = fB.y; cinB = (uint32_t)fB.z;
        }
        fA = edge.load(min(2 * t + 2, n - 1));
        fB = edge.load(min(2 * t + 3, n - 1));
        const uint32_t xi2 = (xw >> ((iA & 15) * 2)) & 15u;
        xw = xs[min(max(iA + 2, 0) >> 4, x_last_word)];
        if (iA >= 0 && iA < n) {
            const bool store = (lane == 31) && edge.has_sink();
            {
                const int4 T = tab[(iA == 0 ? 4 : 0) + (xi2 & 3u)];
                row32<K>(HX, HY, Gy, CX, CY, selS, selI, (uint32_t)T.x, (uint32_t)T.y, (uint32_t)T.z, (uint32_t)T.w,
                         sc.go, sc.ge, hprev, ginA, cprev, cinA, HoA, GoA, coA);
                if (store) edge.store(iA, make_int4(HoA, GoA, (int)coA, 0));
                if (last_block && HoA > best.colBest) { best.colBest = HoA; best.colI = iA; best.colC = coA; }
            }
            if (iA + 1 < n) {
                const int4 T = tab[xi2 >> 2];
                row32<K>(HY, HX, Gy, CY, CX, selS, selI, (uint32_t)T.x, (uint32_t)T.y, (uint32_t)T.z, (uint32_t)T.w,
                         sc.go, sc.ge, hinA, ginB, cinA, cinB, HoB, GoB, coB);
                if (store) edge.store(iA + 1, make_int4(HoB, GoB, (int)coB, 0));
                if (last_block && HoB > best.colBest) { best.colBest = HoB; best.colI = iA + 1; best.colC = coB; }
            }
            hprev = hinB; cprev = cinB;
        }
        // hand-over bookkeeping at chunk granularity (uniform in the warp)
        if (((t - 31) & (CHUNK_STEPS - 1)) == CHUNK_STEPS - 1 && t >= 31) edge.release(min(n, 2 * (t - 31) + 2), lane);
    }
    edge.release(n, lane);
    edge.consumed(n, lane);
    __syncwarp();
    // last row of this block (in Y when n is odd): columns ascending, strict >
    int bv = INT_MIN, bj = INT_MAX;
    uint32_t bc = 0;
    const bool in_y = (n & 1);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = j0 + k;
        const int hk = in_y ? HY[k] : HX[k];
        const uint32_t ck = in_y ? CY[k] : CX[k];
        if (j >= 0 && hk > bv) { bv = hk; bj = j; bc = ck; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int ov = __shfl_xor_sync(FULL_MASK, bv, d), oj = __shfl_xor_sync(FULL_MASK, bj, d);
        const uint32_t oc = __shfl_xor_sync(FULL_MASK, bc, d);
        if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; bc = oc; }
    }
    if (bj != INT_MAX && (bv > best.rowBest || (bv == best.rowBest && bj < best.rowJ))) { best.rowBest = bv; best.rowJ = bj; best.rowC = bc; }
}

__device__ __forceinline__ void finish32(const Best32 &b, const int n, const int m, pa_pair_result *res) {
    pa_pair_result o;
    if (b.rowBest > b.colBest) { o.score = b.rowBest; o.end_i = n - 1; o.end_j = b.rowJ; o.dist = b.rowC & 0xffffu; o.len = b.rowC >> 16; }
    else                       { o.score = b.colBest; o.end_i = b.colI; o.end_j = m - 1; o.dist = b.colC & 0xffffu; o.len = b.colC >> 16; }
    *res = o;
}

// ---------------------------------------------------------------------------
// One pair per warp.  Work items come from `src` (triangle range, explicit
// lists, or the list a previous kernel deferred).  Pairs with a non-A/C/G/T
// sequence go to `deferred` (general kernel); pairs longer than long_len go to
// `deferred_long` (CTA kernel) when that list is given.
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
pa_warp32_kernel(const SeqStore S, const Scoring sc, const PairSource src, const uint64_t count_host,
                 const unsigned int *count_dev, unsigned long long *work_counter, int4 *bbuf_all, const uint32_t bbuf_rows,
                 pa_pair_result *out, uint32_t *deferred, unsigned int *n_deferred,
                 uint32_t *deferred_long, unsigned int *n_deferred_long, const uint32_t long_len) {
    __shared__ __align__(16) uint32_t stage[WARPS_PER_CTA][2][STAGE_WORDS];
    __shared__ int4 tabs[WARPS_PER_CTA][8];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * WARPS_PER_CTA + wib;
    int4 *bbuf = bbuf_all + (size_t)gw * bbuf_rows;
    const uint32_t vrow = bbuf_rows - 1;
    const uint64_t count = count_dev ? (uint64_t)*count_dev : count_host;
    constexpr int W = 32 * K;

    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1ull);
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w >= count) break;
        const uint64_t e = src.idx ? (uint64_t)src.idx[w] : (uint64_t)w;
        uint32_t a, b;
        if (src.ia) { a = src.ia[e]; b = src.ib[e]; }
        else tri_pair(src.first + e, S.n_seq, a, b);
        const int n = (int)S.len[a], m = (int)S.len[b];
        if (n == 0 || m == 0) {   // the reference reads out of bounds here (src/seqpair.cpp:141); defined as "nothing compared"
            if (lane == 0) { pa_pair_result o; o.score = INT_MIN; o.dist = 0; o.len = 0; o.end_i = n - 1; o.end_j = m - 1; out[e] = o; }
            continue;
        }
        if (!(S.pure[a] && S.pure[b])) {
            if (lane == 0) deferred[atomicAdd(n_deferred, 1u)] = (uint32_t)e;
            continue;
        }
        if (deferred_long && (uint32_t)max(n, m) > long_len) {
            if (lane == 0) deferred_long[atomicAdd(n_deferred_long, 1u)] = (uint32_t)e;
            continue;
        }
        __syncwarp();
        const uint32_t *xs = stage_seq(S.p2 + S.off2[a], (uint32_t)(n + 15) >> 4, stage[wib][0], lane);
        const uint32_t *ys = stage_seq(S.p2 + S.off2[b], (uint32_t)(m + 15) >> 4, stage[wib][1], lane);
        build_tab32(tabs[wib], sc, fetch2(S.p2 + S.off2[b], 0), lane);
        if (lane == 0) __stcg(&bbuf[vrow], make_int4(-sc.go, 0, 0, 0));
        __syncwarp();
        const int P = (m + W - 1) / W;
        const int padL = P * W - m;
        Best32 best;
        best.rowBest = INT_MIN; best.rowJ = INT_MAX; best.rowC = 0;
        best.colBest = INT_MIN; best.colI = n - 1; best.colC = 0;
        for (int p = 0; p < P; ++p) {
            GlobalEdge edge;
            edge.feed = p > 0 ? bbuf : bbuf + vrow;
            edge.fmul = p > 0 ? 1 : 0;
            edge.sink = p < P - 1 ? bbuf : nullptr;
            block32<K>(xs, n, ys, p * W - padL, p == 0, p == P - 1, sc, tabs[wib], edge, lane, best);
        }
        best.colBest = __shfl_sync(FULL_MASK, best.colBest, 31);
        best.colI = __shfl_sync(FULL_MASK, best.colI, 31);
        best.colC = __shfl_sync(FULL_MASK, best.colC, 31);
        if (lane == 0) finish32(best, n, m, &out[e]);
    }
}

// ---------------------------------------------------------------------------
// One long pair per CTA: CTA_WARPS warps on consecutive blocks of the pair.
// Work items come from the list `idx` (n_items read from device memory).
// gedge_all: per CTA one column of n rows for the wrap-around edge plus the
// virtual row.
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(CTA_WARPS * 32, 1)
pa_cta32_kernel(const SeqStore S, const Scoring sc, const PairSource src, const unsigned int *n_items,
                unsigned long long *work_counter, int4 *gedge_all, const uint32_t gedge_rows, pa_pair_result *out) {
    __shared__ __align__(16) uint32_t xstage[XSTAGE_WORDS];
    __shared__ __align__(16) int4 rings[CTA_WARPS - 1][RING_ROWS];
    __shared__ int4 tab[8];
    __shared__ int prod[CTA_WARPS], cons[CTA_WARPS];
    __shared__ unsigned long long item_s;
    __shared__ int bestRow[CTA_WARPS], bestJ[CTA_WARPS], bestCol, bestColI;
    __shared__ uint32_t bestRowC[CTA_WARPS], bestColC;
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    int4 *gedge = gedge_all + (size_t)blockIdx.x * gedge_rows;
    const uint32_t vrow = gedge_rows - 1;
    const unsigned long long count = *n_items;
    constexpr int W = 32 * K;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) item_s = atomicAdd(work_counter, 1ull);
        __syncthreads();
        const unsigned long long it = item_s;
        if (it >= count) break;
        const uint64_t e = src.idx ? (uint64_t)src.idx[it] : (uint64_t)it;
        uint32_t a, b;
        if (src.ia) { a = src.ia[e]; b = src.ib[e]; }
        else tri_pair(src.first + e, S.n_seq, a, b);
        const int n = (int)S.len[a], m = (int)S.len[b];
        // stage x for the whole CTA (coalesced 128-bit), tables, counters, virtual row
        {
            const uint4 *g4 = reinterpret_cast<const uint4 *>(S.p2 + S.off2[a]);
            uint4 *s4 = reinterpret_cast<uint4 *>(xstage);
            const uint32_t n4 = (((uint32_t)(n + 15) >> 4) + 3) >> 2;
            for (uint32_t q = threadIdx.x; q < n4; q += blockDim.x) s4[q] = __ldg(&g4[q]);
        }
        const uint32_t *ys = S.p2 + S.off2[b];
        if (w == 0) build_tab32(tab, sc, fetch2(ys, 0), lane);
        if (threadIdx.x < CTA_WARPS) { prod[threadIdx.x] = 0; cons[threadIdx.x] = 0; }
        if (threadIdx.x == 0) { __stcg(&gedge[vrow], make_int4(-sc.go, 0, 0, 0)); bestCol = INT_MIN; bestColI = n - 1; bestColC = 0; }
        __syncthreads();

        const int P = (m + W - 1) / W;
        const int padL = P * W - m;
        Best32 best;
        best.rowBest = INT_MIN; best.rowJ = INT_MAX; best.rowC = 0;
        best.colBest = INT_MIN; best.colI = n - 1; best.colC = 0;
        int round = 0;
        for (int p = w; p < P; p += CTA_WARPS, ++round) {
            RingEdge edge;
            const int prev = (w + CTA_WARPS - 1) % CTA_WARPS;
            // input: block 0 reads the virtual row; warp 0's later blocks read the wrap channel (global);
            // every other block reads the ring its left neighbour fills
            edge.in_ring = nullptr; edge.in_global = gedge + vrow; edge.in_mul = 0; edge.in_prod = nullptr; edge.in_cons = nullptr;
            edge.in_base = 0;
            if (p > 0) {
                if (w == 0) { edge.in_global = gedge; edge.in_mul = 1; edge.in_prod = &prod[prev]; edge.in_base = (round - 1) * n; }
                else { edge.in_ring = rings[prev]; edge.in_prod = &prod[prev]; edge.in_cons = &cons[prev]; edge.in_base = round * n; }
            }
            // output: the last block has none; the last warp writes the wrap channel; others their ring
            edge.out_ring = nullptr; edge.out_global = nullptr; edge.out_prod = nullptr; edge.out_cons = nullptr; edge.out_base = round * n;
            if (p < P - 1) {
                edge.out_prod = &prod[w];
                if (w == CTA_WARPS - 1) edge.out_global = gedge;
                else { edge.out_ring = rings[w]; edge.out_cons = &cons[w]; }
            }
            block32<K>(xstage, n, ys, p * W - padL, p == 0, p == P - 1, sc, tab, edge, lane, best);
        }
        // combine: last row (lowest column wins ties), last column (from the warp that ran the last block)
        if (lane == 0) { bestRow[w] = best.rowBest; bestJ[w] = best.rowJ; bestRowC[w] = best.rowC; }
        if (lane == 31 && ((P - 1) % CTA_WARPS) == w) { bestCol = best.colBest; bestColI = best.colI; bestColC = best.colC; }
        __syncthreads();
        if (threadIdx.x == 0) {
            Best32 f;
            f.rowBest = INT_MIN; f.rowJ = INT_MAX; f.rowC = 0;
            for (int q = 0; q < CTA_WARPS; ++q)
                if (bestJ[q] != INT_MAX && (bestRow[q] > f.rowBest || (bestRow[q] == f.rowBest && bestJ[q] < f.rowJ))) {
                    f.rowBest = bestRow[q]; f.rowJ = bestJ[q]; f.rowC = bestRowC[q];
                }
            f.colBest = bestCol; f.colI = bestColI; f.colC = bestColC;
            finish32(f, n, m, &out[e]);
        }
    }
}

// ---------------------------------------------------------------------------
// pairalign -a.  The move-storing DP kernels (pa_dp_moves.cuh for A/C/G/T pairs, pa_general_dirs_kernel for the rest)
// keep 2 bits per column SLOT (slot = column + pad, the right-aligned layout of the DP), row-major, row stride =
// P*32*K/4 bytes, at dirs + dirs_off[e]: bit 0 = the move is not diagonal, bit 1 = left rather than up (only
// meaningful with bit 0 set).  The kernel below then walks each pair back.
// ---------------------------------------------------------------------------
// One warp per pair: the reference's walk (src/seqpair.cpp:146-178).  ops receives one byte per aligned
// column in the REVERSE order the reference builds them in (it reverses at :183-188; the host does that):
// 0 = x[i] over y[j], 1 = x[i] over a gap, 2 = gap over y[j].  kcols: strip width of the kernel that wrote the
// moves of this pair (16 for A/C/G/T pairs, 8 for the general kernel).  The walk also counts what
// hamming_distance(false) / similarity(false) count over the aligned strings (src/seqpair.cpp:238-274): columns
// with a base on both sides, and those of them whose sets do not intersect -- res[e].len / res[e].dist.
//
// The walk is a chain of dependent loads, one row of the move store (a different cache line) per step, and the
// store of a long pair (226 MB for 30 kb x 30 kb) is in DRAM by the time the walk starts.  So the warp works in
// epochs of WALK_EPOCH rows: all lanes fetch, for every row of the NEXT epoch, the 128 slots around the diagonal
// through the current cell (two 16-byte loads per row, in flight while lane 0 walks the current epoch out of
// shared memory) and park them in the other half of a double buffer.  A step whose slot lies outside its row's
// window (the path drifted more than ~32 columns off the diagonal within two epochs) reads global memory instead:
// the windows are a cache, the walk is exact either way.
constexpr int WALK_EPOCH = 64;
constexpr int WALK_WARPS = 4;
constexpr int WALK_WIN = 128;        // slots per row window (32 bytes)

__global__ void __launch_bounds__(WALK_WARPS * 32)
pa_walk_kernel(const SeqStore S, const uint32_t *ia, const uint32_t *ib, const uint64_t count,
               pa_pair_result *res, const uint8_t *dirs, const unsigned long long *dirs_off,
               uint8_t *ops, const unsigned long long *ops_off, uint32_t *n_ops,
               const int k_pure, const int k_amb, const int k_general) {
    __shared__ __align__(16) uint4 win[WALK_WARPS][2][WALK_EPOCH][2];
    __shared__ int wbase[WALK_WARPS][2][WALK_EPOCH];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint64_t e = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= count) return;
    const uint32_t a = ia[e], b = ib[e];
    const int n = (int)S.len[a], m = (int)S.len[b];
    // ops == nullptr: statistics only (the caller wants res[e].len / res[e].dist, not the op string)
    const bool emit = ops != nullptr;
    uint8_t *o = emit ? ops + ops_off[e] : nullptr;
    uint32_t k = 0;
    if (n == 0 || m == 0) {       // nothing to align (undefined in the reference): the other sequence against gaps
        if (lane == 0) {
            if (emit) {
                for (int q = 0; q < n; ++q) o[k++] = 1;
                for (int q = 0; q < m; ++q) o[k++] = 2;
                n_ops[e] = k;
            }
            res[e].dist = 0; res[e].len = 0;
        }
        return;
    }
    const uint32_t *x4 = S.p4 + S.off4[a], *y4 = S.p4 + S.off4[b];
    uint32_t n_cols = 0, n_diff = 0;
    // strip width of the kernel that stored this pair's moves: plain A/C/G/T, IUPAC codes without gap characters, the rest
    const int W = 32 * ((S.pure[a] && S.pure[b]) ? k_pure : (S.fastok[a] && S.fastok[b]) ? k_amb : k_general);
    const int P = (m + W - 1) / W;
    const int padL = P * W - m;
    const int slots = P * W;                 // >= 256
    const size_t stride = (size_t)slots / 4;
    const uint8_t *d = dirs + dirs_off[e];
    int i = res[e].end_i, j = res[e].end_j;
    // windows of rows i_top, i_top-1, ... (two per lane) along the diagonal that passes (i_top, slot_top)
    uint4 w0[2], w1[2];
    int wb[2];
    auto fetch = [&](const int i_top, const int slot_top) {
#pragma unroll
        for (int h = 0; h < WALK_EPOCH / 32; ++h) {
            const int q = lane + 32 * h;
            const int r = i_top - q;
            int base = (slot_top - q - WALK_WIN / 2) & ~63;
            base = min(max(base, 0), slots - WALK_WIN);
            wb[h] = base;
            if (r >= 0) {
                const uint4 *src = reinterpret_cast<const uint4 *>(d + (size_t)r * stride + (base >> 2));
                w0[h] = __ldcg(src); w1[h] = __ldcg(src + 1);
            } else { w0[h] = make_uint4(0, 0, 0, 0); w1[h] = w0[h]; }
        }
    };
    auto park = [&](const int buf) {
#pragma unroll
        for (int h = 0; h < WALK_EPOCH / 32; ++h) {
            const int q = lane + 32 * h;
            win[wib][buf][q][0] = w0[h]; win[wib][buf][q][1] = w1[h];
            wbase[wib][buf][q] = wb[h];
        }
    };
    fetch(i, j + padL);
    park(0);
    __syncwarp();
    if (lane == 0) {
        if (emit) {
            if (i < n - 1) { for (int pos = n - 1; pos > i; --pos) o[k++] = 1; }
            else if (j < m - 1) { for (int pos = m - 1; pos > j; --pos) o[k++] = 2; }
        }
    }
    for (int buf = 0;; buf ^= 1) {
        i = __shfl_sync(FULL_MASK, i, 0);
        j = __shfl_sync(FULL_MASK, j, 0);
        if (i < 0 && j < 0) break;
        const int i_top = i;
        fetch(i_top - WALK_EPOCH, j + padL - WALK_EPOCH);        // the next epoch's rows, in flight during this walk
        if (lane == 0) {
            const int i_stop = i_top - WALK_EPOCH;                // walk until WALK_EPOCH rows are behind us (or the end)
            const uint8_t *wbytes = reinterpret_cast<const uint8_t *>(&win[wib][buf][0][0]);
            while ((i >= 0 || j >= 0) && i > i_stop) {
                uint32_t mv = 3;
                if (i >= 0 && j >= 0) {
                    const int slot = j + padL, q = i_top - i;
                    const int rel = slot - wbase[wib][buf][q];
                    const uint32_t byte = ((unsigned)rel < (unsigned)WALK_WIN) ? wbytes[q * 32 + (rel >> 2)]
                                                                              : __ldcg(&d[(size_t)i * stride + (slot >> 2)]);
                    mv = (byte >> ((slot & 3) * 2)) & 3u;
                }
                if ((mv & 1u) == 0) {
                    const uint32_t sx = fetch4(x4, i), sy = fetch4(y4, j);       // not on the chain of dependent loads
                    n_cols += (sx != 0 && sy != 0) ? 1u : 0u;
                    n_diff += (sx != 0 && sy != 0 && (sx & sy) == 0) ? 1u : 0u;
                    if (emit) o[k++] = 0;
                    --i; --j;
                }
                else if (j < 0 || (i >= 0 && mv == 1)) { if (emit) o[k++] = 1; --i; }
                else { if (emit) o[k++] = 2; --j; }
            }
        }
        __syncwarp();
        park(buf ^ 1);
        __syncwarp();
    }
    if (lane == 0) { if (emit) n_ops[e] = k; res[e].len = n_cols; res[e].dist = n_diff; }
}

}  // namespace pa
