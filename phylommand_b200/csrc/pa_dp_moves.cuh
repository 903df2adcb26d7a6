// pa_dp_moves.cuh -- pairalign -a for A/C/G/T pairs: the s16x2 two-pairs-per-warp recurrence of pa_dp.cuh with the
// MOVE of every cell stored (2 bits per column slot) instead of counters carried, for the walk of
// src/seqpair.cpp:146-188 (pa_walk_kernel, pa_dp32.cuh).
//
// What a cell costs here: the score PRMT, VIMNMX3, VIADD.16x2, two VIADDMNMX and the two predicate-producing
// VIMNMX.S16x2 on the ALU pipe (7), H + GO and FOUR PREDICATED IMADs on the FMA pipe: bit 0 of a cell's move is
// "not diagonal", bit 1 "left rather than up", and `@!p IMAD mv, mv, one, imm` adds the bit where the predicate says
// so (one = a register holding 1, imm = the bit at this column's position).
// No counters, no increment tables, no selects, and half the hand-over per row (H and Gx only).  The statistics
// (compared / differing columns) are counted by the walk.
//
// Work item: two entries (x, y1), (x, y2) of the caller's pair list that share their first sequence (the order
// pairalign visits the triangle in), or one entry alone.  Strip width 16: one 32-bit word of moves per lane, row and
// pair, so a warp's row is one 128-byte line per pair, in exactly the layout the int32 kernels used (row-major,
// P * 512 slots per row, right-aligned) -- each pair in its OWN geometry: pair 2 of an item may need fewer blocks
// than pair 1, its words then start at the first block that holds one of its columns.
//
// Stores: the wavefront is skewed (lane l is on row 2(t - l) at step t), so storing a lane's word when it is made
// would scatter every warp store over 32 rows, 4 bytes each -- 32 partial-sector writes per instruction, which L2
// answers with a fill read and a second write-back (measured: 1.7 x the move bytes written, 0.7 x read, the kernel
// waiting on memory at 37 % issue utilisation).  Each lane therefore delays its words in a private 8-deep ring in
// shared memory by 7 - (lane mod 8) steps: the eight lanes that share a 32-byte sector of a row then store their
// words of that row in the same instruction -- four full sectors per warp store, streaming (st.global.cs) so that the
// moves do not push the edge rows out of L2.  No exchange between lanes, hence no synchronisation.
//
// Scores beyond int16 (pairs longer than max_len16) use the floating window of align_warp_duo: per-lane 32-bit
// offsets, re-based by WIN_Q when the lane's right edge leaves +-WIN_T; the edge rows carry their offsets.
//
//   pa_warp_duo_moves_kernel   one item per warp, blocks one after the other, edge through L2 scratch rows
//   pa_cta_duo_moves_kernel    one item per CTA: NW warps on NW consecutive blocks, edges through shared-memory
//                              rings (RingEdge of pa_dp32.cuh), for pairs so long that a batch (sized by the memory
//                              their moves need: 226 MB for 30 kb x 30 kb) cannot fill a warp-per-item grid
#pragma once

#include "pa_dp32.cuh"
#include "pa_dp_sets.cuh"

namespace pa {

constexpr int KMOV = 16;                 // columns per lane: 32 bits of moves
constexpr int MOVES_CTA_WARPS = 12;      // warps per item in the CTA form: 3 per scheduler, 170 registers each (no spills)
#ifndef MOVES_WARP_MINB
#define MOVES_WARP_MINB 3
#endif
constexpr int MOVES_DELAY = 8;            // depth of a lane's delay ring (steps): 32-byte sector / 4-byte word
constexpr int MOVES_RING_ROWS = 128;     // rows per shared-memory edge ring (11 rings + the staged x stay below 48 KB)
using MovesRing = RingEdgeT<MOVES_RING_ROWS>;
constexpr size_t moves_cta_smem(const bool sets) {
    return (size_t)(sets ? 2 : 1) * XSTAGE_WORDS * 4 + (size_t)(MOVES_CTA_WARPS - 1) * MOVES_RING_ROWS * 16 +
           (size_t)MOVES_CTA_WARPS * 2 * MOVES_DELAY * 32 * 8;
}

template <int K, int GEC = 0>
__device__ __forceinline__ void duo_moves_row(const uint32_t (&Hs)[K], uint32_t (&Hd)[K], uint32_t (&Gy)[K],
                                              const uint32_t (&selS)[K], const uint32_t Rlo, const uint32_t Rhi,
                                              const uint32_t GOc, const uint32_t GEpk, const uint32_t one,
                                              uint32_t hdiag, uint32_t Gl, uint32_t &Hout, uint32_t &Gxout,
                                              uint32_t &mv1, uint32_t &mv2) {
    static_assert(K <= 16, "the moves of one lane must fit 32 bits");
    const uint32_t GE2 = GEC ? ((uint32_t)GEC & 0xffffu) * 0x10001u : GEpk;
    uint32_t Hdg = hdiag, m1 = 0, m2 = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t s = prmt(Rlo, Rhi, selS[k]);
        const uint32_t Gu = Gy[k];
        const uint32_t h = __vadd2(__vimax3_s16x2(Hdg, Gu, Gl), s);
        const uint32_t o = Hdg * one + GOc;                       // IMAD: both halves + GO (biased storage, see duo_row)
        const uint32_t gy = __viaddmax_s16x2(Gu, GE2, o);
        const uint32_t gx = __viaddmax_s16x2(Gl, GE2, o);
        bool pUhi, pUlo, pDhi, pDlo;
        const uint32_t g = vibmax_s16x2(gy, gx, pUhi, pUlo);      // gy >= gx
        (void)vibmax_s16x2(h, g, pDhi, pDlo);                     // h >= max(gy, gx)
        // @!p IMAD mv, mv, one, imm.  Measured alternatives (profiles/r02_add_placement.txt): written as one * imm + mv, ptxas
        // strength-reduces the power of two and selects / adds on the ALU pipe instead: 34 -> 60 ms on the 1.5 kb set
        m1 = pDlo ? m1 : m1 * one + (1u << (2 * k));
        m1 = pUlo ? m1 : m1 * one + (2u << (2 * k));
        m2 = pDhi ? m2 : m2 * one + (1u << (2 * k));
        m2 = pUhi ? m2 : m2 * one + (2u << (2 * k));
        Hdg = Hs[k];
        Hd[k] = h; Gy[k] = gy;
        Gl = gx;
    }
    Hout = Hd[K - 1]; Gxout = Gl; mv1 = m1; mv2 = m2;
}

// The same row on 4-bit base SETS (sequences with IUPAC codes, no gap character): hit vectors and pre-borrowed
// column scores as in sets_row (pa_dp_sets.cuh), moves as above.
template <int K, int GEC = 0, int DC = 0>
__device__ __forceinline__ void duo_moves_row_sets(const uint32_t (&Hs)[K], uint32_t (&Hd)[K], uint32_t (&Gy)[K],
                                                   const uint32_t (&HV)[K], const uint32_t (&XAc)[K], const uint32_t rowset,
                                                   const uint32_t Dmul, const uint32_t GOc, const uint32_t GEpk, const uint32_t one,
                                                   uint32_t hdiag, uint32_t Gl, uint32_t &Hout, uint32_t &Gxout,
                                                   uint32_t &mv1, uint32_t &mv2) {
    static_assert(K <= 16, "the moves of one lane must fit 32 bits");
    const uint32_t GE2 = GEC ? ((uint32_t)GEC & 0xffffu) * 0x10001u : GEpk;
    uint32_t Hdg = hdiag, m1 = 0, m2 = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t t2 = (HV[k] >> rowset) & 0x00010001u;
        const uint32_t Gu = Gy[k];
        const uint32_t m3x = __vimax3_s16x2(Hdg, Gu, Gl) + XAc[k];
        const uint32_t h = __vadd2(DC ? t2 * (uint32_t)DC : t2 * Dmul, m3x);
        const uint32_t o = Hdg * one + GOc;
        const uint32_t gy = __viaddmax_s16x2(Gu, GE2, o);
        const uint32_t gx = __viaddmax_s16x2(Gl, GE2, o);
        bool pUhi, pUlo, pDhi, pDlo;
        const uint32_t g = vibmax_s16x2(gy, gx, pUhi, pUlo);
        (void)vibmax_s16x2(h, g, pDhi, pDlo);
        m1 = pDlo ? m1 : m1 * one + (1u << (2 * k));
        m1 = pUlo ? m1 : m1 * one + (2u << (2 * k));
        m2 = pDhi ? m2 : m2 * one + (1u << (2 * k));
        m2 = pUhi ? m2 : m2 * one + (2u << (2 * k));
        Hdg = Hs[k];
        Hd[k] = h; Gy[k] = gy;
        Gl = gx;
    }
    Hout = Hd[K - 1]; Gxout = Gl; mv1 = m1; mv2 = m2;
}

struct BestDuo {
    int rowBest1, rowJ1, rowBest2, rowJ2;        // last row so far (columns ascending, strict >), true 32-bit values
    int colBest1, colI1, colBest2, colI2;        // last column (rows ascending, strict >); valid in lane 31
    __device__ __forceinline__ void reset(const int n) {
        rowBest1 = rowBest2 = INT_MIN; rowJ1 = rowJ2 = INT_MAX;
        colBest1 = colBest2 = INT_MIN; colI1 = colI2 = n - 1;
    }
};

// tab[z * 4 + x] = (Rlo, Rhi, -, -): scores of row code x against base codes 0-3, column 0 of pair 1 (byte 4), pad
// (byte 5, zero), column 0 of pair 2 (byte 6); z = 1: row 0 (+GO).  Same bytes as align_warp_duo's table.
__device__ __forceinline__ void build_tab_duo(int4 *tab, const Scoring sc, const uint32_t y10, const uint32_t y20, const int lane) {
    if (lane < 8) {
        const uint32_t xi = lane & 3u, z = lane >> 2;
        const int adj = z ? sc.go : 0;
        const uint32_t Mb = (uint32_t)(sc.match + adj) & 0xffu, Xb = (uint32_t)(sc.mismatch + adj) & 0xffu;
        const uint32_t Mz = (uint32_t)(sc.match + sc.go) & 0xffu, Xz = (uint32_t)(sc.mismatch + sc.go) & 0xffu;
        int4 e;
        e.x = (int)((Xb * 0x01010101u) ^ ((Mb ^ Xb) << (xi * 8u)));
        e.y = (int)(((xi == y10) ? Mz : Xz) | (((xi == y20) ? Mz : Xz) << 16));
        e.z = 0; e.w = 0;
        tab[z * 4 + xi] = e;
    }
}

// One block: slots [lane * K, lane * K + K) of a 32 * K wide strip, all n rows, for both pairs.  j_base1 / j_base2: the
// column of each pair that slot 0 of the block holds (negative: pad slots).  Edge rows are (H, Gx, off1, off2): packed
// stored values and the 32-bit offsets of the frame they were stored in (true value = stored + offset).
// dirp1 / dirp2: this lane's word of row 0 of the block in each pair's move store (nullptr: the pair has no column in
// this block, or its moves are not wanted); dstride: words per row.
// SETS: xs / ys1 / ys2 are 4-bit sets (8 per word) and a cell scores by intersection; tab is not used and row 0 is set in
// closed form (H = s, Gy = Gx = 0: the move is diagonal iff s >= 0, else up).
template <int K, class Edge, bool WIN, int GEC, bool SETS = false, int DC = 0>
__device__ __forceinline__ void block_duo_moves(const uint32_t *xs, const int n, const uint32_t *ys1, const uint32_t *ys2,
                                                const int j_base1, const int j_base2, const bool last_block, const Scoring sc,
                                                const int4 *tab, const Edge &edge, const int lane, BestDuo &best,
                                                uint32_t *dirp1, const uint32_t dstride1, uint32_t *dirp2, const uint32_t dstride2,
                                                uint2 *delay) {        // this warp's 2 x MOVES_DELAY x 32 delay ring
    const int B = WIN ? WIN_BIAS : sc.bias16;
    const uint32_t Bpk = pack16(B, B);
    const int Hinit = -sc.go + B;
    const uint32_t HinitPk = pack16(Hinit, Hinit);
    const uint32_t GOc = sc.go ? pack16(sc.go, sc.go - 1) : 0u, GEpk = pack16(sc.ge, sc.ge);
    const uint32_t one = (uint32_t)sc.one;
    const int j01 = j_base1 + lane * K, j02 = j_base2 + lane * K;
    const uint32_t Dmul = (uint32_t)(sc.match - sc.mismatch);
    uint32_t HX[K], HY[K], Gy[K], selS[K];          // SETS: selS holds the hit vectors
    uint32_t XAc[SETS ? K : 1];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j1 = j01 + k, j2 = j02 + k;
        HX[k] = HinitPk; HY[k] = HinitPk; Gy[k] = Bpk;
        if constexpr (SETS) {
            const uint32_t c1 = j1 >= 0 ? fetch4(ys1, j1) : 0u, c2 = j2 >= 0 ? fetch4(ys2, j2) : 0u;
            selS[k] = hit_vector(c1) | (hit_vector(c2) << 16);
            const int xa1 = j1 < 0 ? 0 : sc.mismatch + (j1 == 0 ? sc.go : 0);
            const int xa2 = j2 < 0 ? 0 : sc.mismatch + (j2 == 0 ? sc.go : 0);
            XAc[k] = pack16(xa1, xa2 - (xa1 < 0 ? 1 : 0));
        } else {
            const uint32_t c1 = j1 < 0 ? 5u : (j1 == 0 ? 4u : fetch2(ys1, j1));
            const uint32_t c2 = j2 < 0 ? 5u : (j2 == 0 ? 6u : fetch2(ys2, j2));
            selS[k] = ((8u | c2) << 12) | (c2 << 8) | ((8u | c1) << 4) | c1;
        }
    }
    uint32_t hprev = HinitPk;
    uint32_t HoA = HinitPk, GoA = Bpk, HoB = HinitPk, GoB = Bpk;
    int off1 = -B, off2 = -B;                         // true value = stored + off
    uint32_t colBestPk = 0x80008000u;                 // !WIN: last-column maxima of both pairs, packed like the scores
    const int n_steps = ((n + 1) >> 1) + 31;
    const int x_last_word = SETS ? (n - 1) >> 3 : (n - 1) >> 4;
    constexpr int CHUNK_STEPS = CHUNK_ROWS / 2;

    edge.acquire(min(n, 2 * CHUNK_ROWS), lane);
    int4 fA = edge.load(0), fB = edge.load(n > 1 ? 1 : 0);
    uint32_t xw = xs[0];

    for (int t = 0; t < n_steps; ++t) {
        const int iA = 2 * (t - lane);
        if ((t & (CHUNK_STEPS - 1)) == 0 && t > 0) {
            edge.consumed(2 * t, lane);
            edge.acquire(min(n, 2 * t + 2 * CHUNK_ROWS), lane);
        }
        if (t >= 31 && ((t - 31) & (CHUNK_STEPS - 1)) == 0 && edge.has_sink())
            edge.reserve(min(n, 2 * (t - 31) + CHUNK_ROWS), lane);
        uint32_t hinA = __shfl_up_sync(FULL_MASK, HoA, 1), ginA = __shfl_up_sync(FULL_MASK, GoA, 1);
        uint32_t hinB = __shfl_up_sync(FULL_MASK, HoB, 1), ginB = __shfl_up_sync(FULL_MASK, GoB, 1);
        if (lane == 0) {
            hinA = (uint32_t)fA.x; ginA = (uint32_t)fA.y;
            hinB = (uint32_t)fB.x; ginB = (uint32_t)fB.y;
        }
        if (WIN) {   // from the left lane's frame (lane 0: the frames the two edge rows were stored in) into mine
            int oA1 = __shfl_up_sync(FULL_MASK, off1, 1), oA2 = __shfl_up_sync(FULL_MASK, off2, 1);
            int oB1 = oA1, oB2 = oA2;
            if (lane == 0) { oA1 = fA.z; oA2 = fA.w; oB1 = fB.z; oB2 = fB.w; }
            const uint32_t dA = pack16(oA1 - off1, oA2 - off2), dB = pack16(oB1 - off1, oB2 - off2);
            hinA = __vadd2(hinA, dA); ginA = __vadd2(ginA, dA);
            hinB = __vadd2(hinB, dB); ginB = __vadd2(ginB, dB);
        }
        fA = edge.load(min(2 * t + 2, n - 1));
        fB = edge.load(min(2 * t + 3, n - 1));
        // codes of rows iA and iA + 1: two 2-bit codes, or (SETS) two 4-bit sets
        const uint32_t xi2 = SETS ? (xw >> ((iA & 7) * 4)) & 0xffu : (xw >> ((iA & 15) * 2)) & 15u;
        xw = xs[min(max(iA + 2, 0) >> (SETS ? 3 : 4), x_last_word)];
        uint2 mvP1 = make_uint2(0u, 0u), mvP2 = mvP1;      // moves of rows iA (x) and iA + 1 (y), pair 1 and pair 2
        if (iA >= 0 && iA < n) {
            const bool store = (lane == 31) && edge.has_sink();
            {
                if constexpr (SETS) {
                    if (iA == 0) {      // first row in closed form (values in this lane's frame: off = -B here)
                        const bool dM = sc.match >= 0, dX = sc.mismatch >= 0;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const uint32_t t2 = (selS[k] >> (xi2 & 15u)) & 0x00010001u;
                            const bool real1 = j01 + k >= 0, real2 = j02 + k >= 0;
                            const bool hit1 = (t2 & 1u) != 0, hit2 = (t2 >> 16) != 0;
                            HY[k] = pack16(real1 ? (hit1 ? sc.match : sc.mismatch) + B : Hinit, real2 ? (hit2 ? sc.match : sc.mismatch) + B : Hinit);
                            Gy[k] = Bpk;
                            if (!(hit1 ? dM : dX)) mvP1.x |= 1u << (2 * k);       // not diagonal -> up (Gy = Gx = 0)
                            if (!(hit2 ? dM : dX)) mvP2.x |= 1u << (2 * k);
                        }
                        HoA = HY[K - 1]; GoA = Bpk;
                    } else {
                        duo_moves_row_sets<K, GEC, DC>(HX, HY, Gy, selS, XAc, xi2 & 15u, Dmul, GOc, GEpk, one, hprev, ginA, HoA, GoA, mvP1.x, mvP2.x);
                    }
                } else {
                    const int4 T = tab[(iA == 0 ? 4 : 0) + (xi2 & 3u)];
                    duo_moves_row<K, GEC>(HX, HY, Gy, selS, (uint32_t)T.x, (uint32_t)T.y, GOc, GEpk, one, hprev, ginA, HoA, GoA, mvP1.x, mvP2.x);
                }
                if (store) edge.store(iA, make_int4((int)HoA, (int)GoA, off1, off2));
                if (last_block) {
                    if (WIN) {
                        const int t1 = lo16(HoA) + off1, t2 = hi16(HoA) + off2;
                        if (t1 > best.colBest1) { best.colBest1 = t1; best.colI1 = iA; }
                        if (t2 > best.colBest2) { best.colBest2 = t2; best.colI2 = iA; }
                    } else {
                        bool ghi, glo;
                        colBestPk = vibmax_s16x2(colBestPk, HoA, ghi, glo);
                        if (!glo) best.colI1 = iA;
                        if (!ghi) best.colI2 = iA;
                    }
                }
            }
            if (iA + 1 < n) {
                if constexpr (SETS) {
                    duo_moves_row_sets<K, GEC, DC>(HY, HX, Gy, selS, XAc, xi2 >> 4, Dmul, GOc, GEpk, one, hinA, ginB, HoB, GoB, mvP1.y, mvP2.y);
                } else {
                    const int4 T = tab[xi2 >> 2];
                    duo_moves_row<K, GEC>(HY, HX, Gy, selS, (uint32_t)T.x, (uint32_t)T.y, GOc, GEpk, one, hinA, ginB, HoB, GoB, mvP1.y, mvP2.y);
                }
                if (store) edge.store(iA + 1, make_int4((int)HoB, (int)GoB, off1, off2));
                if (last_block) {
                    if (WIN) {
                        const int t1 = lo16(HoB) + off1, t2 = hi16(HoB) + off2;
                        if (t1 > best.colBest1) { best.colBest1 = t1; best.colI1 = iA + 1; }
                        if (t2 > best.colBest2) { best.colBest2 = t2; best.colI2 = iA + 1; }
                    } else {
                        bool ghi, glo;
                        colBestPk = vibmax_s16x2(colBestPk, HoB, ghi, glo);
                        if (!glo) best.colI1 = iA + 1;
                        if (!ghi) best.colI2 = iA + 1;
                    }
                }
            }
            hprev = hinB;
            if (WIN) {   // keep this lane's window centred on its right edge
                const uint32_t edge_h = (iA + 1 < n) ? HoB : HoA;
                const int v1 = lo16(edge_h) - B, v2 = hi16(edge_h) - B;
                const int d1 = v1 > WIN_T ? WIN_Q : (v1 < -WIN_T ? -WIN_Q : 0);
                const int d2 = v2 > WIN_T ? WIN_Q : (v2 < -WIN_T ? -WIN_Q : 0);
                if (d1 | d2) {
                    const uint32_t dPk = pack16(d1, d2);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        HX[k] = __vsub2(HX[k], dPk); HY[k] = __vsub2(HY[k], dPk); Gy[k] = __vsub2(Gy[k], dPk);
                    }
                    hprev = __vsub2(hprev, dPk);
                    HoA = __vsub2(HoA, dPk); GoA = __vsub2(GoA, dPk); HoB = __vsub2(HoB, dPk); GoB = __vsub2(GoB, dPk);
                    off1 += d1; off2 += d2;
                }
            }
        }
        {   // delayed, sector-aligned stores of the moves (see the header): this lane's words of j7 steps ago
            const int j7 = 7 - (lane & 7);
            delay[(t & (MOVES_DELAY - 1)) * 32 + lane] = mvP1;
            delay[(MOVES_DELAY + (t & (MOVES_DELAY - 1))) * 32 + lane] = mvP2;
            const uint2 w1 = delay[((t - j7) & (MOVES_DELAY - 1)) * 32 + lane];
            const uint2 w2 = delay[(MOVES_DELAY + ((t - j7) & (MOVES_DELAY - 1))) * 32 + lane];
            const int iD = iA - 2 * j7;
            if (iD >= 0 && iD < n) {
                if (dirp1) { __stcs(&dirp1[(size_t)iD * dstride1], w1.x); if (iD + 1 < n) __stcs(&dirp1[(size_t)(iD + 1) * dstride1], w1.y); }
                if (dirp2) { __stcs(&dirp2[(size_t)iD * dstride2], w2.x); if (iD + 1 < n) __stcs(&dirp2[(size_t)(iD + 1) * dstride2], w2.y); }
            }
        }
        if (((t - 31) & (CHUNK_STEPS - 1)) == CHUNK_STEPS - 1 && t >= 31) edge.release(min(n, 2 * (t - 31) + 2), lane);
    }
    edge.release(n, lane);
    edge.consumed(n, lane);
    __syncwarp();
    if (last_block && !WIN) {   // scores stay far above -32768 (host-checked range), so the sentinel is never a real value
        best.colBest1 = lo16(colBestPk) - B; best.colBest2 = hi16(colBestPk) - B;
    }
    // last row of this block (in Y when n is odd): columns ascending, strict >
    int bv1 = INT_MIN, bj1 = INT_MAX, bv2 = INT_MIN, bj2 = INT_MAX;
    const bool in_y = (n & 1);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j1 = j01 + k, j2 = j02 + k;
        const uint32_t hk = in_y ? HY[k] : HX[k];
        const int h1 = lo16(hk) + off1, h2 = hi16(hk) + off2;
        if (j1 >= 0 && h1 > bv1) { bv1 = h1; bj1 = j1; }
        if (j2 >= 0 && h2 > bv2) { bv2 = h2; bj2 = j2; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int ov1 = __shfl_xor_sync(FULL_MASK, bv1, d), oj1 = __shfl_xor_sync(FULL_MASK, bj1, d);
        if (ov1 > bv1 || (ov1 == bv1 && oj1 < bj1)) { bv1 = ov1; bj1 = oj1; }
        const int ov2 = __shfl_xor_sync(FULL_MASK, bv2, d), oj2 = __shfl_xor_sync(FULL_MASK, bj2, d);
        if (ov2 > bv2 || (ov2 == bv2 && oj2 < bj2)) { bv2 = ov2; bj2 = oj2; }
    }
    if (bj1 != INT_MAX && (bv1 > best.rowBest1 || (bv1 == best.rowBest1 && bj1 < best.rowJ1))) { best.rowBest1 = bv1; best.rowJ1 = bj1; }
    if (bj2 != INT_MAX && (bv2 > best.rowBest2 || (bv2 == best.rowBest2 && bj2 < best.rowJ2))) { best.rowBest2 = bv2; best.rowJ2 = bj2; }
}

// score and end cell of one pair; the walk fills in the compared / differing columns
__device__ __forceinline__ void finish_moves(const int rowBest, const int rowJ, const int colBest, const int colI,
                                             const int n, const int m, pa_pair_result *res) {
    pa_pair_result o;
    o.dist = 0; o.len = 0;
    if (rowBest > colBest) { o.score = rowBest; o.end_i = n - 1; o.end_j = rowJ; }
    else                   { o.score = colBest; o.end_i = colI;  o.end_j = m - 1; }
    *res = o;
}

// Work item w = (items[w].x, items[w].y): entries of the pair list (ia, ib) with ia equal; .y == 0xffffffff: one entry.
struct MovesItem {
    uint32_t a, y1, y2;
    int n, m1, m2;
    uint32_t e1, e2;
    bool two;
};
__device__ __forceinline__ MovesItem load_moves_item(const SeqStore &S, const uint32_t *ia, const uint32_t *ib, const uint2 it) {
    MovesItem r;
    r.e1 = it.x; r.two = it.y != 0xffffffffu; r.e2 = r.two ? it.y : it.x;
    r.a = ia[r.e1]; r.y1 = ib[r.e1]; r.y2 = ib[r.e2];
    r.n = (int)S.len[r.a]; r.m1 = (int)S.len[r.y1]; r.m2 = (int)S.len[r.y2];
    return r;
}

// SETS: the items hold a sequence with IUPAC codes (none with a gap character): 4-bit sets instead of 2-bit codes.
template <int GEC = 0, bool SETS = false, int DC = 0>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, MOVES_WARP_MINB)
pa_warp_duo_moves_kernel(const SeqStore S, const Scoring sc, const uint32_t *ia, const uint32_t *ib, const uint2 *items,
                         const uint32_t n_items, const uint32_t max_len16, unsigned long long *work_counter, int4 *bbuf_all,
                         const uint32_t bbuf_rows, pa_pair_result *out, uint8_t *dirs, const unsigned long long *dirs_off) {
    __shared__ __align__(16) uint32_t stage[WARPS_PER_CTA][3][STAGE_WORDS];
    __shared__ int4 tabs[WARPS_PER_CTA][8];
    __shared__ __align__(8) uint2 delays[WARPS_PER_CTA][2 * MOVES_DELAY * 32];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * WARPS_PER_CTA + wib;
    int4 *bbuf = bbuf_all + (size_t)gw * bbuf_rows;
    const uint32_t vrow = bbuf_rows - 1;
    constexpr int W = 32 * KMOV;
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1ull);
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w >= n_items) break;
        const MovesItem it = load_moves_item(S, ia, ib, items[w]);
        __syncwarp();
        const uint32_t *xs, *ys1, *ys2;
        if (SETS) {
            xs = stage_seq(S.p4 + S.off4[it.a], (uint32_t)(it.n + 7) >> 3, stage[wib][0], lane);
            ys1 = stage_seq(S.p4 + S.off4[it.y1], (uint32_t)(it.m1 + 7) >> 3, stage[wib][1], lane);
            ys2 = stage_seq(S.p4 + S.off4[it.y2], (uint32_t)(it.m2 + 7) >> 3, stage[wib][2], lane);
        } else {
            xs = stage_seq(S.p2 + S.off2[it.a], (uint32_t)(it.n + 15) >> 4, stage[wib][0], lane);
            ys1 = stage_seq(S.p2 + S.off2[it.y1], (uint32_t)(it.m1 + 15) >> 4, stage[wib][1], lane);
            ys2 = stage_seq(S.p2 + S.off2[it.y2], (uint32_t)(it.m2 + 15) >> 4, stage[wib][2], lane);
            build_tab_duo(tabs[wib], sc, fetch2(S.p2 + S.off2[it.y1], 0), fetch2(S.p2 + S.off2[it.y2], 0), lane);
        }
        const bool win = (uint32_t)it.n > max_len16 || (uint32_t)it.m1 > max_len16 || (uint32_t)it.m2 > max_len16;
        const int B = win ? WIN_BIAS : sc.bias16;
        if (lane == 0) __stcg(&bbuf[vrow], make_int4((int)pack16(-sc.go + B, -sc.go + B), (int)pack16(B, B), -B, -B));
        __syncwarp();
        const int mmax = max(it.m1, it.m2);
        const int P = (mmax + W - 1) / W, P1 = (it.m1 + W - 1) / W, P2 = (it.m2 + W - 1) / W;
        const int pad1 = P * W - it.m1, pad2 = P * W - it.m2;
        uint32_t *d1 = reinterpret_cast<uint32_t *>(dirs + dirs_off[it.e1]), *d2 = reinterpret_cast<uint32_t *>(dirs + dirs_off[it.e2]);
        BestDuo best;
        best.reset(it.n);
        for (int p = 0; p < P; ++p) {
            GlobalEdge edge;
            edge.feed = p > 0 ? bbuf : bbuf + vrow;
            edge.fmul = p > 0 ? 1 : 0;
            edge.sink = p < P - 1 ? bbuf : nullptr;
            const int b1 = p - (P - P1), b2 = p - (P - P2);      // block index in each pair's own geometry
            uint32_t *q1 = b1 >= 0 ? d1 + b1 * 32 + lane : nullptr;
            uint32_t *q2 = (it.two && b2 >= 0) ? d2 + b2 * 32 + lane : nullptr;
            if (win) block_duo_moves<KMOV, GlobalEdge, true, GEC, SETS, DC>(xs, it.n, ys1, ys2, p * W - pad1, p * W - pad2, p == P - 1, sc, tabs[wib],
                                                                            edge, lane, best, q1, (uint32_t)P1 * 32u, q2, (uint32_t)P2 * 32u, delays[wib]);
            else block_duo_moves<KMOV, GlobalEdge, false, GEC, SETS, DC>(xs, it.n, ys1, ys2, p * W - pad1, p * W - pad2, p == P - 1, sc, tabs[wib],
                                                                         edge, lane, best, q1, (uint32_t)P1 * 32u, q2, (uint32_t)P2 * 32u, delays[wib]);
        }
        best.colBest1 = __shfl_sync(FULL_MASK, best.colBest1, 31); best.colI1 = __shfl_sync(FULL_MASK, best.colI1, 31);
        best.colBest2 = __shfl_sync(FULL_MASK, best.colBest2, 31); best.colI2 = __shfl_sync(FULL_MASK, best.colI2, 31);
        if (lane == 0) {
            finish_moves(best.rowBest1, best.rowJ1, best.colBest1, best.colI1, it.n, it.m1, &out[it.e1]);
            if (it.two) finish_moves(best.rowBest2, best.rowJ2, best.colBest2, best.colI2, it.n, it.m2, &out[it.e2]);
        }
    }
}

// One item per CTA.  gedge_all: per CTA one column of n rows for the wrap-around edge (last warp -> first warp's next
// block) plus the virtual row.  Always the floating-window form (only pairs beyond LONG_LEN come here).
template <int GEC = 0, bool SETS = false, int DC = 0>
__global__ void __launch_bounds__(MOVES_CTA_WARPS * 32, 1)
pa_cta_duo_moves_kernel(const SeqStore S, const Scoring sc, const uint32_t *ia, const uint32_t *ib, const uint2 *items,
                        const uint32_t n_items, unsigned long long *work_counter, int4 *gedge_all, const uint32_t gedge_rows,
                        pa_pair_result *out, uint8_t *dirs, const unsigned long long *dirs_off) {
    constexpr int NW = MOVES_CTA_WARPS;
    // dynamic shared memory (moves_cta_smem(SETS) bytes, above the 48 KB a kernel gets statically): x (twice the words as
    // 4-bit sets), the edge rings, the delay rings
    constexpr int XW = SETS ? 2 * XSTAGE_WORDS : XSTAGE_WORDS;
    extern __shared__ __align__(16) unsigned char moves_smem[];
    uint32_t *xstage = reinterpret_cast<uint32_t *>(moves_smem);
    int4 (*rings)[MOVES_RING_ROWS] = reinterpret_cast<int4 (*)[MOVES_RING_ROWS]>(moves_smem + XW * 4);
    uint2 *delays = reinterpret_cast<uint2 *>(moves_smem + XW * 4 + (NW - 1) * MOVES_RING_ROWS * 16);
    __shared__ int4 tab[8];
    __shared__ int prod[NW], cons[NW];
    __shared__ unsigned long long item_s;
    __shared__ int bRow1[NW], bJ1[NW], bRow2[NW], bJ2[NW], bCol1, bColI1, bCol2, bColI2;
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    int4 *gedge = gedge_all + (size_t)blockIdx.x * gedge_rows;
    const uint32_t vrow = gedge_rows - 1;
    constexpr int W = 32 * KMOV;
    constexpr int B = WIN_BIAS;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) item_s = atomicAdd(work_counter, 1ull);
        __syncthreads();
        const unsigned long long wi = item_s;
        if (wi >= n_items) break;
        const MovesItem it = load_moves_item(S, ia, ib, items[wi]);
        {
            const uint4 *g4 = reinterpret_cast<const uint4 *>(SETS ? S.p4 + S.off4[it.a] : S.p2 + S.off2[it.a]);
            uint4 *s4 = reinterpret_cast<uint4 *>(xstage);
            const uint32_t n4 = ((SETS ? (uint32_t)(it.n + 7) >> 3 : (uint32_t)(it.n + 15) >> 4) + 3) >> 2;
            for (uint32_t q = threadIdx.x; q < n4; q += blockDim.x) s4[q] = __ldg(&g4[q]);
        }
        const uint32_t *ys1 = SETS ? S.p4 + S.off4[it.y1] : S.p2 + S.off2[it.y1];
        const uint32_t *ys2 = SETS ? S.p4 + S.off4[it.y2] : S.p2 + S.off2[it.y2];
        if (!SETS && w == 0) build_tab_duo(tab, sc, fetch2(ys1, 0), fetch2(ys2, 0), lane);
        if (threadIdx.x < NW) { prod[threadIdx.x] = 0; cons[threadIdx.x] = 0; }
        if (threadIdx.x == 0) {
            __stcg(&gedge[vrow], make_int4((int)pack16(-sc.go + B, -sc.go + B), (int)pack16(B, B), -B, -B));
            bCol1 = bCol2 = INT_MIN; bColI1 = bColI2 = it.n - 1;
        }
        __syncthreads();

        const int n = it.n;
        const int mmax = max(it.m1, it.m2);
        const int P = (mmax + W - 1) / W, P1 = (it.m1 + W - 1) / W, P2 = (it.m2 + W - 1) / W;
        const int pad1 = P * W - it.m1, pad2 = P * W - it.m2;
        uint32_t *d1 = reinterpret_cast<uint32_t *>(dirs + dirs_off[it.e1]), *d2 = reinterpret_cast<uint32_t *>(dirs + dirs_off[it.e2]);
        BestDuo best;
        best.reset(n);
        int round = 0;
        for (int p = w; p < P; p += NW, ++round) {
            MovesRing edge;
            const int prev = (w + NW - 1) % NW;
            edge.in_ring = nullptr; edge.in_global = gedge + vrow; edge.in_mul = 0; edge.in_prod = nullptr; edge.in_cons = nullptr;
            edge.in_base = 0;
            if (p > 0) {
                if (w == 0) { edge.in_global = gedge; edge.in_mul = 1; edge.in_prod = &prod[prev]; edge.in_base = (round - 1) * n; }
                else { edge.in_ring = rings[prev]; edge.in_prod = &prod[prev]; edge.in_cons = &cons[prev]; edge.in_base = round * n; }
            }
            edge.out_ring = nullptr; edge.out_global = nullptr; edge.out_prod = nullptr; edge.out_cons = nullptr; edge.out_base = round * n;
            if (p < P - 1) {
                edge.out_prod = &prod[w];
                if (w == NW - 1) edge.out_global = gedge;
                else { edge.out_ring = rings[w]; edge.out_cons = &cons[w]; }
            }
            const int b1 = p - (P - P1), b2 = p - (P - P2);
            uint32_t *q1 = b1 >= 0 ? d1 + b1 * 32 + lane : nullptr;
            uint32_t *q2 = (it.two && b2 >= 0) ? d2 + b2 * 32 + lane : nullptr;
            block_duo_moves<KMOV, MovesRing, true, GEC, SETS, DC>(xstage, n, ys1, ys2, p * W - pad1, p * W - pad2, p == P - 1, sc, tab, edge, lane,
                                                       best, q1, (uint32_t)P1 * 32u, q2, (uint32_t)P2 * 32u, delays + w * 2 * MOVES_DELAY * 32);
        }
        if (lane == 0) { bRow1[w] = best.rowBest1; bJ1[w] = best.rowJ1; bRow2[w] = best.rowBest2; bJ2[w] = best.rowJ2; }
        if (lane == 31 && ((P - 1) % NW) == w) { bCol1 = best.colBest1; bColI1 = best.colI1; bCol2 = best.colBest2; bColI2 = best.colI2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int r1 = INT_MIN, j1 = INT_MAX, r2 = INT_MIN, j2 = INT_MAX;
            for (int q = 0; q < NW; ++q) {
                if (bJ1[q] != INT_MAX && (bRow1[q] > r1 || (bRow1[q] == r1 && bJ1[q] < j1))) { r1 = bRow1[q]; j1 = bJ1[q]; }
                if (bJ2[q] != INT_MAX && (bRow2[q] > r2 || (bRow2[q] == r2 && bJ2[q] < j2))) { r2 = bRow2[q]; j2 = bJ2[q]; }
            }
            finish_moves(r1, j1, bCol1, bColI1, n, it.m1, &out[it.e1]);
            if (it.two) finish_moves(r2, j2, bCol2, bColI2, n, it.m2, &out[it.e2]);
        }
    }
}

}  // namespace pa
