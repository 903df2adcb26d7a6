// pa_dp_sets.cuh -- the s16x2 two-pairs-per-warp kernel for sequences with IUPAC ambiguity codes (any density,
// no gap character): same wavefront, same biased 16-bit storage, same floating window for long pairs as
// align_warp_duo (pa_dp.cuh), but rows and columns are 4-bit base SETS and a cell scores `match` when the two
// sets intersect, `mismatch` otherwise (seqpair::cost, src/seqpair.cpp:191-236 with the 7 / -5 matrix of
// src/pairalign.cpp:682: every pair of sets that share a base scores the match value).
//
// Per column the lane keeps three registers, all set up once per pass:
//   HV  hit vectors of the two pairs' column sets: bit r (pair 1) / bit 16 + r (pair 2) is set iff (r & set) != 0,
//       so ONE funnel-free shift by the row's set r and one mask give both hit bits: t2 = (HV >> r) & 0x00010001;
//   XAc what every cell of the column scores before the hit bonus: mismatch (+GO in column 0, 0 in a pad slot),
//       packed and pre-borrowed (high half minus one when the low half is negative) so that adding it to the
//       negative-biased max3 is a plain 32-bit add -- on the FMA pipe, like H + GO in duo_row;
//   LC  the "this slot is a real column" flags of the two pairs, at the bit where each pair's counter keeps its
//       column count.
// A cell then costs the ALU pipe 12 instructions for two DP cells -- SHF, LOP3 (hit bits), VIMNMX3, VIADD.16x2,
// 2 VIADDMNMX, 2 VIMNMX with predicates, 2 LOP3 (counter increments), 2 predicated SEL -- one more than the PRMT
// kernel pays, plus 5 on the FMA pipe (H+GO, max3+XAc, hit*(match-mismatch), two counter adds); on this chip an
// ALU-pipe instruction holds the scheduler's dispatch port for 2 cycles and a two-register IMAD for 1, so a pair of
// cells costs 29 dispatch cycles against the PRMT kernel's 25.
//
// Counters: pair 1 keeps (columns << 16 | matching columns), pair 2 (matching columns << 16 | columns), so both
// increments are one bit-select each out of (t2, LC); mismatches = columns - matches at the very end.
//
// Row 0 is not run through the recurrence: H = s, Gy = Gx = 0 (src/seqpair.cpp:103-107) in closed form, once per
// pass and lane.  The other rows need no first-row table and no per-row adjustment.
#pragma once

#include "pa_dp.cuh"

namespace pa {

__device__ __forceinline__ uint32_t hit_vector(const uint32_t set) {       // bit r = ((r & set) != 0), r = 0..15
    return ((set & 1u) ? 0xAAAAu : 0u) | ((set & 2u) ? 0xCCCCu : 0u) | ((set & 4u) ? 0xF0F0u : 0u) | ((set & 8u) ? 0xFF00u : 0u);
}

// (a & ~m) | (b & m) as ONE LOP3 the optimiser cannot take apart (it would otherwise split the masks off and fold
// the rest into three-input IADD3s on the ALU pipe, which is the pipe this kernel is bound by)
__device__ __forceinline__ uint32_t bitsel(const uint32_t a, const uint32_t b, const uint32_t m) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
}

// DC != 0: match - mismatch is the compile-time constant DC (12 for pairalign's scoring) and multiplies the hit bits as an
// immediate; an IMAD with a register multiplier costs several times more (profiles/r02_add_placement.txt).
template <int K, int GEC = 0, int DC = 0>
__device__ __forceinline__ void sets_row(const uint32_t (&Hs)[K], uint32_t (&Hd)[K], uint32_t (&Gy)[K],
                                         const uint32_t (&C1s)[K], uint32_t (&C1d)[K],
                                         const uint32_t (&C2s)[K], uint32_t (&C2d)[K],
                                         const uint32_t (&HV)[K], const uint32_t (&XAc)[K], const uint32_t (&LC)[K],
                                         const uint32_t rowset, const uint32_t Dmul, const uint32_t GOc, const uint32_t GEpk,
                                         uint32_t hdiag, uint32_t Gl, uint32_t cd1, uint32_t cd2, uint32_t cl1, uint32_t cl2,
                                         uint32_t &Hout, uint32_t &Gxout, uint32_t &c1out, uint32_t &c2out) {
    const uint32_t GE2 = GEC ? ((uint32_t)GEC & 0xffffu) * 0x10001u : GEpk;
    uint32_t Hdg = hdiag;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t t2 = (HV[k] >> rowset) & 0x00010001u;              // hit bit of pair 1 | hit bit of pair 2 << 16
        const uint32_t inc1 = bitsel(t2, LC[k], 0x00010000u);             // column flag << 16 | hit
        const uint32_t inc2 = bitsel(t2, LC[k], 0x00000001u);             // hit << 16 | column flag
        const uint32_t Gu = Gy[k];
        const uint32_t cu1 = C1s[k], cu2 = C2s[k];
        const uint32_t m3x = __vimax3_s16x2(Hdg, Gu, Gl) + XAc[k];        // 32-bit add, exact per half (pre-borrowed)
        // + (match - mismatch) where the sets intersect.  A packed add on purpose: when h is the result of a 32-bit IMAD,
        // ptxas leaves a dead PRMT per cell behind the predicate-producing VIMNMX.S16x2 below (same ALU-pipe cost, one
        // FMA-pipe instruction more)
        const uint32_t h = __vadd2(DC ? t2 * (uint32_t)DC : t2 * Dmul, m3x);
        const uint32_t o = Hdg + GOc;
        const uint32_t gy = __viaddmax_s16x2(Gu, GE2, o);
        const uint32_t gx = __viaddmax_s16x2(Gl, GE2, o);
        bool pUhi, pUlo, pDhi, pDlo;
        const uint32_t g = vibmax_s16x2(gy, gx, pUhi, pUlo);
        (void)vibmax_s16x2(h, g, pDhi, pDlo);                             // h >= max(gy, gx)
        const uint32_t cdi1 = cd1 + inc1, cdi2 = cd2 + inc2;
        const uint32_t c1 = pDlo ? cdi1 : (pUlo ? cu1 : cl1);
        const uint32_t c2 = pDhi ? cdi2 : (pUhi ? cu2 : cl2);
        Hdg = Hs[k]; cd1 = cu1; cd2 = cu2;
        Hd[k] = h; Gy[k] = gy; C1d[k] = c1; C2d[k] = c2;
        Gl = gx; cl1 = c1; cl2 = c2;
    }
    Hout = Hd[K - 1]; Gxout = Gl; c1out = cl1; c2out = cl2;
}

// xs / ys1 / ys2: 4-bit sets, 8 per word.  Everything else as align_warp_duo.
template <int K, bool WIN = false, int GEC = 0, int DC = 0>
__device__ __forceinline__ void align_warp_sets(const uint32_t *xs, const int n, const uint32_t *ys1, const int m1,
                                                const uint32_t *ys2, const int m2, const Scoring sc, int4 *bbuf,
                                                const uint32_t vrow, pa_pair_result *res1, pa_pair_result *res2, const int lane) {
    constexpr int W = 32 * K;
    const int mmax = m1 > m2 ? m1 : m2;
    const int P = (mmax + W - 1) / W;
    const int pad1 = P * W - m1, pad2 = P * W - m2;
    const int B = WIN ? WIN_BIAS : sc.bias16;
    const uint32_t Bpk = pack16(B, B);
    const int Hinit = -sc.go + B;
    const uint32_t HinitPk = pack16(Hinit, Hinit);
    const uint32_t GOc = sc.go ? pack16(sc.go, sc.go - 1) : 0u, GEpk = pack16(sc.ge, sc.ge);
    const uint32_t Dmul = (uint32_t)(sc.match - sc.mismatch);            // >= 0 (host-checked)
    const bool dM = sc.match >= 0, dX = sc.mismatch >= 0;                // first row: the move is D iff the score is >= 0

    int rowBest1 = INT_MIN, rowJ1 = 0, rowBest2 = INT_MIN, rowJ2 = 0;
    uint32_t rowC1 = 0, rowC2 = 0;
    uint32_t colBestPk = 0x80008000u;
    int colB1 = INT_MIN, colB2 = INT_MIN;
    int colI1 = n - 1, colI2 = n - 1;
    uint32_t colC1 = 0, colC2 = 0;

    if (lane == 0) {
        __stcg(&bbuf[vrow], make_int4((int)HinitPk, (int)Bpk, 0, 0));
        if (WIN) __stcg(&bbuf[vrow + 1], make_int4(-B, -B, 0, 0));
    }
    __syncwarp();
    const int n_steps = ((n + 1) >> 1) + 31;
    const int x_last_word = (n - 1) >> 3;

    for (int p = 0; p < P; ++p) {
        const bool last_pass = (p == P - 1);
        const int s0 = p * W + lane * K;
        uint32_t HX[K], HY[K], Gy[K], C1X[K], C1Y[K], C2X[K], C2Y[K], HV[K], XAc[K], LC[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j1 = s0 + k - pad1, j2 = s0 + k - pad2;
            HX[k] = HinitPk; HY[k] = HinitPk; Gy[k] = Bpk; C1X[k] = 0; C2X[k] = 0; C1Y[k] = 0; C2Y[k] = 0;
            const uint32_t c1 = j1 >= 0 ? fetch4(ys1, j1) : 0u, c2 = j2 >= 0 ? fetch4(ys2, j2) : 0u;
            HV[k] = hit_vector(c1) | (hit_vector(c2) << 16);
            const int xa1 = j1 < 0 ? 0 : sc.mismatch + (j1 == 0 ? sc.go : 0);
            const int xa2 = j2 < 0 ? 0 : sc.mismatch + (j2 == 0 ? sc.go : 0);
            XAc[k] = pack16(xa1, xa2 - (xa1 < 0 ? 1 : 0));
            LC[k] = (j1 >= 0 ? 0x00010000u : 0u) | (j2 >= 0 ? 1u : 0u);
        }
        uint32_t hprev = HinitPk, c1prev = 0, c2prev = 0;
        uint32_t HoA = HinitPk, GoA = Bpk, c1oA = 0, c2oA = 0, HoB = HinitPk, GoB = Bpk, c1oB = 0, c2oB = 0;
        const int4 *feed = p > 0 ? bbuf : bbuf + vrow;
        const int fmul = p > 0 ? (WIN ? 2 : 1) : 0;
        int4 fA = __ldcg(&feed[0]), fB = __ldcg(&feed[fmul * (n > 1 ? 1 : 0)]);
        int4 fOA = make_int4(0, 0, 0, 0), fOB = fOA;
        if (WIN) { fOA = __ldcg(&feed[1]); fOB = __ldcg(&feed[fmul * (n > 1 ? 1 : 0) + 1]); }
        int off1 = -B, off2 = -B;
        uint32_t xw = xs[min(max(-2 * lane, 0) >> 3, x_last_word)];

        for (int t = 0; t < n_steps; ++t) {
            const int iA = 2 * (t - lane);
            uint32_t hinA = __shfl_up_sync(FULL_MASK, HoA, 1), ginA = __shfl_up_sync(FULL_MASK, GoA, 1);
            uint32_t c1inA = __shfl_up_sync(FULL_MASK, c1oA, 1), c2inA = __shfl_up_sync(FULL_MASK, c2oA, 1);
            uint32_t hinB = __shfl_up_sync(FULL_MASK, HoB, 1), ginB = __shfl_up_sync(FULL_MASK, GoB, 1);
            uint32_t c1inB = __shfl_up_sync(FULL_MASK, c1oB, 1), c2inB = __shfl_up_sync(FULL_MASK, c2oB, 1);
            if (lane == 0) {
                hinA = (uint32_t)fA.x; ginA = (uint32_t)fA.y; c1inA = (uint32_t)fA.z; c2inA = (uint32_t)fA.w;
                hinB = (uint32_t)fB.x; ginB = (uint32_t)fB.y; c1inB = (uint32_t)fB.z; c2inB = (uint32_t)fB.w;
            }
            if (WIN) {
                int oA1 = __shfl_up_sync(FULL_MASK, off1, 1), oA2 = __shfl_up_sync(FULL_MASK, off2, 1);
                int oB1 = oA1, oB2 = oA2;
                if (lane == 0) { oA1 = fOA.x; oA2 = fOA.y; oB1 = fOB.x; oB2 = fOB.y; }
                const uint32_t dA = pack16(oA1 - off1, oA2 - off2), dB = pack16(oB1 - off1, oB2 - off2);
                hinA = __vadd2(hinA, dA); ginA = __vadd2(ginA, dA);
                hinB = __vadd2(hinB, dB); ginB = __vadd2(ginB, dB);
            }
            fA = __ldcg(&feed[fmul * min(2 * t + 2, n - 1)]);
            fB = __ldcg(&feed[fmul * min(2 * t + 3, n - 1)]);
            if (WIN) {
                fOA = __ldcg(&feed[fmul * min(2 * t + 2, n - 1) + 1]);
                fOB = __ldcg(&feed[fmul * min(2 * t + 3, n - 1) + 1]);
            }
            const uint32_t xi2 = (xw >> ((iA & 7) * 4)) & 0xffu;       // sets of rows iA (low nibble) and iA + 1
            xw = xs[min(max(iA + 2, 0) >> 3, x_last_word)];
            if (iA >= 0 && iA < n) {
                const bool store = (lane == 31) && !last_pass;
                const uint32_t mA = xi2 & 15u, mB = xi2 >> 4;
                if (iA == 0) {
                    // first row in closed form: H = s, Gy = Gx = 0; the move is D iff s >= 0 and the counters start there.
                    // Pad slots keep their fixed point.  Values are stored in this lane's frame (off = -B here).
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const uint32_t t2 = (HV[k] >> mA) & 0x00010001u;
                        const bool real1 = (LC[k] & 0x00010000u) != 0, real2 = (LC[k] & 1u) != 0;
                        const bool hit1 = (t2 & 1u) != 0, hit2 = (t2 >> 16) != 0;
                        const int h1 = real1 ? (hit1 ? sc.match : sc.mismatch) + B : Hinit;
                        const int h2 = real2 ? (hit2 ? sc.match : sc.mismatch) + B : Hinit;
                        HY[k] = pack16(h1, h2);
                        Gy[k] = Bpk;
                        C1Y[k] = (real1 && (hit1 ? dM : dX)) ? bitsel(t2, LC[k], 0x00010000u) : 0u;
                        C2Y[k] = (real2 && (hit2 ? dM : dX)) ? bitsel(t2, LC[k], 0x00000001u) : 0u;
                    }
                    HoA = HY[K - 1]; GoA = Bpk; c1oA = C1Y[K - 1]; c2oA = C2Y[K - 1];
                } else {
                    sets_row<K, GEC, DC>(HX, HY, Gy, C1X, C1Y, C2X, C2Y, HV, XAc, LC, mA, Dmul, GOc, GEpk,
                                     hprev, ginA, c1prev, c2prev, c1inA, c2inA, HoA, GoA, c1oA, c2oA);
                }
                if (store) {
                    if (WIN) {
                        __stcg(&bbuf[2 * iA], make_int4((int)HoA, (int)GoA, (int)c1oA, (int)c2oA));
                        __stcg(&bbuf[2 * iA + 1], make_int4(off1, off2, 0, 0));
                    } else __stcg(&bbuf[iA], make_int4((int)HoA, (int)GoA, (int)c1oA, (int)c2oA));
                }
                if (last_pass) {
                    if (WIN) {
                        const int t1 = lo16(HoA) + off1, t2 = hi16(HoA) + off2;
                        if (t1 > colB1) { colB1 = t1; colI1 = iA; colC1 = c1oA; }
                        if (t2 > colB2) { colB2 = t2; colI2 = iA; colC2 = c2oA; }
                    } else {
                        bool ghi, glo;
                        colBestPk = vibmax_s16x2(colBestPk, HoA, ghi, glo);
                        if (!glo) { colI1 = iA; colC1 = c1oA; }
                        if (!ghi) { colI2 = iA; colC2 = c2oA; }
                    }
                }
                if (iA + 1 < n) {
                    sets_row<K, GEC, DC>(HY, HX, Gy, C1Y, C1X, C2Y, C2X, HV, XAc, LC, mB, Dmul, GOc, GEpk,
                                     hinA, ginB, c1inA, c2inA, c1inB, c2inB, HoB, GoB, c1oB, c2oB);
                    if (store) {
                        if (WIN) {
                            __stcg(&bbuf[2 * iA + 2], make_int4((int)HoB, (int)GoB, (int)c1oB, (int)c2oB));
                            __stcg(&bbuf[2 * iA + 3], make_int4(off1, off2, 0, 0));
                        } else __stcg(&bbuf[iA + 1], make_int4((int)HoB, (int)GoB, (int)c1oB, (int)c2oB));
                    }
                    if (last_pass) {
                        if (WIN) {
                            const int t1 = lo16(HoB) + off1, t2 = hi16(HoB) + off2;
                            if (t1 > colB1) { colB1 = t1; colI1 = iA + 1; colC1 = c1oB; }
                            if (t2 > colB2) { colB2 = t2; colI2 = iA + 1; colC2 = c2oB; }
                        } else {
                            bool ghi, glo;
                            colBestPk = vibmax_s16x2(colBestPk, HoB, ghi, glo);
                            if (!glo) { colI1 = iA + 1; colC1 = c1oB; }
                            if (!ghi) { colI2 = iA + 1; colC2 = c2oB; }
                        }
                    }
                }
                hprev = hinB; c1prev = c1inB; c2prev = c2inB;
                if (WIN) {
                    const uint32_t edge = (iA + 1 < n) ? HoB : HoA;
                    const int v1 = lo16(edge) - B, v2 = hi16(edge) - B;
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
        }
        __syncwarp();
        int bv1 = INT_MIN, bj1 = INT_MAX, bv2 = INT_MIN, bj2 = INT_MAX;
        uint32_t bc1 = 0, bc2 = 0;
        const bool in_y = (n & 1);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j1 = s0 + k - pad1, j2 = s0 + k - pad2;
            const uint32_t hk = in_y ? HY[k] : HX[k];
            const uint32_t ck1 = in_y ? C1Y[k] : C1X[k], ck2 = in_y ? C2Y[k] : C2X[k];
            const int h1 = lo16(hk) + (WIN ? off1 : -B), h2 = hi16(hk) + (WIN ? off2 : -B);
            if (j1 >= 0 && h1 > bv1) { bv1 = h1; bj1 = j1; bc1 = ck1; }
            if (j2 >= 0 && h2 > bv2) { bv2 = h2; bj2 = j2; bc2 = ck2; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int ov1 = __shfl_xor_sync(FULL_MASK, bv1, d), oj1 = __shfl_xor_sync(FULL_MASK, bj1, d);
            const uint32_t oc1 = __shfl_xor_sync(FULL_MASK, bc1, d);
            if (ov1 > bv1 || (ov1 == bv1 && oj1 < bj1)) { bv1 = ov1; bj1 = oj1; bc1 = oc1; }
            const int ov2 = __shfl_xor_sync(FULL_MASK, bv2, d), oj2 = __shfl_xor_sync(FULL_MASK, bj2, d);
            const uint32_t oc2 = __shfl_xor_sync(FULL_MASK, bc2, d);
            if (ov2 > bv2 || (ov2 == bv2 && oj2 < bj2)) { bv2 = ov2; bj2 = oj2; bc2 = oc2; }
        }
        if (bj1 != INT_MAX && bv1 > rowBest1) { rowBest1 = bv1; rowJ1 = bj1; rowC1 = bc1; }
        if (bj2 != INT_MAX && bv2 > rowBest2) { rowBest2 = bv2; rowJ2 = bj2; rowC2 = bc2; }
    }
    colBestPk = __shfl_sync(FULL_MASK, colBestPk, 31);
    colB1 = __shfl_sync(FULL_MASK, colB1, 31); colB2 = __shfl_sync(FULL_MASK, colB2, 31);
    colI1 = __shfl_sync(FULL_MASK, colI1, 31); colC1 = __shfl_sync(FULL_MASK, colC1, 31);
    colI2 = __shfl_sync(FULL_MASK, colI2, 31); colC2 = __shfl_sync(FULL_MASK, colC2, 31);
    if (lane == 0) {
        const int colBest1 = WIN ? colB1 : lo16(colBestPk) - B, colBest2 = WIN ? colB2 : hi16(colBestPk) - B;
        pa_pair_result o;
        if (res1) {     // pair 1 counts (columns << 16 | matches)
            const uint32_t c = rowBest1 > colBest1 ? rowC1 : colC1;
            if (rowBest1 > colBest1) { o.score = rowBest1; o.end_i = n - 1; o.end_j = rowJ1; }
            else                     { o.score = colBest1; o.end_i = colI1; o.end_j = m1 - 1; }
            o.len = c >> 16; o.dist = (c >> 16) - (c & 0xffffu);
            *res1 = o;
        }
        if (res2) {     // pair 2 counts (matches << 16 | columns)
            const uint32_t c = rowBest2 > colBest2 ? rowC2 : colC2;
            if (rowBest2 > colBest2) { o.score = rowBest2; o.end_i = n - 1; o.end_j = rowJ2; }
            else                     { o.score = colBest2; o.end_i = colI2; o.end_j = m2 - 1; }
            o.len = c & 0xffffu; o.dist = (c & 0xffffu) - (c >> 16);
            *res2 = o;
        }
    }
}

// The items of the triangle range with an ambiguous (not plain A/C/G/T) sequence; same work items as
// pa_warp_duo_kernel<0>, which leaves exactly these to this launch (amb_launch).  setok[s]: sequence s has no gap
// character.  Items this kernel cannot take (a gap character, too long without the window) were deferred by the
// plain launch.
// GEC / DC: compile-time gap extension / match - mismatch (the host launches <-1, 12> for pairalign's scoring), 0 = run time.
template <int GEC = 0, int DC = 0>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
pa_warp_sets_kernel(const SeqStore S, const Scoring sc, const uint64_t first, const uint4 *items, const uint64_t n_items,
                    const uint32_t max_len16, unsigned long long *work_counter, int4 *bbuf_all, const uint32_t bbuf_rows,
                    pa_pair_result *out, const int win_ok) {
    __shared__ __align__(16) uint32_t stage[WARPS_PER_CTA][3][STAGE_WORDS];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * WARPS_PER_CTA + wib;
    int4 *bbuf = bbuf_all + (size_t)gw * bbuf_rows;
    const uint32_t N = S.n_seq;

    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1ull);
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w >= n_items) break;
        const uint4 it = __ldg(&items[w]);
        const uint32_t a = it.x, b1 = it.y;
        uint32_t b2 = it.z;
        const uint64_t q0 = tri_row_start(a, N) - a - 1;
        const uint64_t q1 = q0 + b1;
        bool use1 = true;
        bool use2 = (b2 != NO_PARTNER);
        if (!use2) b2 = b1;
        const uint64_t q2 = q0 + b2;
        const int n = (int)S.len[a], m1 = (int)S.len[b1], m2 = (int)S.len[b2];
        const uint32_t lim = win_ok ? 0xffffffffu : max_len16;
        const bool okx = S.fastok[a] && n > 0 && (uint32_t)n <= lim;
        use1 = use1 && okx && S.fastok[b1] && m1 > 0 && (uint32_t)m1 <= lim;
        use2 = use2 && okx && S.fastok[b2] && m2 > 0 && (uint32_t)m2 <= lim;
        if (!use1 && !use2) continue;
        const uint32_t y1 = use1 ? b1 : b2, y2 = use2 ? b2 : b1;
        const int my1 = use1 ? m1 : m2, my2 = use2 ? m2 : m1;
        if (S.pure[a] && S.pure[y1] && S.pure[y2]) continue;      // the plain launch has it
        __syncwarp();
        const uint32_t *xs = stage_seq(S.p4 + S.off4[a], (uint32_t)(n + 7) >> 3, stage[wib][0], lane);
        const uint32_t *ys1 = stage_seq(S.p4 + S.off4[y1], (uint32_t)(my1 + 7) >> 3, stage[wib][1], lane);
        const uint32_t *ys2 = stage_seq(S.p4 + S.off4[y2], (uint32_t)(my2 + 7) >> 3, stage[wib][2], lane);
        __syncwarp();
        pa_pair_result *r1 = use1 ? &out[q1 - first] : nullptr, *r2 = use2 ? &out[q2 - first] : nullptr;
        const bool too_long = (uint32_t)n > max_len16 || (uint32_t)my1 > max_len16 || (uint32_t)my2 > max_len16;
        if (too_long) align_warp_sets<12, true, GEC, DC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 2, r1, r2, lane);
        else align_warp_sets<12, false, GEC, DC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 1, r1, r2, lane);
    }
}

}  // namespace pa
