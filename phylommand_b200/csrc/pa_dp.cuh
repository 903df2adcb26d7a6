// pa_dp.cuh -- device code of the pairalign hot path for sm_100a.
//
// What the reference computes per pair (all citations relative to the
// reference checkout): seqpair::align() fills three n x m int matrices
// (src/seqpair.cpp:95-131), picks the best cell of the last column / last row
// (src/seqpair.cpp:134-143), walks back re-deciding the move at every cell from
// the three values AT that cell (src/seqpair.cpp:159-178), and then
// hamming_distance()/similarity() count columns over the aligned strings
// (src/seqpair.cpp:238-274).
//
// How it is computed here: nothing is stored.  Because the move at (i,j)
// depends only on (H,Gy,Gx) at (i,j), every cell has exactly one predecessor
// and the pair (compared columns, mismatching columns) of the reference's
// traceback path ending in (i,j) obeys the same dependency pattern as the
// scores.  One warp owns one pair; each lane keeps a strip of K columns in
// registers and the warp sweeps the rows as a skewed wavefront (lane l works
// on row t-l at step t), passing the strip's right edge to lane l+1 with
// __shfl_up_sync.  Sequences wider than 32*K columns take several passes; the
// right edge of a pass goes through a per-warp scratch row in global memory
// (16 B per row, L2 resident).
//
// Recurrence (0-based, i over x = rows, j over y = columns):
//   H (i,j) = max(H(i-1,j-1), Gy(i-1,j), Gx(i,j-1)) + s(i,j)
//   Gy(i,j) = max(H(i-1,j-1)+GO, Gy(i-1,j)+GE)
//   Gx(i,j) = max(H(i-1,j-1)+GO, Gx(i,j-1)+GE)
//   first row / column: H = s, Gy = Gx = 0            (src/seqpair.cpp:103-120)
//   move(i,j) = D if H>=Gy && H>=Gx, else U if Gy>=Gx, else L
//   cnt(i,j)  = D ? cnt(i-1,j-1)+[both non-gap](1 column, mismatch?) :
//               U ? cnt(i-1,j) : cnt(i,j-1);   cnt = 0 outside the matrix
// cnt packs (columns << 16 | mismatches).
#pragma once

#ifndef PA_DUO_ADD_VARIANT
#define PA_DUO_ADD_VARIANT 0    // measured on B200 (profiles/r02_add_placement.txt): 0 plain adds 2534 GCUPS, 1 H+GO as IMAD with a register
#endif                          // multiplier 2461, 2 all three adds so 2394, 3 counter adds as IMAD with an immediate multiplier 2462


#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

#include "../../include/pairalign_b200.h"

namespace pa {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr int WARPS_PER_CTA = 4;
constexpr int STAGE_WORDS = 256;     // per sequence per warp: 1 KB = 4096 2-bit or 2048 4-bit codes

struct SeqStore {
    const uint32_t *p2;      // 2 bit/base (A0 G1 C2 T3), 16 bases per word, every sequence 16-byte aligned
    const uint32_t *p4;      // 4 bit/base IUPAC sets, 8 bases per word, every sequence 16-byte aligned
    const uint32_t *off2;    // word offset of sequence s in p2
    const uint32_t *off4;    // word offset of sequence s in p4
    const uint32_t *len;     // encoded length
    const uint8_t  *pure;    // 1 if the sequence holds only A/C/G/T
    const uint8_t  *fastok;  // 1 if the sequence holds no gap character: plain, or IUPAC codes for the 4-bit-set s16x2 kernel
    uint32_t n_seq;
};

// bias16: see duo_row (s16x2 kernels only).  one, one2: both 1 -- multipliers that turn an add into an IMAD the compiler cannot
// move to the ALU pipe; two of them so that ptxas can keep one in a uniform register (IMAD R, R, UR, R) while the other
// sits in a vector register next to a uniform addend (IMAD R, R, R, UR): either way two vector-register operands.
struct Scoring { int match, mismatch, go, ge; int bias16 = 0; int one = 1; int one2 = 1; };

struct PairSource {
    uint64_t first;          // triangle mode: global index of element 0
    const uint32_t *ia;      // explicit mode (ia != nullptr): pair k = (ia[k], ib[k])
    const uint32_t *ib;
    const uint32_t *idx;     // optional indirection: work item w handles element idx[w]
};

// Row-major upper-triangle index -> (a,b), a<b.  Rows before r hold
// r*(2N-r-1)/2 pairs.
__host__ __device__ inline uint64_t tri_row_start(uint64_t r, uint64_t N) { return r * (2 * N - r - 1) / 2; }

__host__ __device__ inline void tri_pair(uint64_t q, uint32_t N, uint32_t &a, uint32_t &b) {
    const double t = 2.0 * (double)N - 1.0;
    double disc = t * t - 8.0 * (double)q;
    if (disc < 0.0) disc = 0.0;
    long long r = (long long)((t - sqrt(disc)) * 0.5);
    if (r < 0) r = 0;
    if (r > (long long)N - 2) r = (long long)N - 2;
    while (r > 0 && tri_row_start((uint64_t)r, N) > q) --r;
    while (r < (long long)N - 2 && tri_row_start((uint64_t)r + 1, N) <= q) ++r;
    a = (uint32_t)r;
    b = (uint32_t)(r + 1 + (long long)(q - tri_row_start((uint64_t)r, N)));
}

__device__ __forceinline__ uint32_t fetch2(const uint32_t *w, int j) { return (w[j >> 4] >> ((j & 15) * 2)) & 3u; }
__device__ __forceinline__ uint32_t fetch4(const uint32_t *w, int j) { return (w[j >> 3] >> ((j & 7) * 4)) & 15u; }

__device__ __forceinline__ int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }

// prmt.b32 in its default mode: selector nibble bits 0-2 pick one of the 8
// source bytes (a: 0-3, b: 4-7); bit 3 replicates that byte's sign bit instead.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// compile-time switches handed to generic lambdas
struct FlagFalse { static constexpr bool value = false; };
struct FlagTrue { static constexpr bool value = true; };

// ---------------------------------------------------------------------------
// One pair on one warp.  xs/ys: packed codes (2-bit when !GENERAL, 4-bit sets
// when GENERAL), generic pointers (shared staging or global).
//
// Column layout: the m columns are RIGHT-aligned in P*32*K slots, so the last
// column is always the last slot of lane 31 in the last pass and its per-row
// values are exactly what that lane hands on anyway.  The padL = P*32*K-m slots
// before column 0 are neutral pad columns:
//   !GENERAL: pad cells score 0 and count nothing; with the virtual values
//     H=-GO, Gy=Gx=0 that is a fixed point of the recurrence, and -GO is what
//     column 0 / row 0 need to see so that (with GO added to their scores)
//     H = s and Gy = Gx = 0 come out of the ordinary recurrence: no per-cell
//     boundary test.  Scores come from two byte tables indexed with PRMT.
//   GENERAL: explicit flags per column (any parameters, gaps -> INT_MIN with
//     32-bit wrap-around exactly as the reference binary behaves).
// ---------------------------------------------------------------------------
template <int K, bool GENERAL, bool DIRS = false>
__device__ __forceinline__ void align_warp(const uint32_t *xs, const int n, const uint32_t *ys, const int m,
                                           const Scoring sc, int4 *bbuf, pa_pair_result *res, const int lane,
                                           uint8_t *dirs = nullptr) {
    constexpr int W = 32 * K;
    const int P = (m + W - 1) / W;
    const int padL = P * W - m;
    const int Hinit = GENERAL ? 0 : -sc.go;

    int rowBest = INT_MIN, rowJ = 0;
    uint32_t rowC = 0;
    int colBest = INT_MIN, colI = n - 1;
    uint32_t colC = 0;

    // fast-path constants
    uint32_t y0 = 0, baseN = 0, baseZ = 0, dN = 0, dZ = 0, c0m = 0, c0x = 0;
    if (!GENERAL) {
        y0 = fetch2(ys, 0);
        const uint32_t Mn = (uint32_t)sc.match & 0xffu, Xn = (uint32_t)sc.mismatch & 0xffu;
        const uint32_t Mz = (uint32_t)(sc.match + sc.go) & 0xffu, Xz = (uint32_t)(sc.mismatch + sc.go) & 0xffu;
        baseN = Xn * 0x01010101u; dN = Mn ^ Xn;
        baseZ = Xz * 0x01010101u; dZ = Mz ^ Xz;
        c0m = Mz; c0x = Xz;
    }

    for (int p = 0; p < P; ++p) {
        const int j0 = p * W + lane * K - padL;   // column of this lane's k = 0 (negative: pad)
        int H[K], Gy[K];
        uint32_t C[K];
        uint32_t selS[K];   // !GENERAL: PRMT selector of the score byte;   GENERAL: 4-bit set of the column
        uint32_t selI[K];   // !GENERAL: PRMT selector of the count increment; GENERAL: 0 normal, 1 column 0, 2 pad
        // GENERAL: what a cell of this column scores / counts when the two sets do not intersect -- the mismatch score
        // and (1 column, 1 mismatch), or INT_MIN and nothing when the column is a gap character.  The row has the same
        // pair of constants; a cell takes the smaller of the two (src/seqpair.cpp:192-193: either side empty -> INT_MIN).
        int cXG[GENERAL ? K : 1];
        uint32_t cInc[GENERAL ? K : 1];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = j0 + k;
            H[k] = Hinit; Gy[k] = 0; C[k] = 0;
            if (!GENERAL) {
                if (j < 0)       { selS[k] = 0xddd5u; selI[k] = 0x5555u; }
                else if (j == 0) { selS[k] = 0xccc4u; selI[k] = 0x5654u; }
                else { const uint32_t c = fetch2(ys, j); selS[k] = 0x8880u + 0x1111u * c; selI[k] = 0x5650u | c; }
            } else {
                if (j < 0)       { selS[k] = 0; selI[k] = 2; }
                else             { selS[k] = fetch4(ys, j); selI[k] = (j == 0) ? 1u : 0u; }
                cXG[k] = selS[k] ? sc.mismatch : INT_MIN;
                cInc[k] = selS[k] ? 0x10001u : 0u;
            }
        }
        int hprev = Hinit;
        uint32_t cprev = 0;
        int Hout = Hinit, Gxout = 0;
        uint32_t cout = 0;
        int4 nxt = make_int4(Hinit, 0, 0, 0);
        if (p > 0 && lane == 0) nxt = __ldcg(&bbuf[0]);
        uint32_t xcur = 0, xprev = 0;
        if (lane < n) xcur = GENERAL ? fetch4(xs, lane) : fetch2(xs, lane);

        const int T = n + 31;
        for (int t = 0; t < T; ++t) {
            const int r = t & 31;
            if (r == 0 && t > 0) {
                xprev = xcur;
                const int ii = t + lane;
                xcur = 0;
                if (ii < n) xcur = GENERAL ? fetch4(xs, ii) : fetch2(xs, ii);
            }
            const uint32_t xv = (lane <= r) ? xcur : xprev;
            const uint32_t xi = __shfl_sync(FULL_MASK, xv, (r - lane) & 31);
            int hin = __shfl_up_sync(FULL_MASK, Hout, 1);
            int gin = __shfl_up_sync(FULL_MASK, Gxout, 1);
            uint32_t cin = __shfl_up_sync(FULL_MASK, cout, 1);
            if (lane == 0) {
                hin = nxt.x; gin = nxt.y; cin = (uint32_t)nxt.z;
                if (p > 0 && t + 1 < n) nxt = __ldcg(&bbuf[t + 1]);
            }
            const int i = t - lane;
            if (i >= 0 && i < n) {
                int Hd = hprev, Gl = gin;
                uint32_t cd = cprev, cl = cin;
                if (!GENERAL) {
                    const uint32_t sh = xi * 8u;
                    const bool z = (i == 0);
                    const uint32_t Rlo = (z ? baseZ : baseN) ^ ((z ? dZ : dN) << sh);
                    const uint32_t Rhi = (xi == y0) ? c0m : c0x;            // byte4: column-0 score, byte5: 0 (pad)
                    const uint32_t Mlo = 0x01010101u ^ (1u << sh);           // mismatch flags per base code
                    const uint32_t Mhi = 0x00010000u | (xi != y0 ? 1u : 0u); // byte4: col-0 flag, byte5: 0, byte6: 1
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int s = (int)prmt(Rlo, Rhi, selS[k]);
                        const uint32_t inc = prmt(Mlo, Mhi, selI[k]);
                        const int Gu = Gy[k];
                        const uint32_t cu = C[k];
                        const int h = __vimax3_s32(Hd, Gu, Gl) + s;
                        const int o = Hd + sc.go;
                        const int gy = __viaddmax_s32(Gu, sc.ge, o);
                        const int gx = __viaddmax_s32(Gl, sc.ge, o);
                        const bool pD = (h >= gy) && (h >= gx);
                        const bool pU = (gy >= gx);
                        const uint32_t cdi = cd + inc;
                        const uint32_t c = pD ? cdi : (pU ? cu : cl);
                        Hd = H[k]; cd = cu;
                        H[k] = h; Gy[k] = gy; C[k] = c;
                        Gl = gx; cl = c;
                    }
                } else {
                    // xi == 0: a gap character in x -- every cell of the row scores INT_MIN and counts nothing
                    const int rXG = xi ? sc.mismatch : INT_MIN;
                    const uint32_t rInc = xi ? 0x10001u : 0u;
                    const bool row0 = (i == 0);
                    uint32_t mv = 0;
                    // EDGE: this pass holds pad slots and / or column 0 (pass 0 only); the other passes carry no flags
                    auto cells = [&](auto edge_c) {
                        constexpr bool EDGE = decltype(edge_c)::value;
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const uint32_t ym = selS[k];
                            const bool hit = (xi & ym) != 0;
                            const int s = hit ? sc.match : min(rXG, cXG[k]);
                            const uint32_t inc = hit ? 0x10000u : min(rInc, cInc[k]);
                            const int Gu = Gy[k];
                            const uint32_t cu = C[k];
                            int h = wadd(__vimax3_s32(Hd, Gu, Gl), s);
                            const int o = wadd(Hd, sc.go);
                            int gy = max(o, wadd(Gu, sc.ge));
                            int gx = max(o, wadd(Gl, sc.ge));
                            if (EDGE) {
                                const uint32_t fl = selI[k];
                                if (fl != 0) { gy = 0; gx = 0; }
                                if (fl == 2) h = 0;
                            }
                            const bool pD = (h >= gy) && (h >= gx);
                            const bool pU = (gy >= gx);
                            uint32_t c = pD ? cd + inc : (pU ? cu : cl);
                            if (EDGE && selI[k] == 2) c = 0;
                            if (DIRS) mv |= (pD ? 0u : (pU ? 1u : 3u)) << (2 * k);      // bit 0: not diagonal, bit 1: left rather than up
                            Hd = H[k]; cd = cu;
                            H[k] = h; Gy[k] = gy; C[k] = c;
                            Gl = gx; cl = c;
                        }
                    };
                    if (row0) {
                        // first row: H = s, Gy = Gx = 0 (src/seqpair.cpp:103-107), so the move is D iff s >= 0 and the
                        // counters start there; no recurrence, and the other rows need no first-row test
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            const bool hit = (xi & selS[k]) != 0;
                            const int s = hit ? sc.match : min(rXG, cXG[k]);
                            const uint32_t inc = hit ? 0x10000u : min(rInc, cInc[k]);
                            const int h = (selI[k] == 2) ? 0 : s;
                            const bool pD = (h >= 0);
                            const uint32_t c = pD ? inc : 0u;
                            if (DIRS) mv |= (pD ? 0u : 1u) << (2 * k);
                            H[k] = h; Gy[k] = 0; C[k] = c;
                            cl = c;
                        }
                        Gl = 0;
                    } else if (p == 0) cells(FlagTrue{});
                    else cells(FlagFalse{});
                    if (DIRS) {   // 2 bits per slot, row-major, K/4 bytes per lane (K = 8: one 16-bit store)
                        static_assert(!DIRS || K == 8, "the move store assumes K = 8");
                        reinterpret_cast<uint16_t *>(dirs)[(size_t)i * ((size_t)P * 32) + (size_t)p * 32 + lane] = (uint16_t)mv;
                    }
                }
                hprev = hin; cprev = cin;
                Hout = H[K - 1]; Gxout = Gl; cout = cl;
                if (lane == 31) {
                    if (p < P - 1) {
                        __stcg(&bbuf[i], make_int4(Hout, Gxout, (int)cout, 0));
                    } else {
                        // last column, rows ascending, strict >  (src/seqpair.cpp:137-139)
                        if (Hout > colBest) { colBest = Hout; colI = i; colC = cout; }
                        // nothing above INT_MIN anywhere: the reference stays on its initial (n-1, m-1)  (:132-133)
                        if (i == n - 1 && colBest == INT_MIN) colC = cout;
                    }
                }
            }
        }
        __syncwarp();
        // last row of this pass: columns ascending, strict >  (src/seqpair.cpp:140-142)
        int bv = INT_MIN, bj = INT_MAX;
        uint32_t bc = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = j0 + k;
            if (j >= 0 && H[k] > bv) { bv = H[k]; bj = j; bc = C[k]; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int ov = __shfl_xor_sync(FULL_MASK, bv, d);
            const int oj = __shfl_xor_sync(FULL_MASK, bj, d);
            const uint32_t oc = __shfl_xor_sync(FULL_MASK, bc, d);
            if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; bc = oc; }
        }
        if (bj != INT_MAX && bv > rowBest) { rowBest = bv; rowJ = bj; rowC = bc; }
    }
    colBest = __shfl_sync(FULL_MASK, colBest, 31);
    colI = __shfl_sync(FULL_MASK, colI, 31);
    colC = __shfl_sync(FULL_MASK, colC, 31);
    if (lane == 0) {
        pa_pair_result o;
        if (rowBest > colBest) { o.score = rowBest; o.end_i = n - 1; o.end_j = rowJ; o.dist = rowC & 0xffffu; o.len = rowC >> 16; }
        else                   { o.score = colBest; o.end_i = colI;  o.end_j = m - 1; o.dist = colC & 0xffffu; o.len = colC >> 16; }
        *res = o;
    }
}

// ---------------------------------------------------------------------------
// Two pairs (x, y1) and (x, y2) that share the row sequence on one warp, packed
// as signed 16-bit halves of every score register (low half = pair 1, high half
// = pair 2) and computed with the s16x2 DPX instructions of sm_100a
// (VIMNMX3.S16x2, VIADDMNMX.S16x2, VIADD.16x2, and VIMNMX.S16x2 with its two
// predicate outputs for the move).  The (columns, mismatches) counters stay 32
// bit per pair.  Valid while every score fits int16 (host checks:
// match*len <= 32000 and |gap_ext|*len + |mismatch| + |gap_open| <= 32000).
//
// Both column sets are right-aligned in the same P*32*K slots, so pad columns
// (a fixed point of the recurrence, see align_warp) absorb the length
// difference and the last column of BOTH pairs is lane 31's last slot.
// Score table bytes: 0-3 base codes, 4 column 0 of pair 1, 5 zero (pad),
// 6 column 0 of pair 2.  Increment table bytes: 0-3 mismatch flag per base code,
// 4 column-0 flag pair 1, 5 zero, 6 one, 7 column-0 flag pair 2.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int lo16(uint32_t v) { return (int)(short)(v & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t v) { return (int)v >> 16; }
__device__ __forceinline__ uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

// BIAS.  Every packed state is stored as true value + B with B < 0 chosen so that both halves are always
// NEGATIVE (B = sc.bias16 from the host for plain 16-bit scores, WIN_BIAS inside the floating window).  The
// recurrence only adds constants and takes maxima, so a common bias passes straight through it.  What it buys:
// H + GO, one of the five packed additions per cell, becomes a plain 32-bit add -- with the low half negative,
// low + (GO & 0xffff) carries into the high half on every single cell (for GO < 0), so adding the constant
// GOc = (GO - 1) << 16 | (GO & 0xffff) is exact for both halves.  ptxas issues that add as IMAD.IADD on the FMA
// pipe; the ALU pipe, which bounds this kernel, has 24 instructions fewer per step of 24 cells (K = 12).
//
// One row of the packed recurrence over this lane's K columns.  Reads the
// previous row from (Hs, C1s, C2s), writes this row to (Hd, C1d, C2d); Gy is
// updated in place.  Source and destination arrays are different registers
// (the caller ping-pongs two rows per step), so nothing has to be copied to
// keep H(i-1,j-1) / cnt(i-1,j-1) alive for the next column.
// max per signed half plus "a >= b" per half: the VIMNMX.S16x2 form with two predicate outputs.
// Same PTX as CUDA's __vibmax_s16x2, but with early-clobber outputs: the toolkit's version lets the
// result share a register with `a` when `a` dies here, and then compares the maximum with itself.
__device__ __forceinline__ uint32_t vibmax_s16x2(const uint32_t a, const uint32_t b, bool &ge_hi, bool &ge_lo) {
    uint32_t val, phi, plo;
    asm("{.reg .pred pu, pv;\n\t"
        ".reg .s16 rs0, rs1, rs2, rs3;\n\t"
        "max.s16x2 %0, %3, %4;\n\t"
        "mov.b32 {rs0, rs1}, %0;\n\t"
        "mov.b32 {rs2, rs3}, %3;\n\t"
        "setp.eq.s16 pv, rs0, rs2;\n\t"
        "setp.eq.s16 pu, rs1, rs3;\n\t"
        "selp.b32 %1, 1, 0, pu;\n\t"
        "selp.b32 %2, 1, 0, pv;}"
        : "=&r"(val), "=&r"(phi), "=&r"(plo) : "r"(a), "r"(b));
    ge_hi = (phi != 0);
    ge_lo = (plo != 0);
    return val;
}

// GEC != 0: the gap-extension penalty is the compile-time constant GEC and travels as an immediate operand of
// VIADDMNMX.S16x2 -- two register reads fewer per cell (measured: +2 % on config 2).
template <int K, int GEC = 0>
__device__ __forceinline__ void duo_row(const uint32_t (&Hs)[K], uint32_t (&Hd)[K], uint32_t (&Gy)[K],
                                        const uint32_t (&C1s)[K], uint32_t (&C1d)[K],
                                        const uint32_t (&C2s)[K], uint32_t (&C2d)[K],
                                        const uint32_t (&selS)[K], const uint32_t (&selI1)[K], const uint32_t (&selI2)[K],
                                        const uint32_t Rlo, const uint32_t Rhi, const uint32_t Mlo, const uint32_t Mhi,
                                        const uint32_t GOc, const uint32_t GEpk, const uint32_t one, const uint32_t one2,
                                        uint32_t hdiag, uint32_t Gl, uint32_t cd1, uint32_t cd2, uint32_t cl1, uint32_t cl2,
                                        uint32_t &Hout, uint32_t &Gxout, uint32_t &c1out, uint32_t &c2out) {
    const uint32_t GE2 = GEC ? ((uint32_t)GEC & 0xffffu) * 0x10001u : GEpk;
    uint32_t Hdg = hdiag;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t s = prmt(Rlo, Rhi, selS[k]);
        const uint32_t inc1 = prmt(Mlo, Mhi, selI1[k]);
        const uint32_t inc2 = prmt(Mlo, Mhi, selI2[k]);
        const uint32_t Gu = Gy[k];
        const uint32_t cu1 = C1s[k], cu2 = C2s[k];
        const uint32_t h = __vadd2(__vimax3_s16x2(Hdg, Gu, Gl), s);
#if PA_DUO_ADD_VARIANT == 0
        const uint32_t o = Hdg + GOc;
#else
        const uint32_t o = Hdg * one + GOc;                          // IMAD: both halves + GO (see above); one = 1, opaque to ptxas,
                                                                     // which otherwise moves this add to the ALU pipe when it likes
#endif
        const uint32_t gy = __viaddmax_s16x2(Gu, GE2, o);
        const uint32_t gx = __viaddmax_s16x2(Gl, GE2, o);
        bool pUhi, pUlo, pDhi, pDlo;
        const uint32_t g = vibmax_s16x2(gy, gx, pUhi, pUlo);         // gy >= gx
        (void)vibmax_s16x2(h, g, pDhi, pDlo);                        // h >= max(gy, gx)
#if PA_DUO_ADD_VARIANT == 2
        const uint32_t cdi1 = cd1 * one2 + inc1, cdi2 = cd2 * one2 + inc2;
#elif PA_DUO_ADD_VARIANT == 3
        const uint32_t cdi1 = inc1 * 3u + cd1, cdi2 = inc2 * 3u + cd2;       // counters hold 3 x (columns, mismatches): IMAD with an immediate
#else
        const uint32_t cdi1 = cd1 + inc1, cdi2 = cd2 + inc2;
#endif
        const uint32_t c1 = pDlo ? cdi1 : (pUlo ? cu1 : cl1);
        const uint32_t c2 = pDhi ? cdi2 : (pUhi ? cu2 : cl2);
        Hdg = Hs[k]; cd1 = cu1; cd2 = cu2;
        Hd[k] = h; Gy[k] = gy; C1d[k] = c1; C2d[k] = c2;
        Gl = gx; cl1 = c1; cl2 = c2;
    }
    Hout = Hd[K - 1]; Gxout = Gl; c1out = cl1; c2out = cl2;
}

// tab: this warp's 8-entry shared table, tab[z*4 + x] = (Rlo, Rhi, Mlo, Mhi) for row code x, z = 1 for
// row 0 (scores carry +GO there).  vrow: index of the spare scratch row that holds the virtual column
// left of column 0, so that pass 0 and later passes feed lane 0 through the same loads.
// WIN (floating window): scores of long pairs do not fit 16 bits, but the values a lane holds at any moment (K
// columns x 2 rows, plus what its neighbour hands over) lie within a few hundred of each other -- every state is at
// most one gap opening away from its diagonal neighbour.  Each lane therefore keeps its halves relative to its own
// 32-bit offsets (off1, off2: true value = stored + offset), converts what it receives from the left lane's frame
// into its own (one packed add per value and step), and re-bases by WIN_Q whenever its right-edge value leaves
// +-WIN_T.  The recurrence itself is untouched: it only ever combines values of one frame.  The edge rows handed to
// the next pass carry their offsets in a second int4 (bbuf rows 2i, 2i+1), maxima are compared as true 32-bit values.
constexpr int WIN_T = 8192, WIN_Q = 4096;
constexpr int WIN_PREFETCH = 8;      // steps between the L2 prefetch of an edge row and its use (a step is ~1 us of the warp)
constexpr int WIN_BIAS = -16384;     // centre of the window in stored terms: values stay in about [-29000, -3800]

template <int K, bool WIN = false, int GEC = 0>
__device__ __forceinline__ void align_warp_duo(const uint32_t *xs, const int n, const uint32_t *ys1, const int m1,
                                               const uint32_t *ys2, const int m2, const Scoring sc, int4 *bbuf,
                                               const uint32_t vrow, int4 *tab,
                                               pa_pair_result *res1, pa_pair_result *res2, const int lane) {
    constexpr int W = 32 * K;
    const int mmax = m1 > m2 ? m1 : m2;
    const int P = (mmax + W - 1) / W;
    const int pad1 = P * W - m1, pad2 = P * W - m2;
    const int B = WIN ? WIN_BIAS : sc.bias16;      // stored = true + B, negative in both halves (duo_row)
    const uint32_t Bpk = pack16(B, B);
    const int Hinit = -sc.go + B;
    const uint32_t HinitPk = pack16(Hinit, Hinit);
    const uint32_t GOc = sc.go ? pack16(sc.go, sc.go - 1) : 0u, GEpk = pack16(sc.ge, sc.ge);
    const uint32_t one = (uint32_t)sc.one, one2 = (uint32_t)sc.one2;

    int rowBest1 = INT_MIN, rowJ1 = 0, rowBest2 = INT_MIN, rowJ2 = 0;
    uint32_t rowC1 = 0, rowC2 = 0;
    // last-column maxima of both pairs, packed like the scores; -32768 stands for "nothing yet"
    uint32_t colBestPk = 0x80008000u;
    int colB1 = INT_MIN, colB2 = INT_MIN;      // WIN: the same maxima as true 32-bit values
    int colI1 = n - 1, colI2 = n - 1;
    uint32_t colC1 = 0, colC2 = 0;

    const uint32_t y10 = fetch2(ys1, 0), y20 = fetch2(ys2, 0);
    {   // score / increment tables of this (x, y1, y2): one entry per row code and row-0 flag
        if (lane < 8) {
            const uint32_t xi = lane & 3u, z = lane >> 2;
            const int adj = z ? sc.go : 0;
            const uint32_t Mb = (uint32_t)(sc.match + adj) & 0xffu, Xb = (uint32_t)(sc.mismatch + adj) & 0xffu;
            const uint32_t Mz = (uint32_t)(sc.match + sc.go) & 0xffu, Xz = (uint32_t)(sc.mismatch + sc.go) & 0xffu;
            const uint32_t sh = xi * 8u;
            int4 e;
            e.x = (int)((Xb * 0x01010101u) ^ ((Mb ^ Xb) << sh));
            e.y = (int)(((xi == y10) ? Mz : Xz) | (((xi == y20) ? Mz : Xz) << 16));
            e.z = (int)(0x01010101u ^ (1u << sh));
            e.w = (int)(0x00010000u | (xi != y10 ? 1u : 0u) | (xi != y20 ? 0x01000000u : 0u));
            tab[z * 4 + xi] = e;
        }
        if (lane == 0) {
            __stcg(&bbuf[vrow], make_int4((int)HinitPk, (int)Bpk, 0, 0));
            if (WIN) __stcg(&bbuf[vrow + 1], make_int4(-B, -B, 0, 0));
        }
        __syncwarp();
    }
    const int n_steps = ((n + 1) >> 1) + 31;        // two rows per step, lanes one step apart
    const int x_last_word = (n - 1) >> 4;

    for (int p = 0; p < P; ++p) {
        const bool last_pass = (p == P - 1);
        const int s0 = p * W + lane * K;           // slot of this lane's k = 0
        // X: the odd row (2u-1, then 2u+1), Y: the even row 2u
        uint32_t HX[K], HY[K], Gy[K], C1X[K], C1Y[K], C2X[K], C2Y[K], selS[K], selI1[K], selI2[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j1 = s0 + k - pad1, j2 = s0 + k - pad2;
            HX[k] = HinitPk; HY[k] = HinitPk; Gy[k] = Bpk; C1X[k] = 0; C2X[k] = 0; C1Y[k] = 0; C2Y[k] = 0;
            uint32_t c1, c2, i1, i2;
            if (j1 < 0)       { c1 = 5; i1 = 0x5555u; }
            else if (j1 == 0) { c1 = 4; i1 = 0x5654u; }
            else              { c1 = fetch2(ys1, j1); i1 = 0x5650u | c1; }
            if (j2 < 0)       { c2 = 5; i2 = 0x5555u; }
            else if (j2 == 0) { c2 = 6; i2 = 0x5657u; }
            else              { c2 = fetch2(ys2, j2); i2 = 0x5650u | c2; }
            selS[k] = ((8u | c2) << 12) | (c2 << 8) | ((8u | c1) << 4) | c1;
            selI1[k] = i1; selI2[k] = i2;
        }
        // what the left neighbour handed over for the previous odd row: diagonal of the next even row
        uint32_t hprev = HinitPk, c1prev = 0, c2prev = 0;
        uint32_t HoA = HinitPk, GoA = Bpk, c1oA = 0, c2oA = 0, HoB = HinitPk, GoB = Bpk, c1oB = 0, c2oB = 0;
        // lane 0's left edge: rows of the previous pass's right edge, or the virtual row in pass 0
        const int4 *feed = p > 0 ? bbuf : bbuf + vrow;
        const int fmul = p > 0 ? (WIN ? 2 : 1) : 0;
        int4 fA = __ldcg(&feed[0]), fB = __ldcg(&feed[fmul * (n > 1 ? 1 : 0)]);
        int4 fOA = make_int4(0, 0, 0, 0), fOB = fOA;       // WIN: the offsets of the two feed rows
        if (WIN) { fOA = __ldcg(&feed[1]); fOB = __ldcg(&feed[fmul * (n > 1 ? 1 : 0) + 1]); }
        int off1 = -B, off2 = -B;                 // true value = stored + off
        uint32_t xw = xs[min(max(-2 * lane, 0) >> 4, x_last_word)];

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
            if (WIN) {   // from the left lane's frame (lane 0: the frames the two edge rows were stored in) into mine
                int oA1 = __shfl_up_sync(FULL_MASK, off1, 1), oA2 = __shfl_up_sync(FULL_MASK, off2, 1);
                int oB1 = oA1, oB2 = oA2;
                if (lane == 0) { oA1 = fOA.x; oA2 = fOA.y; oB1 = fOB.x; oB2 = fOB.y; }
                const uint32_t dA = pack16(oA1 - off1, oA2 - off2), dB = pack16(oB1 - off1, oB2 - off2);
                hinA = __vadd2(hinA, dA); ginA = __vadd2(ginA, dA);
                hinB = __vadd2(hinB, dB); ginB = __vadd2(ginB, dB);
            }
            // prefetch lane 0's next two rows (every lane issues the same address) and this lane's next x word
            fA = __ldcg(&feed[fmul * min(2 * t + 2, n - 1)]);
            fB = __ldcg(&feed[fmul * min(2 * t + 3, n - 1)]);
            if (WIN) {
                fOA = __ldcg(&feed[fmul * min(2 * t + 2, n - 1) + 1]);
                fOB = __ldcg(&feed[fmul * min(2 * t + 3, n - 1) + 1]);
                // the edge rows of a 30 kb pass (0.9 MB per warp, 1.1 GB per grid) left L2 long before the next pass
                // reads them: fetch the line WIN_PREFETCH steps ahead into L2 so that the loads above find it there
                if (p > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(&feed[2 * min(2 * t + 2 + 2 * WIN_PREFETCH, n - 1)]));
            }
            const uint32_t xi2 = (xw >> ((iA & 15) * 2)) & 15u;
            xw = xs[min(max(iA + 2, 0) >> 4, x_last_word)];
            if (iA >= 0 && iA < n) {
                const bool store = (lane == 31) && !last_pass;
                {   // even row iA: previous row in X, result in Y
                    const int4 T = tab[(iA == 0 ? 4 : 0) + (xi2 & 3u)];
                    duo_row<K, GEC>(HX, HY, Gy, C1X, C1Y, C2X, C2Y, selS, selI1, selI2, (uint32_t)T.x, (uint32_t)T.y,
                               (uint32_t)T.z, (uint32_t)T.w, GOc, GEpk, one, one2,
                               hprev, ginA, c1prev, c2prev, c1inA, c2inA, HoA, GoA, c1oA, c2oA);
                    if (store) {
                        if (WIN) {
                            __stcg(&bbuf[2 * iA], make_int4((int)HoA, (int)GoA, (int)c1oA, (int)c2oA));
                            __stcg(&bbuf[2 * iA + 1], make_int4(off1, off2, 0, 0));
                        } else __stcg(&bbuf[iA], make_int4((int)HoA, (int)GoA, (int)c1oA, (int)c2oA));
                    }
                    if (last_pass) {   // last column, rows ascending, strict >; every lane tracks, lane 31 is read
                        if (WIN) {
                            const int t1 = lo16(HoA) + off1, t2 = hi16(HoA) + off2;
                            if (t1 > colB1) { colB1 = t1; colI1 = iA; colC1 = c1oA; }
                            if (t2 > colB2) { colB2 = t2; colI2 = iA; colC2 = c2oA; }
                        } else {
                            bool ghi, glo;                               // best >= candidate: keep
                            colBestPk = vibmax_s16x2(colBestPk, HoA, ghi, glo);
                            if (!glo) { colI1 = iA; colC1 = c1oA; }
                            if (!ghi) { colI2 = iA; colC2 = c2oA; }
                        }
                    }
                }
                if (iA + 1 < n) {   // odd row iA+1: previous row in Y, result in X
                    const int4 T = tab[xi2 >> 2];
                    duo_row<K, GEC>(HY, HX, Gy, C1Y, C1X, C2Y, C2X, selS, selI1, selI2, (uint32_t)T.x, (uint32_t)T.y,
                               (uint32_t)T.z, (uint32_t)T.w, GOc, GEpk, one, one2,
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
                if (WIN) {   // keep this lane's window centred on its right edge
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
        // last row of this pass (in Y when n is odd, in X when even): columns ascending, strict >
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
        // scores stay far above -32768 (host-checked range), so the sentinel is never a real value
        const int colBest1 = WIN ? colB1 : lo16(colBestPk) - B, colBest2 = WIN ? colB2 : hi16(colBestPk) - B;
        pa_pair_result o;
        if (res1) {
            if (rowBest1 > colBest1) { o.score = rowBest1; o.end_i = n - 1; o.end_j = rowJ1; o.dist = rowC1 & 0xffffu; o.len = rowC1 >> 16; }
            else                     { o.score = colBest1; o.end_i = colI1; o.end_j = m1 - 1; o.dist = colC1 & 0xffffu; o.len = colC1 >> 16; }
            *res1 = o;
        }
        if (res2) {
            if (rowBest2 > colBest2) { o.score = rowBest2; o.end_i = n - 1; o.end_j = rowJ2; o.dist = rowC2 & 0xffffu; o.len = rowC2 >> 16; }
            else                     { o.score = colBest2; o.end_i = colI2; o.end_j = m2 - 1; o.dist = colC2 & 0xffffu; o.len = colC2 >> 16; }
            *res2 = o;
        }
    }
}

// Coalesced 128-bit staging of one packed sequence into this warp's shared
// buffer; returns the pointer to read codes from (global if it does not fit).
__device__ __forceinline__ const uint32_t *stage_seq(const uint32_t *g, uint32_t n_words, uint32_t *sm, int lane) {
    if (n_words > (uint32_t)STAGE_WORDS) return g;
    const uint4 *g4 = reinterpret_cast<const uint4 *>(g);
    uint4 *s4 = reinterpret_cast<uint4 *>(sm);
    const uint32_t n4 = (n_words + 3) >> 2;
    for (uint32_t w = lane; w < n4; w += 32) s4[w] = __ldg(&g4[w]);
    return sm;
}

// Persistent warps pull pair indices from a global counter.
//   !GENERAL: pairs with a non-A/C/G/T sequence are appended to `deferred`
//             and handled by the GENERAL instantiation afterwards.
//   count_dev != nullptr: the number of work items is read from device memory
//             (a list another kernel of the same call has just written).
template <int K, bool GENERAL>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
pa_warp_dp_kernel(const SeqStore S, const Scoring sc, const PairSource src, const uint64_t count_host,
                  const unsigned int *count_dev, unsigned long long *work_counter, int4 *bbuf_all,
                  const uint32_t bbuf_rows, pa_pair_result *out, uint32_t *deferred, unsigned int *n_deferred) {
    __shared__ __align__(16) uint32_t stage[WARPS_PER_CTA][2][STAGE_WORDS];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * WARPS_PER_CTA + wib;
    int4 *bbuf = bbuf_all + (size_t)gw * bbuf_rows;
    // the item count of a follow-up launch is whatever the previous kernel deferred (no host round trip)
    const uint64_t count = count_dev ? (uint64_t)*count_dev : count_host;

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
        if (!GENERAL) {
            if (!(S.pure[a] && S.pure[b])) {
                if (lane == 0) deferred[atomicAdd(n_deferred, 1u)] = (uint32_t)e;
                continue;
            }
        }
        __syncwarp();
        const uint32_t *xs, *ys;
        if (GENERAL) {
            xs = stage_seq(S.p4 + S.off4[a], (uint32_t)(n + 7) >> 3, stage[wib][0], lane);
            ys = stage_seq(S.p4 + S.off4[b], (uint32_t)(m + 7) >> 3, stage[wib][1], lane);
        } else {
            xs = stage_seq(S.p2 + S.off2[a], (uint32_t)(n + 15) >> 4, stage[wib][0], lane);
            ys = stage_seq(S.p2 + S.off2[b], (uint32_t)(m + 15) >> 4, stage[wib][1], lane);
        }
        __syncwarp();
        align_warp<K, GENERAL>(xs, n, ys, m, sc, bbuf, &out[e], lane);
    }
}

// Triangle range with pairs taken two at a time.  A work item is (a, b1, b2): the pairs (a, b1) and (a, b2) share
// the row sequence a.  Both pairs sweep the passes of the LONGER column sequence (the shorter one's pad slots are
// computed like real ones), so the two partners of an item should have about the same length: the partners b of a
// row that lie in the range are taken in order of length (`order`: all sequence indices, longest first), two at a
// time.  With neighbours in file order instead, a 400-900 bp set did 13 % more slots than it has cells.
// pa_duo_items_kernel writes the items of one range (one CTA per row, block-wide compaction of `order`), item
// `local_start(a) + j` = (a, j-th pair of partners); the DP kernels read items by index, results go to the pairs'
// triangle indices.  Pairs the s16x2 path cannot take (a non-A/C/G/T sequence, or longer than max_len16) are
// appended to `deferred` as range-relative indices.
constexpr uint32_t NO_PARTNER = 0xffffffffu;
__global__ void __launch_bounds__(256)
pa_duo_items_kernel(const uint32_t *order, const uint32_t N, const uint32_t a_lo, const uint32_t b_lo, const uint32_t a_hi,
                    const uint32_t b_hi, const unsigned long long *row_item_start, const uint32_t items_first, uint4 *items) {
    __shared__ uint32_t warp_tot[8];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (uint32_t a = a_lo + blockIdx.x; a <= a_hi; a += gridDim.x) {
        const uint32_t blo = (a == a_lo) ? b_lo : a + 1, bhi = (a == a_hi) ? b_hi : N - 1;
        // items of the rows before a in this range: the first row may be partial, the others are whole rows
        const unsigned long long base = (a == a_lo) ? 0ull : (unsigned long long)items_first + (row_item_start[a] - row_item_start[a_lo + 1]);
        uint32_t running = 0;
        for (uint32_t i0 = 0; i0 < N; i0 += 256) {
            const uint32_t i = i0 + threadIdx.x;
            const uint32_t idx = i < N ? order[i] : NO_PARTNER;
            const bool keep = (idx != NO_PARTNER) && idx >= blo && idx <= bhi;
            const uint32_t bal = __ballot_sync(FULL_MASK, keep);
            if (lane == 0) warp_tot[wib] = __popc(bal);
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) { const uint32_t t = warp_tot[k]; if (k < wib) before += t; total += t; }
            if (keep) {
                const uint32_t pos = running + before + __popc(bal & ((1u << lane) - 1u));
                uint32_t *it = reinterpret_cast<uint32_t *>(&items[base + (pos >> 1)]);
                if (pos & 1u) it[2] = idx;
                else { it[0] = a; it[1] = idx; }
            }
            running += total;
            __syncthreads();
        }
        if (threadIdx.x == 0 && (running & 1u)) reinterpret_cast<uint32_t *>(&items[base + (running >> 1)])[2] = NO_PARTNER;
    }
}

// Strip width (columns per lane) for a pair whose longer column sequence has m bases.  Column slots come in passes
// of 32*K and the pad slots of the first pass are computed like real ones, so a fixed K wastes up to a whole pass
// (m = 400 with K = 12: 768 slots).  Every pass has the same number of wavefront steps and a step costs about
// step_cost + 32*K issue slots, hence cost(K) ~ passes(K) * (step_cost + 32*K).  step_cost = 256 reproduces the
// measured K = 8 : K = 12 ratio (shuffle latency and loop overhead weigh more than their ~45 instructions).
// Widths in use: 8, 10, 11, 12, 13 (DUO_KSET).  K = 9 is left out: its build runs at 2/3 of the speed of its
// neighbours on the 400-900 bp set (measured, profiles/r01_duo_strip_width.txt); K = 7 never wins.
constexpr int DUO_KMIN = 8, DUO_KMAX = 13;
constexpr uint32_t DUO_KSET = (1u << 8) | (1u << 10) | (1u << 11) | (1u << 12) | (1u << 13);
constexpr int DUO_STEP_COST = 256;
__host__ __device__ __forceinline__ int duo_pick_k(const int m, const uint32_t kmask, const int step_cost) {
    int best_k = DUO_KMAX, best = 0x7fffffff;
#pragma unroll
    for (int k = DUO_KMAX; k >= DUO_KMIN; --k) {
        const int cost = ((m + 32 * k - 1) / (32 * k)) * (step_cost + 32 * k);
        if (((kmask >> k) & 1u) && cost < best) { best = cost; best_k = k; }
    }
    return best_k;
}

// K = 0: the strip width is chosen per work item (duo_pick_k); K > 0: fixed.
// GEC: compile-time gap extension (the host launches the GEC = -1 build for pairalign's own scoring), 0 = sc.ge.
template <int K, int MINB = 1, int GEC = 0>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, MINB)
pa_warp_duo_kernel(const SeqStore S, const Scoring sc, const uint64_t first, const uint4 *items, const uint64_t n_items,
                   const uint32_t max_len16, unsigned long long *work_counter, int4 *bbuf_all, const uint32_t bbuf_rows,
                   pa_pair_result *out, uint32_t *deferred, unsigned int *n_deferred,
                   const uint32_t kmask = DUO_KSET, const int step_cost = DUO_STEP_COST, const int win_ok = 0,
                   const int amb_launch = 0) {
    __shared__ __align__(16) uint32_t stage[WARPS_PER_CTA][3][STAGE_WORDS];
    __shared__ int4 tabs[WARPS_PER_CTA][8];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * WARPS_PER_CTA + wib;
    int4 *bbuf = bbuf_all + (size_t)gw * bbuf_rows;     // bbuf_rows - 1 usable rows + the virtual-column row
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
        // pairs this path cannot take go to the general / 32-bit kernels
        // longer than max_len16: floating-window variant (K = 0 kernel, when the host allows it), else deferred
        const uint32_t lim = (K == 0 && win_ok) ? 0xffffffffu : max_len16;
        // sequences with IUPAC ambiguity codes but no gap character (fastok, not pure): their items belong to the launch
        // of the 4-bit-set form (pa_warp_sets_kernel) when the host runs one (amb_launch); each item is done by exactly
        // one of the two launches, and only this one defers
        const uint8_t *good = (K == 0 && amb_launch) ? S.fastok : S.pure;
        const bool okx = good[a] && n > 0 && (uint32_t)n <= lim;
        const bool ok1 = okx && good[b1] && m1 > 0 && (uint32_t)m1 <= lim;
        const bool ok2 = okx && good[b2] && m2 > 0 && (uint32_t)m2 <= lim;
        if (lane == 0) {
            if (use1 && !ok1) deferred[atomicAdd(n_deferred, 1u)] = (uint32_t)(q1 - first);
            if (use2 && !ok2) deferred[atomicAdd(n_deferred, 1u)] = (uint32_t)(q2 - first);
        }
        use1 = use1 && ok1;
        use2 = use2 && ok2;
        if (!use1 && !use2) continue;
        // a half that is not wanted mirrors the other one
        const uint32_t y1 = use1 ? b1 : b2, y2 = use2 ? b2 : b1;
        const int my1 = use1 ? m1 : m2, my2 = use2 ? m2 : m1;
        // items with an ambiguous (fastok, not pure) sequence are the set-form launch's
        if (!(S.pure[a] && S.pure[y1] && S.pure[y2])) continue;
        __syncwarp();
        const uint32_t *xs = stage_seq(S.p2 + S.off2[a], (uint32_t)(n + 15) >> 4, stage[wib][0], lane);
        const uint32_t *ys1 = stage_seq(S.p2 + S.off2[y1], (uint32_t)(my1 + 15) >> 4, stage[wib][1], lane);
        const uint32_t *ys2 = stage_seq(S.p2 + S.off2[y2], (uint32_t)(my2 + 15) >> 4, stage[wib][2], lane);
        __syncwarp();
        pa_pair_result *r1 = use1 ? &out[q1 - first] : nullptr, *r2 = use2 ? &out[q2 - first] : nullptr;
        // y1 / y2 / my1 / my2: what is actually aligned (a half that is not wanted mirrors the other one)
        const bool too_long = (uint32_t)n > max_len16 || (uint32_t)my1 > max_len16 || (uint32_t)my2 > max_len16;
        if constexpr (K == 0) {
            if (too_long) {
                // bbuf rows come in pairs here (values, offsets); the virtual-column row is the last pair
                align_warp_duo<12, true, GEC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 2, tabs[wib], r1, r2, lane);
                continue;
            }
            // strip width per work item: the fewest issue slots for these lengths (duo_pick_k)
            switch (duo_pick_k(my1 > my2 ? my1 : my2, kmask & DUO_KSET, step_cost)) {
                case 8:  align_warp_duo<8, false, GEC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 1, tabs[wib], r1, r2, lane); break;
                case 10: align_warp_duo<10, false, GEC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 1, tabs[wib], r1, r2, lane); break;
                case 11: align_warp_duo<11, false, GEC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 1, tabs[wib], r1, r2, lane); break;
                case 12: align_warp_duo<12, false, GEC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 1, tabs[wib], r1, r2, lane); break;
                default: align_warp_duo<DUO_KMAX, false, GEC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 1, tabs[wib], r1, r2, lane); break;
            }
        } else {
            align_warp_duo<K, false, GEC>(xs, n, ys1, my1, ys2, my2, sc, bbuf, bbuf_rows - 1, tabs[wib], r1, r2, lane);
        }
    }
}

// pairalign -A: position-wise comparison over min(n,m) columns of the raw
// encoded sequences (src/pairalign.cpp:681, src/seqpair.cpp:238-274).  One warp
// per pair, lanes stride over 8-base words.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
pa_aligned_stats_kernel(const SeqStore S, const PairSource src, const uint64_t count,
                        unsigned long long *work_counter, pa_pair_result *out) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(work_counter, 1ull);
        w = __shfl_sync(FULL_MASK, w, 0);
        if (w >= count) break;
        uint32_t a, b;
        if (src.ia) { a = src.ia[w]; b = src.ib[w]; }
        else tri_pair(src.first + w, S.n_seq, a, b);
        const int n = (int)S.len[a], m = (int)S.len[b];
        const int L = n < m ? n : m;
        const uint32_t *xw = S.p4 + S.off4[a], *yw = S.p4 + S.off4[b];
        uint32_t d = 0, l = 0;
        for (int wd = lane; wd * 8 < L; wd += 32) {
            uint32_t x = __ldg(&xw[wd]), y = __ldg(&yw[wd]);
            const int rem = L - wd * 8;
            if (rem < 8) { const uint32_t keep = (1u << (rem * 4)) - 1u; x &= keep; y &= keep; }
            // per-nibble "non-zero" flags in bit 0 of each nibble
            uint32_t nx = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;
            uint32_t ny = (y | (y >> 1) | (y >> 2) | (y >> 3)) & 0x11111111u;
            const uint32_t xy = x & y;
            uint32_t nb = (xy | (xy >> 1) | (xy >> 2) | (xy >> 3)) & 0x11111111u;
            const uint32_t cmp = nx & ny;
            l += __popc(cmp);
            d += __popc(cmp & ~nb);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            d += __shfl_xor_sync(FULL_MASK, d, s);
            l += __shfl_xor_sync(FULL_MASK, l, s);
        }
        if (lane == 0) { pa_pair_result o; o.score = 0; o.dist = d; o.len = l; o.end_i = n - 1; o.end_j = m - 1; out[w] = o; }
    }
}

}  // namespace pa
