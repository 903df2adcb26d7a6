/* pa_oracle.c -- CPU restatement of phylommand's seqpair path (see pa_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: never linked into the product library.
 * Citations are relative to /root/reference/.
 */
#include "pa_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* 32-bit two's-complement wrap-around add: what the reference binary does when
 * cost() returns INT_MIN for a '-' in unaligned input (src/seqpair.cpp:193). */
static inline int32_t wadd(int32_t a, int32_t b) {
    return (int32_t)((uint32_t)a + (uint32_t)b);
}

/* src/seqpair.cpp:22-54 -- bit0=A bit1=G bit2=C bit3=T */
int pa_oracle_char_mask(unsigned char c) {
    switch (c) {
    case ' ': case '\n': case '\r': case '\t': return -1;   /* src/seqpair.cpp:81 */
    case 'A': case 'a': return 1;
    case 'G': case 'g': return 2;
    case 'C': case 'c': return 4;
    case 'T': case 't': return 8;
    case '-': return 0;
    case 'R': case 'r': return 1 | 2;
    case 'Y': case 'y': return 4 | 8;
    case 'S': case 's': return 2 | 4;
    case 'W': case 'w': return 1 | 8;
    case 'K': case 'k': return 2 | 8;
    case 'M': case 'm': return 1 | 4;
    case 'B': case 'b': return 2 | 4 | 8;
    case 'D': case 'd': return 1 | 2 | 8;
    case 'H': case 'h': return 1 | 4 | 8;
    case 'V': case 'v': return 1 | 2 | 4;
    case 'N': case 'n': case '.': return 15;
    default: return -2;                                     /* src/seqpair.cpp:86 */
    }
}

/* src/seqpair.cpp:74-92 -- note the loop starts at index 1. */
size_t pa_oracle_encode(const char *text, size_t len, uint8_t *out, size_t *n_unknown) {
    size_t n = 0, unk = 0;
    for (size_t i = 1; i < len; ++i) {
        int m = pa_oracle_char_mask((unsigned char)text[i]);
        if (m == -1) continue;
        if (m == -2) { ++unk; continue; }
        out[n++] = (uint8_t)m;
    }
    if (n_unknown) *n_unknown = unk;
    return n;
}

/* src/seqpair.cpp:62-72 -- first equal mask in ascending char order. */
char pa_oracle_mask_char(uint8_t mask) {
    static const char tab[16] = {
        /* 0 */ '-', /* 1 A */ 'A', /* 2 G */ 'G', /* 3 AG */ 'R',
        /* 4 C */ 'C', /* 5 AC */ 'M', /* 6 GC */ 'S', /* 7 AGC */ 'V',
        /* 8 T */ 'T', /* 9 AT */ 'W', /* 10 GT */ 'K', /* 11 AGT */ 'D',
        /* 12 CT */ 'Y', /* 13 ACT */ 'H', /* 14 GCT */ 'B', /* 15 */ '.'};
    return tab[mask & 15];
}

/* src/seqpair.cpp:192-205 with the matrix of src/seqpair.h:72-74. */
int32_t pa_oracle_cost(uint8_t x, uint8_t y, int32_t match, int32_t mismatch) {
    if (x == 0 || y == 0) return INT_MIN;   /* no (i,j) pair is ever set: value stays INT_MIN */
    return (x & y) ? match : mismatch;      /* best entry among compatible bases */
}

static void stats_over_columns(const uint8_t *ax, const uint8_t *ay, int32_t n,
                               uint32_t *dist, uint32_t *len) {
    uint32_t d = 0, l = 0;
    for (int32_t k = 0; k < n; ++k) {
        /* src/seqpair.cpp:245-246 and :265-266 with gap=false */
        if (ax[k] == 0 || ay[k] == 0) continue;
        ++l;
        if ((ax[k] & ay[k]) == 0) ++d;      /* src/seqpair.cpp:247-250 */
    }
    *dist = d;
    *len = l;
}

static int align_full_impl(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                           int32_t match, int32_t mismatch, int32_t GO, int32_t GE,
                           pa_oracle_result *res, uint8_t *ax_out, uint8_t *ay_out, int32_t *alen_out, uint8_t *ops_out) {
    if (n <= 0 || m <= 0 || !res) return -1;
    size_t cells = (size_t)n * (size_t)m;
    int32_t *A = (int32_t *)malloc(cells * sizeof(int32_t));    /* aligned  :98 */
    int32_t *Gy = (int32_t *)malloc(cells * sizeof(int32_t));   /* gap_y    :99 */
    int32_t *Gx = (int32_t *)malloc(cells * sizeof(int32_t));   /* gap_x    :100 */
    uint8_t *rx = (uint8_t *)malloc((size_t)n + m + 1);
    uint8_t *ry = (uint8_t *)malloc((size_t)n + m + 1);
    uint8_t *ro = (uint8_t *)malloc((size_t)n + m + 1);   /* 0: x over y, 1: x over a gap, 2: a gap over y */
    if (!A || !Gy || !Gx || !rx || !ry || !ro) { free(A); free(Gy); free(Gx); free(rx); free(ry); free(ro); return -1; }

#define AT(M_, i_, j_) M_[(size_t)(i_) * (size_t)m + (size_t)(j_)]
    for (int32_t i = 0; i < n; ++i) {
        for (int32_t j = 0; j < m; ++j) {
            int32_t c = pa_oracle_cost(x[i], y[j], match, mismatch);
            if (i == 0 || j == 0) {                              /* :103-120 */
                if (i == 0 && j == 0) AT(A, i, j) = c;
                else if (i == 0) AT(A, i, j) = wadd(AT(Gx, i, j - 1), c);
                else AT(A, i, j) = wadd(AT(Gy, i - 1, j), c);
                AT(Gy, i, j) = 0;
                AT(Gx, i, j) = 0;
            } else {                                             /* :121-129 */
                int32_t d = AT(A, i - 1, j - 1), u = AT(Gy, i - 1, j), l = AT(Gx, i, j - 1);
                if (d >= u && d >= l) AT(A, i, j) = wadd(d, c);
                else if (u > d && u > l) AT(A, i, j) = wadd(u, c);
                else AT(A, i, j) = wadd(l, c);
                int32_t open = wadd(d, GO);
                int32_t uy = wadd(u, GE), lx = wadd(l, GE);
                AT(Gy, i, j) = (open > uy) ? open : uy;
                AT(Gx, i, j) = (open > lx) ? open : lx;
            }
        }
    }
    /* end cell: last column rows ascending, then last row columns ascending, strict > (:134-143) */
    int32_t i = n - 1, j = m - 1;
    int32_t best = INT_MIN;
    for (int32_t pos = 0; pos < n; ++pos)
        if (AT(A, pos, m - 1) > best) { i = pos; j = m - 1; best = AT(A, pos, m - 1); }
    for (int32_t pos = 0; pos < m; ++pos)
        if (AT(A, n - 1, pos) > best) { j = pos; i = n - 1; best = AT(A, n - 1, pos); }
    res->score = best;
    res->end_i = i;
    res->end_j = j;

    int32_t k = 0;
    if (i < n - 1) {                                             /* :146-151 */
        for (int32_t pos = n - 1; pos > i; --pos) { rx[k] = x[pos]; ry[k] = 0; ro[k] = 1; ++k; }
    } else if (j < m - 1) {                                      /* :152-157 */
        for (int32_t pos = m - 1; pos > j; --pos) { rx[k] = 0; ry[k] = y[pos]; ro[k] = 2; ++k; }
    }
    while (i >= 0 || j >= 0) {                                   /* :159-178 */
        if (i >= 0 && j >= 0 && AT(A, i, j) >= AT(Gy, i, j) && AT(A, i, j) >= AT(Gx, i, j)) {
            rx[k] = x[i]; ry[k] = y[j]; ro[k] = 0; ++k; --i; --j;
        } else if (j < 0 || (i >= 0 && AT(Gy, i, j) >= AT(A, i, j) && AT(Gy, i, j) >= AT(Gx, i, j))) {
            rx[k] = x[i]; ry[k] = 0; ro[k] = 1; ++k; --i;
        } else if (i < 0 || (j >= 0 && AT(Gx, i, j) >= AT(A, i, j) && AT(Gx, i, j) >= AT(Gy, i, j))) {
            rx[k] = 0; ry[k] = y[j]; ro[k] = 2; ++k; --j;
        }
    }
#undef AT
    /* reverse (:183-188) */
    for (int32_t a = 0, b = k - 1; a < b; ++a, --b) {
        uint8_t t = rx[a]; rx[a] = rx[b]; rx[b] = t;
        t = ry[a]; ry[a] = ry[b]; ry[b] = t;
        t = ro[a]; ro[a] = ro[b]; ro[b] = t;
    }
    stats_over_columns(rx, ry, k, &res->dist, &res->len);
    if (ax_out) memcpy(ax_out, rx, (size_t)k);
    if (ay_out) memcpy(ay_out, ry, (size_t)k);
    if (alen_out) *alen_out = k;
    if (ops_out) memcpy(ops_out, ro, (size_t)k);
    free(A); free(Gy); free(Gx); free(rx); free(ry); free(ro);
    return 0;
}

int pa_oracle_align_full(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                         int32_t match, int32_t mismatch, int32_t GO, int32_t GE,
                         pa_oracle_result *res, uint8_t *ax_out, uint8_t *ay_out, int32_t *alen_out) {
    return align_full_impl(x, n, y, m, match, mismatch, GO, GE, res, ax_out, ay_out, alen_out, NULL);
}

/* The same walk as one op byte per aligned column (what pa_align_pairs_ops returns): needed where the gapped
 * mask strings are ambiguous, i.e. when a sequence itself holds '-' (mask 0). */
int pa_oracle_align_ops(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                        int32_t match, int32_t mismatch, int32_t GO, int32_t GE,
                        pa_oracle_result *res, uint8_t *ops_out, int32_t *alen_out) {
    return align_full_impl(x, n, y, m, match, mismatch, GO, GE, res, NULL, NULL, alen_out, ops_out);
}

/* Forward-only form.  The traceback's move at (i,j) depends only on the three
 * matrices at (i,j) (src/seqpair.cpp:160-177), so the (dist,len) the reference
 * would count over its path ending in (i,j) obeys
 *   D: cnt(i-1,j-1) + [both masks non-empty]*(1 column, mismatch?)
 *   U: cnt(i-1,j)      L: cnt(i,j-1)      cnt = 0 outside the matrix,
 * D iff A>=Gy && A>=Gx, else U iff Gy>=Gx, else L.  Overhang columns pair a
 * base with a gap and are never counted (src/seqpair.cpp:246). */
int pa_oracle_align_forward(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                            int32_t match, int32_t mismatch, int32_t GO, int32_t GE,
                            pa_oracle_result *res) {
    if (n <= 0 || m <= 0 || !res) return -1;
    int32_t *A = (int32_t *)calloc((size_t)m, sizeof(int32_t));     /* A(i-1,.) then A(i,.) */
    int32_t *Gy = (int32_t *)calloc((size_t)m, sizeof(int32_t));
    uint32_t *cd = (uint32_t *)calloc((size_t)m, sizeof(uint32_t));
    uint32_t *cl = (uint32_t *)calloc((size_t)m, sizeof(uint32_t));
    if (!A || !Gy || !cd || !cl) { free(A); free(Gy); free(cd); free(cl); return -1; }

    int32_t best = INT_MIN, bi = n - 1, bj = m - 1;
    uint32_t bd = 0, bl = 0;
    for (int32_t i = 0; i < n; ++i) {
        int32_t diagA = 0, leftGx = 0;
        uint32_t diag_d = 0, diag_l = 0, left_d = 0, left_l = 0;
        for (int32_t j = 0; j < m; ++j) {
            int32_t c = pa_oracle_cost(x[i], y[j], match, mismatch);
            int32_t upA = A[j], upGy = Gy[j];
            uint32_t up_d = cd[j], up_l = cl[j];
            int32_t a, gy, gx;
            if (i == 0 || j == 0) {
                if (i == 0 && j == 0) a = c;
                else if (i == 0) a = wadd(leftGx, c);
                else a = wadd(upGy, c);
                gy = 0; gx = 0;
            } else {
                int32_t mx = diagA;
                if (upGy > mx) mx = upGy;
                if (leftGx > mx) mx = leftGx;
                a = wadd(mx, c);
                int32_t open = wadd(diagA, GO);
                int32_t uy = wadd(upGy, GE), lx = wadd(leftGx, GE);
                gy = (open > uy) ? open : uy;
                gx = (open > lx) ? open : lx;
            }
            uint32_t nd, nl;
            if (a >= gy && a >= gx) {            /* D: predecessor (i-1,j-1) or outside */
                nd = (i > 0 && j > 0) ? diag_d : 0;
                nl = (i > 0 && j > 0) ? diag_l : 0;
                if (x[i] != 0 && y[j] != 0) { ++nl; if ((x[i] & y[j]) == 0) ++nd; }
            } else if (gy >= gx) {               /* U: predecessor (i-1,j) or outside */
                nd = (i > 0) ? up_d : 0;
                nl = (i > 0) ? up_l : 0;
            } else {                             /* L: predecessor (i,j-1) or outside */
                nd = (j > 0) ? left_d : 0;
                nl = (j > 0) ? left_l : 0;
            }
            /* shift the window */
            diagA = upA; diag_d = up_d; diag_l = up_l;
            A[j] = a; Gy[j] = gy; cd[j] = nd; cl[j] = nl;
            leftGx = gx; left_d = nd; left_l = nl;
            if (j == m - 1 && a > best) { best = a; bi = i; bj = j; bd = nd; bl = nl; }
        }
    }
    for (int32_t j = 0; j < m; ++j)
        if (A[j] > best) { best = A[j]; bi = n - 1; bj = j; bd = cd[j]; bl = cl[j]; }
    /* every candidate equals INT_MIN: the reference keeps its initial i=n-1, j=m-1 (:132-133) */
    if (best == INT_MIN) { bd = cd[m - 1]; bl = cl[m - 1]; }
    res->score = best; res->end_i = bi; res->end_j = bj; res->dist = bd; res->len = bl;
    free(A); free(Gy); free(cd); free(cl);
    return 0;
}

/* pa_oracle_align_ops for pairs whose three full matrices do not fit (30 kb x 30 kb: 10.8 GB): the forward-only
 * recurrence above with the move of every cell kept in 2 bits (225 MB for such a pair), then the reference's walk
 * (src/seqpair.cpp:146-178) over the moves.  Checked against pa_oracle_align_ops on small pairs (tests/test_oracle.py);
 * it exists so that the op strings of the CUDA path can be compared exactly at BASELINE.json config 5 sizes. */
int pa_oracle_align_ops_compact(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                                int32_t match, int32_t mismatch, int32_t GO, int32_t GE,
                                pa_oracle_result *res, uint8_t *ops_out, int32_t *alen_out) {
    if (n <= 0 || m <= 0 || !res || !ops_out) return -1;
    const size_t row_bytes = ((size_t)m + 3) / 4;
    uint8_t *mv = (uint8_t *)calloc((size_t)n * row_bytes, 1);      /* 0 D, 1 U, 2 L */
    int32_t *A = (int32_t *)calloc((size_t)m, sizeof(int32_t));
    int32_t *Gy = (int32_t *)calloc((size_t)m, sizeof(int32_t));
    if (!mv || !A || !Gy) { free(mv); free(A); free(Gy); return -1; }
    int32_t best = INT_MIN, bi = n - 1, bj = m - 1;
    for (int32_t i = 0; i < n; ++i) {
        int32_t diagA = 0, leftGx = 0;
        uint8_t *row = mv + (size_t)i * row_bytes;
        for (int32_t j = 0; j < m; ++j) {
            int32_t c = pa_oracle_cost(x[i], y[j], match, mismatch);
            int32_t upA = A[j], upGy = Gy[j];
            int32_t a, gy, gx;
            if (i == 0 || j == 0) {
                if (i == 0 && j == 0) a = c;
                else if (i == 0) a = wadd(leftGx, c);
                else a = wadd(upGy, c);
                gy = 0; gx = 0;
            } else {
                int32_t mx = diagA;
                if (upGy > mx) mx = upGy;
                if (leftGx > mx) mx = leftGx;
                a = wadd(mx, c);
                int32_t open = wadd(diagA, GO);
                int32_t uy = wadd(upGy, GE), lx = wadd(leftGx, GE);
                gy = (open > uy) ? open : uy;
                gx = (open > lx) ? open : lx;
            }
            unsigned move = (a >= gy && a >= gx) ? 0u : (gy >= gx ? 1u : 2u);
            row[j >> 2] |= (uint8_t)(move << ((j & 3) * 2));
            diagA = upA;
            A[j] = a; Gy[j] = gy;
            leftGx = gx;
            if (j == m - 1 && a > best) { best = a; bi = i; bj = j; }
        }
    }
    for (int32_t j = 0; j < m; ++j)
        if (A[j] > best) { best = A[j]; bi = n - 1; bj = j; }
    res->score = best; res->end_i = bi; res->end_j = bj;
    int32_t i = bi, j = bj, k = 0;
    uint32_t d = 0, l = 0;
    if (i < n - 1) { for (int32_t pos = n - 1; pos > i; --pos) ops_out[k++] = 1; }
    else if (j < m - 1) { for (int32_t pos = m - 1; pos > j; --pos) ops_out[k++] = 2; }
    while (i >= 0 || j >= 0) {
        unsigned move = 3;
        if (i >= 0 && j >= 0) move = (mv[(size_t)i * row_bytes + (size_t)(j >> 2)] >> ((j & 3) * 2)) & 3u;
        if (move == 0) {
            if (x[i] != 0 && y[j] != 0) { ++l; if ((x[i] & y[j]) == 0) ++d; }
            ops_out[k++] = 0; --i; --j;
        } else if (j < 0 || (i >= 0 && move == 1)) { ops_out[k++] = 1; --i; }
        else { ops_out[k++] = 2; --j; }
    }
    for (int32_t a = 0, b = k - 1; a < b; ++a, --b) { uint8_t t = ops_out[a]; ops_out[a] = ops_out[b]; ops_out[b] = t; }
    res->dist = d; res->len = l;
    if (alen_out) *alen_out = k;
    free(mv); free(A); free(Gy);
    return 0;
}

void pa_oracle_aligned_stats(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                             pa_oracle_result *res) {
    int32_t k = n < m ? n : m;                   /* src/seqpair.cpp:239,259 */
    stats_over_columns(x, y, k, &res->dist, &res->len);
    res->score = 0; res->end_i = n - 1; res->end_j = m - 1;
}

double pa_oracle_similarity(uint32_t dist, uint32_t len) {
    /* hamming_distance() returns int and is divided by double(length): src/seqpair.cpp:272-273 */
    if (len > 0) return 1.0 - ((int)dist / (double)(int)len);
    return 1.0;
}
double pa_oracle_pdist(uint32_t dist, uint32_t len) { return 1 - pa_oracle_similarity(dist, len); }
double pa_oracle_jc(uint32_t dist, uint32_t len) {
    double p = 1 - pa_oracle_similarity(dist, len);          /* src/seqpair.h:97 */
    return log(1.0 - (4.0 / 3.0) * p) * (-3.0 / 4.0);        /* src/seqpair.h:98 */
}
double pa_oracle_diff(uint32_t dist, uint32_t len) {
    return pa_oracle_jc(dist, len) - (1.0 - pa_oracle_similarity(dist, len));  /* src/pairalign.cpp:818 */
}

/* ---- all-pairs driver (CPU baseline "port" and large-case checker) ---- */
typedef struct {
    const uint8_t *codes; const uint64_t *offsets; uint32_t n_seq;
    int32_t match, mismatch, go, ge;
    uint64_t first, last; int tid, nthreads; pa_oracle_result *out; int rc;
} ap_job;

static void pair_from_index(uint64_t k, uint32_t n, uint32_t *a, uint32_t *b) {
    /* row-major upper triangle: row a holds n-1-a pairs */
    uint32_t r = 0; uint64_t rem = k;
    while (rem >= (uint64_t)(n - 1 - r)) { rem -= (uint64_t)(n - 1 - r); ++r; }
    *a = r; *b = r + 1 + (uint32_t)rem;
}

static void *ap_worker(void *arg) {
    ap_job *jb = (ap_job *)arg;
    for (uint64_t k = jb->first + (uint64_t)jb->tid; k < jb->last; k += (uint64_t)jb->nthreads) {
        uint32_t a, b;
        pair_from_index(k, jb->n_seq, &a, &b);
        const uint8_t *xa = jb->codes + jb->offsets[a];
        const uint8_t *yb = jb->codes + jb->offsets[b];
        int32_t n = (int32_t)(jb->offsets[a + 1] - jb->offsets[a]);
        int32_t m = (int32_t)(jb->offsets[b + 1] - jb->offsets[b]);
        if (pa_oracle_align_forward(xa, n, yb, m, jb->match, jb->mismatch, jb->go, jb->ge,
                                    &jb->out[k - jb->first]) != 0) { jb->rc = -1; return 0; }
    }
    return 0;
}

int pa_oracle_all_pairs(const uint8_t *codes, const uint64_t *offsets, uint32_t n_seq,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                        uint64_t first, uint64_t last, int n_threads, pa_oracle_result *out) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256]; ap_job jobs[256];
    for (int t = 0; t < n_threads; ++t) {
        ap_job j = {codes, offsets, n_seq, match, mismatch, gap_open, gap_ext, first, last, t, n_threads, out, 0};
        jobs[t] = j;
        if (n_threads == 1) ap_worker(&jobs[t]);
        else if (pthread_create(&th[t], 0, ap_worker, &jobs[t]) != 0) return -1;
    }
    int rc = 0;
    for (int t = 0; t < n_threads; ++t) {
        if (n_threads > 1) pthread_join(th[t], 0);
        if (jobs[t].rc) rc = -1;
    }
    return rc;
}
