/* pa_oracle.h -- CPU restatement of phylommand's seqpair path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (phylommand_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity status: PINNED.  The restatement is checked against the unmodified
 * reference compiled from /root/reference/src (oracle/_ref, see oracle/Makefile)
 * and against golden vectors generated from it (tests/golden/, made by
 * oracle/make_golden.py).  The reference ships no golden vectors of its own.
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef PA_ORACLE_H
#define PA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-pair integer result; mirrors pa_pair_result in include/pairalign_b200.h. */
typedef struct {
    int32_t  score;   /* return value of seqpair::align()  (src/seqpair.cpp:189)          */
    uint32_t dist;    /* hamming_distance(false)           (src/seqpair.cpp:238-256)      */
    uint32_t len;     /* compared columns in similarity()  (src/seqpair.cpp:258-268)      */
    int32_t  end_i;   /* end cell chosen by the last row/column scan (src/seqpair.cpp:137-143) */
    int32_t  end_j;
} pa_oracle_result;

/* IUPAC char -> 4-bit mask, A=1 G=2 C=4 T=8 (src/seqpair.cpp:22-54).
 * Returns 0..15 for a known character ('-' -> 0), -1 for whitespace that is
 * silently skipped (src/seqpair.cpp:81), -2 for a character that is not in
 * the alphabet (warned about and skipped, src/seqpair.cpp:86). */
int pa_oracle_char_mask(unsigned char c);

/* translate_to_binary (src/seqpair.cpp:74-92): drops text[0], skips white
 * space and unknown characters.  out must hold len bytes.  Returns the number
 * of masks written; *n_unknown (optional) counts skipped unknown characters. */
size_t pa_oracle_encode(const char *text, size_t len, uint8_t *out, size_t *n_unknown);

/* translate_to_string (src/seqpair.cpp:62-72): mask -> first character in
 * ascending map<char> order whose mask is equal ('-' for 0, '.' for 15). */
char pa_oracle_mask_char(uint8_t mask);

/* cost() with the (match,mismatch) matrix pairalign installs
 * (src/seqpair.cpp:192-205, src/pairalign.cpp:682): INT_MIN if either mask is
 * empty, match if the masks intersect, else mismatch. */
int32_t pa_oracle_cost(uint8_t x, uint8_t y, int32_t match, int32_t mismatch);

/* Literal restatement of seqpair::align() (src/seqpair.cpp:95-190) with full
 * n*m matrices and the reference's traceback, followed by
 * hamming_distance(false)/similarity(false) on the aligned strings.
 * 32-bit wrap-around arithmetic (what the reference binary does when '-' is
 * present in unaligned input).  ax/ay (optional, capacity n+m+1 each) receive
 * the aligned mask strings, *alen their length.  Requires n>0 and m>0.
 * Returns 0, or -1 on allocation failure / bad arguments. */
int pa_oracle_align_full(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                         int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                         pa_oracle_result *res, uint8_t *ax, uint8_t *ay, int32_t *alen);

/* Forward-only restatement: same recurrences, but (dist,len) of the
 * reference's traceback path are carried through the DP (every cell has one
 * predecessor: src/seqpair.cpp:159-178), O(m) memory.  Bit-identical results
 * to pa_oracle_align_full. */
/* pa_oracle_align_full's walk as op bytes in alignment order: 0 = x[i] over y[j], 1 = x[i] over a gap,
 * 2 = a gap over y[j] (src/seqpair.cpp:146-188). */
int pa_oracle_align_ops(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                        pa_oracle_result *res, uint8_t *ops, int32_t *alen);

/* The same op string from a forward pass that keeps 2 bits per cell instead of three int matrices: for pairs of
 * tens of kilobases (30 kb x 30 kb: 225 MB). */
int pa_oracle_align_ops_compact(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                                int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                                pa_oracle_result *res, uint8_t *ops, int32_t *alen);

int pa_oracle_align_forward(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                            int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                            pa_oracle_result *res);

/* -A / --aligned path (src/pairalign.cpp:681 skips align()): position-wise
 * hamming_distance(false)/similarity(false) over min(n,m) columns. */
void pa_oracle_aligned_stats(const uint8_t *x, int32_t n, const uint8_t *y, int32_t m,
                             pa_oracle_result *res);

/* similarity() / 1-similarity() / jc_distance() / jc-(1-sim) from the integer
 * counts, in the reference's exact floating-point expression order
 * (src/seqpair.cpp:272-273, src/seqpair.h:96-99, src/pairalign.cpp:818). */
double pa_oracle_similarity(uint32_t dist, uint32_t len);
double pa_oracle_pdist(uint32_t dist, uint32_t len);
double pa_oracle_jc(uint32_t dist, uint32_t len);
double pa_oracle_diff(uint32_t dist, uint32_t len);

/* All-pairs forward alignment over n_seq encoded sequences (concatenated in
 * codes, sequence s = codes[offsets[s] .. offsets[s+1])), row-major upper
 * triangle order (0,1),(0,2)..(n-2,n-1), optionally restricted to the pair
 * index range [first,last) and run on n_threads OpenMP-free pthreads.
 * Used by bench.py as the "port" CPU baseline and by the large parity tests. */
int pa_oracle_all_pairs(const uint8_t *codes, const uint64_t *offsets, uint32_t n_seq,
                        int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                        uint64_t first, uint64_t last, int n_threads, pa_oracle_result *out);

#ifdef __cplusplus
}
#endif
#endif
