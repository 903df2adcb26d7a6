/* pairalign_b200.h -- C-ABI of the B200 (sm_100a) pairalign hot path.
 *
 * This is the drop-in boundary for phylommand's `seqpair` class as it is used
 * by pairalign's all-pairs loop.  The reference has no FFI of its own; the
 * seam is the public surface of `seqpair` (reference src/seqpair.h:44-120)
 * called once per pair from align_pair() (src/pairalign.cpp:675-685).  The
 * entry points below are the *batched* equivalents: load the sequences once,
 * align many pairs per call, read back one integer record per pair.  Every
 * floating-point figure the CLI prints is derived on the host from that record
 * with the reference's own expressions (pa_similarity .. pa_jc_minus_p), so the
 * text output is bit-identical.
 *
 * Plain C types only; no C++/torch types, exceptions or streams cross this
 * boundary.  All functions return PA_OK (0) or a negative PA_E* code;
 * pa_last_error() gives a message for the calling thread's last failure.
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with PA_ENODEVICE.
 *
 * Citations are relative to the reference checkout (RybergGroup/phylommand).
 */
#ifndef PAIRALIGN_B200_H
#define PAIRALIGN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_API_VERSION 1

enum {
    PA_OK = 0,
    PA_EINVAL = -1,     /* bad argument                                        */
    PA_ENODEVICE = -2,  /* no usable CUDA device / not initialised             */
    PA_ECUDA = -3,      /* a CUDA call failed (message in pa_last_error())     */
    PA_ENOMEM = -4,     /* host or device allocation failed                    */
    PA_ERANGE = -5      /* sequence too long (> PA_MAX_SEQ_LEN) or index range */
};

/* Longest encoded sequence accepted.  The reference itself breaks earlier:
 * `new int[n*m]` with an int product (src/seqpair.cpp:98) overflows beyond
 * 46340 x 46340. */
#define PA_MAX_SEQ_LEN 65535u

/* Scoring.  pairalign hard-codes match 7 / mismatch -5 (src/pairalign.cpp:682)
 * and gap open -15 / extend -1 (src/seqpair.h:47-48,57-58); replaces
 * seqpair::set_cost_matrix(int,int) (src/seqpair.h:72-74) and
 * seqpair::set_gap_penalty(int,int) (src/seqpair.h:75-78). */
typedef struct {
    int32_t match;
    int32_t mismatch;
    int32_t gap_open;
    int32_t gap_ext;
    int32_t aligned;   /* != 0: pairalign -A, skip the DP (src/pairalign.cpp:681) and
                          compare position-wise (src/seqpair.cpp:238-274)          */
} pa_params;

/* One record per pair; replaces the return value of seqpair::align()
 * (src/seqpair.cpp:189), hamming_distance(false) (src/seqpair.cpp:238-256)
 * and the column count of similarity(false) (src/seqpair.cpp:258-268). */
typedef struct {
    int32_t  score;  /* best last-row/last-column cell (src/seqpair.cpp:134-143); 0 with -A */
    uint32_t dist;   /* compared columns whose IUPAC sets do not intersect                   */
    uint32_t len;    /* compared columns (neither side a gap)                                */
    int32_t  end_i;  /* end cell of the alignment proper, 0-based, in the encoded sequences   */
    int32_t  end_j;
} pa_pair_result;

/* Phase timings of the last align call on this thread's context, measured with
 * CUDA events on the streams the work was issued on (milliseconds). */
typedef struct {
    double   h2d_ms;         /* pair lists / parameters host->device                       */
    double   kernel_ms;      /* all DP kernels of the call (max over devices)              */
    double   d2h_ms;         /* results device->host                                       */
    double   total_ms;       /* wall clock of the call                                     */
    uint64_t cells;          /* DP cells computed: sum of n*m over the pairs               */
    uint64_t pairs;
    uint32_t kernel_launches;
    uint32_t n_devices;
    double   dp_fast_ms;     /* the 32-bit 2-bit register-wavefront kernel (one pair per warp) */
    double   dp_general_ms;  /* the IUPAC/gap (int32 wrap) kernel alone                    */
    double   dp_duo_ms;      /* the s16x2 kernel (two pairs per warp); with -A: the stats kernel */
    double   dp_cta_ms;      /* the CTA-per-pair kernel for long pairs                       */
    double   walk_ms;        /* pa_align_pairs_ops: the walk back over the stored moves (in kernel_ms) */
} pa_timing;

/* ---- life cycle --------------------------------------------------------- */

/* Create the process-wide context on the given CUDA devices (devices==NULL or
 * n_dev<=0: device 0 only).  Calling it again re-initialises. */
int pa_init(const int *devices, int n_dev);
void pa_shutdown(void);
int pa_device_count(void);          /* devices in the context (0 if none)      */
int pa_visible_devices(void);       /* CUDA devices this process can see (0 if none) */
int pa_api_version(void);
const char *pa_last_error(void);

/* ---- sequence front end (replaces seqpair's constructor) ------------------ */

/* IUPAC character -> 4-bit set, A=1 G=2 C=4 T=8 (set_DNA_alphapet,
 * src/seqpair.cpp:22-60).  Returns 0..15 ('-' -> 0, 'N' and '.' -> 15),
 * -1 for white space that is skipped (src/seqpair.cpp:81), -2 for a character
 * outside the alphabet (skipped with a warning, src/seqpair.cpp:86). */
int pa_char_to_mask(unsigned char c);

/* 4-bit set -> character, translate_to_string (src/seqpair.cpp:62-72):
 * first match in ascending char order, so 0 -> '-', 15 -> '.', upper case. */
char pa_mask_to_char(uint8_t mask);

/* translate_to_binary (src/seqpair.cpp:74-92): encodes text[1..len) -- the
 * reference silently drops the first character -- skipping white space and
 * unknown characters.  out needs len bytes.  Returns the number of masks
 * written; *n_unknown (optional) counts unknown characters and
 * unknown_chars (optional, capacity unknown_cap) receives them in order so the
 * caller can reproduce the reference's stderr warnings. */
size_t pa_encode_sequence(const char *text, size_t len, uint8_t *out,
                          size_t *n_unknown, char *unknown_chars, size_t unknown_cap);

/* Upload n_seq encoded sequences (one 4-bit mask per byte, concatenated;
 * sequence s is masks[offsets[s] .. offsets[s+1])) to every device of the
 * context.  The devices pack them (2 bit/base and 4 bit/base, pa_pack_kernel)
 * and keep them resident until the next upload. */
int pa_upload_sequences(const uint8_t *masks, const uint64_t *offsets, uint32_t n_seq);
uint32_t pa_num_sequences(void);

/* ---- the hot path ---------------------------------------------------------- */

/* Number of unordered pairs of the uploaded set: n(n-1)/2. */
uint64_t pa_num_pairs(void);

/* Pair k of the row-major upper triangle (0,1),(0,2)..(0,n-1),(1,2).. -- the
 * order pairalign's loop visits them when the sequences are uploaded in
 * std::map order (src/seqdatabase.cpp:69-115, src/indexedfasta.h:41-59). */
int pa_pair_from_index(uint64_t k, uint32_t *a, uint32_t *b);

/* Align pairs [first, first+count) of the upper triangle; out[k-first] is
 * written for each (host memory, caller-owned).  The range is split over the
 * devices of the context, balanced by DP cells.  Replaces the body of the
 * while-loop of cluster() (src/pairalign.cpp:527-628) up to the statistics. */
int pa_align_all_pairs(const pa_params *params, uint64_t first, uint64_t count,
                       pa_pair_result *out);

/* Align an explicit list of pairs (ia[k], ib[k]) -> out[k]. */
int pa_align_pairs(const pa_params *params, const uint32_t *ia, const uint32_t *ib,
                   uint64_t count, pa_pair_result *out);

/* Alignment of one pair as pairalign -a prints it: get_x()/get_y() after
 * align() (src/seqpair.cpp:146-188, src/seqpair.h:83-84).  ax/ay receive
 * 4-bit sets (0 = gap), capacity cap each (n+m is always enough); *alen the
 * number of columns.  res (optional) receives the pair record. */
int pa_align_pair_traceback(const pa_params *params, uint32_t a, uint32_t b,
                            uint8_t *ax, uint8_t *ay, uint32_t cap, uint32_t *alen,
                            pa_pair_result *res);

/* pairalign -a for many pairs at once.  The alignment of pair k comes back as an op string
 * ops[op_offsets[k] .. op_offsets[k] + n_ops[k]), one byte per aligned column in alignment order:
 * 0 = a base of x over a base of y, 1 = a base of x over a gap, 2 = a gap over a base of y
 * (bases are consumed from the encoded sequences in order).  op_offsets (count+1 entries) is filled by
 * the call with the prefix sum of len(ia[k]) + len(ib[k]); ops_cap must be at least op_offsets[count].
 * res (optional) receives the pair records.  Replaces get_x()/get_y() after align() (src/seqpair.h:83-84,
 * src/seqpair.cpp:146-188) for every pair of the loop.  Every device of the context takes a contiguous share
 * of the list (balanced by DP cells); on each, pairs are processed in batches sized to device memory
 * (2 bits per DP cell). */
int pa_align_pairs_ops(const pa_params *params, const uint32_t *ia, const uint32_t *ib, uint64_t count,
                       uint8_t *ops, uint64_t ops_cap, uint64_t *op_offsets, uint32_t *n_ops, pa_pair_result *res);

/* Split [first, first+count) into n_parts contiguous ranges with nearly equal
 * DP cells (sum of n*m); bounds gets n_parts+1 ascending pair indices.  Used
 * to shard the triangle over devices / ranks. */
int pa_partition_pairs(uint64_t first, uint64_t count, uint32_t n_parts, uint64_t *bounds);

/* The same split computed from the encoded lengths alone: needs no device and no
 * context (every rank of a multi-process run calls it and takes its own range).
 * cells (optional, n_parts entries) receives the DP cells of each part. */
int pa_partition_by_length(const uint32_t *lengths, uint32_t n_seq, uint64_t first, uint64_t count,
                           uint32_t n_parts, uint64_t *bounds, uint64_t *cells);

/* Sum of n*m over pairs [first, first+count). */
uint64_t pa_count_cells(uint64_t first, uint64_t count);

int pa_get_timing(pa_timing *t);

/* ---- per-pair statistics (host; the reference's exact expressions) -------- */

/* similarity(false): len>0 ? 1.0-(dist/double(len)) : 1.0  (src/seqpair.cpp:272-273) */
double pa_similarity(uint32_t dist, uint32_t len);
/* 1-similarity()                                           (src/pairalign.cpp:842)   */
double pa_pdistance(uint32_t dist, uint32_t len);
/* jc_distance(): log(1.0-(4.0/3.0)*p)*(-3.0/4.0)           (src/seqpair.h:96-99)     */
double pa_jc_distance(uint32_t dist, uint32_t len);
/* jc_distance()-(1.0-similarity())                         (src/pairalign.cpp:818)   */
double pa_jc_minus_p(uint32_t dist, uint32_t len);

/* ---- neighbour joining: the consumer of the -m distance matrix ------------- */

/* One join of njtree::build_nj_tree (src/nj_tree.cpp:79-103).  Node ids: 0..n-1 are the
 * taxa in matrix order, n.. the joins in creation order.  The lengths are the reference's
 * float expressions (src/nj_tree.cpp:92-94) stored in double like node::branchlength. */
typedef struct {
    uint32_t left, right;
    double   left_len, right_len;
} pa_nj_join;

#define PA_NJ_MAX_TAXA 65000u

/* Neighbour joining with the reference's arithmetic (float sums in its order, float Q values,
 * first strictly smallest pair, new node first in the order; src/nj_tree.cpp:32-205) on the
 * current CUDA device.  dist: row-major upper triangle, n(n-1)/2 floats, as `treeator -n` reads
 * them from pairalign -m output (read_distance_matrix, src/nj_tree.cpp:252-352).  joins
 * receives n-2 records; the tree is (root_left:0, root_right:root_right_len)
 * (src/nj_tree.cpp:193-201).  kernel_ms (optional): CUDA-event time of the joining with the
 * matrix resident in HBM.  Replaces njtree::build_nj_tree(). */
int pa_nj_build(const float *dist, uint32_t n, pa_nj_join *joins, uint32_t *root_left,
                uint32_t *root_right, double *root_right_len, double *kernel_ms);
/* Kernels launched and algorithmic bytes moved by the last pa_nj_build of this thread. */
int pa_nj_last_stats(uint64_t *launches, uint64_t *bytes);

/* ---- measurement and test support (bench.py, tests/; not part of the drop-in surface) ---------- */

/* Same as pa_align_all_pairs but the results stay on device 0 of the context
 * (d_out: device pointer to count records) and nothing is copied back; used to
 * time the kernels with inputs and outputs resident in HBM. */
int pa_align_all_pairs_device(const pa_params *params, uint64_t first, uint64_t count,
                              void *d_out);

/* Device-free: the longest sequence the s16x2 kernel takes with plain 16-bit scores for these parameters
 * (longer A/C/G/T pairs use its floating-window form or the int32 kernels) and the bias it stores states with
 * (stored = true + bias; both 16-bit halves stay negative).  0 / 0 when the parameters
 * are outside the byte-table range (general kernel).  The tests use it to build inputs at the limit. */
int pa_s16_limits(const pa_params *params, uint32_t *max_len, int32_t *bias);

/* INT32 issue-rate micro-benchmark on device 0 of the context: independent
 * chains of the instruction classes the DP uses.  which: 0 IADD3, 1 VIMNMX3,
 * 2 VIADDMNMX, 3 IMAD, 4 PRMT, 5 ISETP+SEL pair, 6 mixed IADD3+IMAD,
 * 7 VIMNMX3.S16x2, 8 VIADDMNMX.S16x2, 9 VIADD.16x2, 10 LOP3, 11 SHFL.
 * Returns giga warp-lane operations per second in *gops (one operation = one
 * instruction executed by one thread) and the SM clock seen in *sm_mhz. */
int pa_int32_peak(int which, double *gops, double *sm_mhz);

#ifdef __cplusplus
}
#endif
#endif /* PAIRALIGN_B200_H */
