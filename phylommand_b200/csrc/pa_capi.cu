// pa_capi.cu -- implementation of include/pairalign_b200.h: context, sequence
// store, launch logic and result transfer.  No CPU fallback exists here: every
// compute entry point needs a CUDA device.
#include "pa_dp.cuh"
#include "pa_dp32.cuh"
#include "pa_dp_sets.cuh"
#include "pa_dp_moves.cuh"
#include "pa_peak.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

using namespace pa;

constexpr int KDUO = 12;      // columns per lane, s16x2 two-pairs-per-warp path
constexpr int KFAST = 16;     // columns per lane, int32 2-bit path (warp kernel and CTA kernel)
constexpr uint32_t LONG_LEN = 8192;   // longer A/C/G/T pairs take a whole CTA (pa_cta32_kernel)
constexpr int KGEN = 8;       // columns per lane, IUPAC/gap path
constexpr uint64_t CHUNK_PAIRS = 1ull << 22;   // pairs per launch (84 MB of records)

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(PA_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct Device {
    int id = -1;
    int n_sm = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    // sequence store
    uint32_t *p2 = nullptr, *p4 = nullptr, *off2 = nullptr, *off4 = nullptr, *len = nullptr;
    uint8_t *pure = nullptr;
    uint8_t *raw = nullptr;                   // the uploaded 4-bit sets, one per byte: input of pa_pack_kernel only
    unsigned long long *raw_off = nullptr;    // their offsets (n_seq + 1)
    size_t cap_raw = 0, cap_raw_off = 0;
    // scratch
    unsigned long long *counters = nullptr;   // [0] fast work counter, [1] general work counter
    unsigned int *n_deferred = nullptr;       // [0] deferred by the s16x2 kernel, [1] non-A/C/G/T, [2] long pairs
    uint32_t *deferred = nullptr, *deferred2 = nullptr, *deferred3 = nullptr;
    size_t deferred_cap = 0;
    unsigned long long *row_items = nullptr;  // items (pairs of pairs) in rows before a; n_seq+1 entries
    size_t cap_row_items = 0, cap_p2 = 0, cap_p4 = 0, cap_off2 = 0, cap_off4 = 0, cap_len = 0, cap_pure = 0, cap_bbuf = 0;
    uint8_t *fastok = nullptr;                // no gap character: the sequence can run on the s16x2 kernels
    size_t cap_fastok = 0;
    uint32_t *order = nullptr;                // sequence indices, longest first (partners of an s16x2 work item: similar lengths)
    uint4 *duo_items = nullptr;               // (a, b1, b2) work items of the chunk being launched (pa_duo_items_kernel)
    size_t cap_order = 0, cap_duo_items = 0;
    int4 *bbuf = nullptr;
    uint32_t bbuf_rows = 0;
    uint32_t n_warps = 0;
    int grid_duo = 0, grid_duo3 = 0, grid_duo8 = 0, grid_duo_auto = 0, grid_sets = 0, grid_fast = 0, grid_cta = 0, grid_gen = 0, grid_stats = 0;
    pa_pair_result *d_out[2] = {nullptr, nullptr};
    size_t d_out_cap = 0;
    pa_pair_result *h_stage[2] = {nullptr, nullptr};
    uint32_t *d_ia = nullptr, *d_ib = nullptr;
    size_t d_pairs_cap = 0;
    // pairalign -a: move store, op strings and their offsets for one batch
    uint8_t *d_dirs = nullptr, *d_ops = nullptr;
    unsigned long long *d_dirs_off = nullptr, *d_ops_off = nullptr;
    uint32_t *d_nops = nullptr;
    pa_pair_result *d_res = nullptr;
    size_t cap_dirs = 0, cap_ops = 0, cap_tb_pairs = 0, cap_items = 0;
    uint2 *d_items = nullptr;                 // work items of the move-storing s16x2 kernels (entries of the batch, in twos)
    int grid_moves_warp = 0, grid_moves_warp_sets = 0, grid_moves_cta = 0;
    bool moves_smem_set = false;
    // Events of one in-flight chunk (two chunks are in flight: slot = chunk & 1).  k[0]..k[1] s16x2 stage,
    // k[1]..k[2] 32-bit warp stage, k[2]..k[3] CTA-per-pair stage, k[3]..k[4] general stage; k[4] also releases the
    // D2H copy of the chunk on copy_stream, `done` marks its end (and lets the next kernel reuse d_out[slot]).
    struct ChunkEvents {
        cudaEvent_t k[5] = {};
        cudaEvent_t done = nullptr;
        bool duo = false, fast = false, cta = false, gen = false;   // which DP kernels the chunk launched
    } ce[2];
    cudaEvent_t ev[6] = {};                       // pa_align_pairs_ops, pair-list upload
    double cta_ms = 0;
    // timing accumulators of the last call
    double duo_ms = 0, fast_ms = 0, gen_ms = 0, d2h_ms = 0, h2d_ms = 0;
    uint32_t launches = 0;
};

// Lengths of the uploaded set with the prefix sums that balancing needs.
struct Triangle {
    uint32_t n_seq = 0;
    std::vector<uint32_t> len;
    std::vector<uint64_t> pref;                        // pref[s] = sum of len[0..s)
    std::vector<unsigned __int128> row_cells_exact;    // DP cells in rows before r
    void build(const uint32_t *l, uint32_t n) {
        n_seq = n;
        len.assign(l, l + n);
        pref.assign((size_t)n + 1, 0);
        for (uint32_t s = 0; s < n; ++s) pref[s + 1] = pref[s] + len[s];
        row_cells_exact.assign((size_t)n + 1, 0);
        for (uint32_t r = 0; r + 1 < n; ++r)
            row_cells_exact[r + 1] = row_cells_exact[r] + (unsigned __int128)len[r] * (pref[n] - pref[r + 1]);
        if (n) row_cells_exact[n] = row_cells_exact[n - 1];
    }
    uint64_t pairs() const { return n_seq < 2 ? 0 : (uint64_t)n_seq * (n_seq - 1) / 2; }
    // cells of all pairs with triangle index < q
    unsigned __int128 cells_before(uint64_t q) const {
        const uint64_t N = n_seq;
        if (N < 2) return 0;
        if (q >= pairs()) return row_cells_exact[N - 1];
        uint32_t a, b;
        tri_pair(q, (uint32_t)N, a, b);
        return row_cells_exact[a] + (unsigned __int128)len[a] * (pref[b] - pref[a + 1]);
    }
    // n_parts contiguous ranges of [first, first+count) with nearly equal cells
    void partition(uint64_t first, uint64_t count, uint32_t n_parts, uint64_t *bounds) const {
        const unsigned __int128 lo = cells_before(first), hi = cells_before(first + count);
        bounds[0] = first;
        for (uint32_t p = 1; p < n_parts; ++p) {
            const unsigned __int128 target = lo + (hi - lo) * p / n_parts;
            uint64_t a = bounds[p - 1], b = first + count;     // smallest q in [a,b] with cells_before(q) >= target
            while (a < b) {
                const uint64_t mid = a + (b - a) / 2;
                if (cells_before(mid) >= target) b = mid; else a = mid + 1;
            }
            bounds[p] = a;
        }
        bounds[n_parts] = first + count;
    }
};

struct Context {
    std::vector<Device> dev;
    // host copy of what partitioning needs
    uint32_t n_seq = 0;
    std::vector<uint32_t> len;
    Triangle tri;
    uint8_t *host_masks = nullptr;         // the uploaded 4-bit sets, pinned: the devices pack from them (pa_pack_kernel)
    uint64_t host_masks_cap = 0;           // and pa_align_pair_traceback rebuilds strings from them
    std::vector<unsigned long long> host_offsets;     // offsets into host_masks (n_seq + 1)
    std::vector<uint8_t> host_pure, host_fastok;
    std::vector<unsigned long long> row_items;   // pairs-of-pairs work items in rows before r
    std::vector<uint32_t> order;                 // sequence indices, longest first (stable)
    // geometry of the packed set (what a device needs, beside the members above, to take the set over: upload_device_*)
    std::vector<uint32_t> off2, off4;
    uint64_t n_bases = 0;
    size_t words2 = 4, words4 = 4;
    // pa_init brings the devices up on one host thread each: 0 coming up, 1 ready, 2 failed (with its message)
    std::unique_ptr<std::atomic<int>[]> dev_state;
    std::vector<std::string> dev_err;
    std::vector<int> dev_code;                   // status code that goes with dev_err (PA_ENODEVICE: not an sm_100 part)
    std::vector<std::thread> bringup;
    std::atomic<int> occ_ready{0};               // 1: occ holds the occupancies asked on the first device, 2: that failed
    struct Occ { int duo, duo3, duo8, duo_auto, sets, moves_warp, moves_warp_sets, moves_cta, fast, cta, gen, stats; } occ = {};
    bool all_pure = true;
    bool all_fast = true;              // no sequence holds a gap character (everything can run on the s16x2 kernels)
    bool any_sparse = false;           // at least one gap-free sequence has IUPAC ambiguity codes (set form of the s16x2 kernel)
    bool no_amb = false;               // PAIRALIGN_NO_AMB=1: ambiguity codes always take the general int32 kernel (comparison, tests)
    bool force_32bit = false;          // PAIRALIGN_FORCE_32BIT=1: skip the s16x2 kernel (testing / comparison)
    int kduo = 0;                      // strip width of the s16x2 kernel; 0: per work item (duo_pick_k)
    int kduo_forced = 0;               // PAIRALIGN_KDUO=8|12 forces one width (tuning)
    bool no_win = false;               // PAIRALIGN_NO_WIN=1: long pairs go to the int32 kernels as before (tuning, tests)
    uint32_t kduo_mask = DUO_KSET;     // PAIRALIGN_KDUO_SET: bit k set = width k allowed in the per-item choice (tuning)
    int kduo_step_cost = DUO_STEP_COST;          // PAIRALIGN_KDUO_A: per-step overhead of the cost model, in instructions (tuning)
    int duo_minb = 1;                  // PAIRALIGN_DUO_MINB=3: register-capped build of the s16x2 kernel, 3 CTAs per SM (tuning)
    bool file_order = false;           // PAIRALIGN_ITEM_ORDER=file: partners of a work item are neighbours in file order (comparison)
    bool force_cta = false;            // PAIRALIGN_FORCE_CTA=1: every long pair takes a CTA regardless of how many there are
    uint64_t est_long_pairs = 0;       // pairs of the whole triangle with a sequence longer than LONG_LEN
    bool no_cta = false;               // PAIRALIGN_NO_CTA=1: long pairs stay on the one-pair-per-warp kernel (comparison)
    uint32_t max_len = 0, min_len = 0;
    pa_timing timing = {};
};

std::mutex g_mu;
Context *g_ctx = nullptr;

void free_device(Device &d) {
    if (d.id < 0) return;
    cudaSetDevice(d.id);
    cudaFree(d.p2); cudaFree(d.p4); cudaFree(d.off2); cudaFree(d.off4); cudaFree(d.len); cudaFree(d.pure);
    cudaFree(d.fastok); cudaFree(d.order); cudaFree(d.duo_items);
    cudaFree(d.counters); cudaFree(d.n_deferred); cudaFree(d.deferred); cudaFree(d.deferred2); cudaFree(d.deferred3); cudaFree(d.bbuf);
    cudaFree(d.row_items); cudaFree(d.raw); cudaFree(d.raw_off);
    for (auto &c : d.ce) { for (auto &e : c.k) if (e) cudaEventDestroy(e); if (c.done) cudaEventDestroy(c.done); }
    for (int k = 0; k < 2; ++k) { cudaFree(d.d_out[k]); if (d.h_stage[k]) cudaFreeHost(d.h_stage[k]); }
    cudaFree(d.d_ia); cudaFree(d.d_ib);
    cudaFree(d.d_dirs); cudaFree(d.d_ops); cudaFree(d.d_dirs_off); cudaFree(d.d_ops_off); cudaFree(d.d_nops); cudaFree(d.d_res); cudaFree(d.d_items);
    for (auto &e : d.ev) if (e) cudaEventDestroy(e);
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    d = Device();
}

SeqStore store_of(const Device &d, uint32_t n_seq) {
    SeqStore s;
    s.p2 = d.p2; s.p4 = d.p4; s.off2 = d.off2; s.off4 = d.off4; s.len = d.len; s.pure = d.pure; s.n_seq = n_seq;
    s.fastok = d.fastok;
    return s;
}

// The PRMT byte tables need every score (and score+GO) to fit a signed byte and
// the pad fixed point needs GO <= 0, GE <= 0; anything else runs on the general kernel.
bool fast_params_ok(const pa_params &p) {
    auto fits = [](long long v) { return v >= -128 && v <= 127; };
    return p.gap_open <= 0 && p.gap_ext <= 0 && p.gap_open >= -127 && fits(p.match) && fits(p.mismatch) &&
           fits((long long)p.match + p.gap_open) && fits((long long)p.mismatch + p.gap_open);
}

unsigned __int128 cells_before(const Context &c, uint64_t q) { return c.tri.cells_before(q); }

int ensure_out(Device &d, size_t n) {
    if (n <= d.d_out_cap) return PA_OK;
    for (int k = 0; k < 2; ++k) {
        if (d.d_out[k]) cudaFree(d.d_out[k]);
        if (d.h_stage[k]) cudaFreeHost(d.h_stage[k]);
        d.d_out[k] = nullptr; d.h_stage[k] = nullptr;
    }
    d.d_out_cap = 0;
    for (int k = 0; k < 2; ++k) {
        CU(cudaMalloc(&d.d_out[k], n * sizeof(pa_pair_result)));
        CU(cudaMallocHost(&d.h_stage[k], n * sizeof(pa_pair_result)));
    }
    d.d_out_cap = n;
    return PA_OK;
}

int ensure_deferred(Device &d, size_t n) {
    if (n <= d.deferred_cap) return PA_OK;
    cudaFree(d.deferred); cudaFree(d.deferred2); cudaFree(d.deferred3);
    d.deferred = d.deferred2 = d.deferred3 = nullptr; d.deferred_cap = 0;
    CU(cudaMalloc(&d.deferred, n * sizeof(uint32_t)));
    CU(cudaMalloc(&d.deferred2, n * sizeof(uint32_t)));
    CU(cudaMalloc(&d.deferred3, n * sizeof(uint32_t)));
    d.deferred_cap = n;
    return PA_OK;
}

// Longest sequence the s16x2 kernel may take with plain 16-bit scores, and the bias B it stores them with
// (stored = true + B, see duo_row in pa_dp.cuh): true values lie in [-(|ge| L + |mismatch| + |go|), match L], and
// the stored ones must stay negative (<= -16) and, after one more gap extension, above -32768.  256 of margin.
uint32_t max_len16(const pa_params &p, int *bias = nullptr) {
    const long long m = p.match > 0 ? p.match : 1;
    const long long ge = p.gap_ext < 0 ? -(long long)p.gap_ext : 1;
    const long long fixed = std::llabs((long long)p.mismatch) + std::llabs((long long)p.gap_open);
    long long r = (32752 - 256 - fixed - ge) / (m + ge);
    if (r < 0) r = 0;
    r = std::min<long long>(r, 8192);
    if (bias) *bias = (int)(-16 - m * r);
    return (uint32_t)r;
}

// Launch the kernels for `count` elements (triangle range starting at `first`,
// or the explicit lists ia/ib) writing records to d_out (device).  Asynchronous
// on d.stream; events ev[0..3] bracket the two DP kernels.
// pairalign -a, pairs with IUPAC codes or '-': the general kernel keeping its moves (2 bits per slot, K = 8)
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
pa_general_dirs_kernel(const SeqStore S, const Scoring sc, const uint32_t *ia, const uint32_t *ib, const uint64_t count,
                       unsigned long long *work_counter, int4 *bbuf_all, const uint32_t bbuf_rows,
                       pa_pair_result *out, uint8_t *dirs, const unsigned long long *dirs_off, const int all_pairs) {
    const int lane = threadIdx.x & 31;
    const uint32_t gw = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    int4 *bbuf = bbuf_all + (size_t)gw * bbuf_rows;
    for (;;) {
        unsigned long long e = 0;
        if (lane == 0) e = atomicAdd(work_counter, 1ull);
        e = __shfl_sync(FULL_MASK, e, 0);
        if (e >= count) break;
        const uint32_t a = ia[e], b = ib[e];
        const int n = (int)S.len[a], m = (int)S.len[b];
        if (n == 0 || m == 0) {
            if (lane == 0) { pa_pair_result o; o.score = INT_MIN; o.dist = 0; o.len = 0; o.end_i = n - 1; o.end_j = m - 1; out[e] = o; }
            continue;
        }
        if (!all_pairs && S.pure[a] && S.pure[b]) continue;       // the A/C/G/T kernel has it
        align_warp<KGEN, true, true>(S.p4 + S.off4[a], n, S.p4 + S.off4[b], m, sc, bbuf, &out[e], lane, dirs + dirs_off[e]);
    }
}

// Packing on the device (pa_upload_sequences): one CTA per sequence, one thread per 16 bases -> one 2-bit word
// (A0 G1 C2 T3, placeholder 0 for anything else) and two 4-bit words; every word of the sequence's 16-byte-aligned
// slot is written, so nothing has to be cleared first.  pure[s] = every set of s has exactly one bit.
__global__ void __launch_bounds__(128)
pa_pack_kernel(const uint8_t *raw, const unsigned long long *raw_off, const uint32_t *len, const uint32_t *off2,
               const uint32_t *off4, const uint32_t n_seq, uint32_t *p2, uint32_t *p4, uint8_t *pure) {
    __shared__ int s_bad;
    for (uint32_t s = blockIdx.x; s < n_seq; s += gridDim.x) {
        if (threadIdx.x == 0) s_bad = 0;
        __syncthreads();
        const uint32_t L = len[s];
        const uint8_t *src = raw + raw_off[s];
        const uint32_t n2 = ((L + 15) / 16 + 3) / 4 * 4, n4 = ((L + 7) / 8 + 3) / 4 * 4;
        uint32_t *o2 = p2 + off2[s], *o4 = p4 + off4[s];
        int bad = 0;
        for (uint32_t u = threadIdx.x; u < n2 || 2 * u < n4; u += blockDim.x) {
            uint32_t w2 = 0, w4lo = 0, w4hi = 0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t pos = u * 16 + k;
                if (pos < L) {
                    const uint32_t v = src[pos] & 15u;
                    const uint32_t code = (v == 2) ? 1u : (v == 4) ? 2u : (v == 8) ? 3u : 0u;
                    bad |= (v != 1 && v != 2 && v != 4 && v != 8);
                    w2 |= code << (2 * k);
                    if (k < 8) w4lo |= v << (4 * k); else w4hi |= v << (4 * (k - 8));
                }
            }
            if (u < n2) o2[u] = w2;
            if (2 * u < n4) o4[2 * u] = w4lo;
            if (2 * u + 1 < n4) o4[2 * u + 1] = w4hi;
        }
        if (bad) s_bad = 1;
        __syncthreads();
        if (threadIdx.x == 0) pure[s] = s_bad ? 0 : 1;
        __syncthreads();
    }
}

// The s16x2 work items of the triangle range [first, first+count): rows a_lo..a_hi, the first from partner b_lo and the
// last up to partner b_hi; every row contributes ceil(partners / 2) items (pa_duo_items_kernel fills them in).
struct ItemRange { uint32_t a_lo, b_lo, a_hi, b_hi, items_first; uint64_t n_items; };
ItemRange item_range(const Context &c, uint64_t first, uint64_t count) {
    ItemRange r;
    tri_pair(first, c.n_seq, r.a_lo, r.b_lo);
    tri_pair(first + count - 1, c.n_seq, r.a_hi, r.b_hi);
    if (r.a_lo == r.a_hi) {
        r.items_first = (r.b_hi - r.b_lo + 2) / 2;
        r.n_items = r.items_first;
    } else {
        r.items_first = (c.n_seq - r.b_lo + 1) / 2;
        r.n_items = (uint64_t)r.items_first + (c.row_items[r.a_hi] - c.row_items[r.a_lo + 1]) + (r.b_hi - r.a_hi + 1) / 2;
    }
    return r;
}

// Launch the kernels for `count` elements (triangle range starting at `first`,
// or the explicit lists ia/ib) writing records to d_out (device).  Asynchronous
// on d.stream.  Stages (each later stage takes what the previous one deferred,
// its item count read from device memory):
//   1. s16x2 kernel, two pairs per warp      triangle ranges, A/C/G/T, len <= max_len16
//   2. 32-bit 2-bit kernel, one pair per warp longer A/C/G/T pairs, explicit pair lists
//   3. general kernel                         IUPAC sets, '-', any scoring parameters
// Events: ev[0]..ev[1] stage 1, ev[1]..ev_mid stage 2, ev[2]..ev[3] stage 3.
int launch_chunk(Context &c, Device &d, const pa_params &p, uint64_t first, uint64_t count,
                 const uint32_t *d_ia, const uint32_t *d_ib, pa_pair_result *d_out, Device::ChunkEvents &E) {
    const SeqStore S = store_of(d, c.n_seq);
    Scoring sc{p.match, p.mismatch, p.gap_open, p.gap_ext};
    PairSource src{first, d_ia, d_ib, nullptr};
    CU(cudaMemsetAsync(d.counters, 0, 5 * sizeof(unsigned long long), d.stream));
    CU(cudaMemsetAsync(d.n_deferred, 0, 3 * sizeof(unsigned int), d.stream));
    E.duo = E.fast = E.cta = E.gen = false;
    const int threads = WARPS_PER_CTA * 32;
    if (p.aligned) {
        E.duo = true;
        CU(cudaEventRecord(E.k[0], d.stream));
        pa_aligned_stats_kernel<<<d.grid_stats, threads, 0, d.stream>>>(S, src, count, d.counters, d_out);
        CU(cudaGetLastError());
        for (int k = 1; k < 5; ++k) CU(cudaEventRecord(E.k[k], d.stream));
        d.launches += 1;
        return PA_OK;
    }
    const bool fast = fast_params_ok(p);
    int bias16 = 0;
    const uint32_t l16 = fast ? max_len16(p, &bias16) : 0;
    sc.bias16 = bias16;
    const bool duo = fast && !d_ia && l16 >= 16 && !c.force_32bit;
    // a CTA per pair pays off when there are too few long pairs to keep every warp of a one-item-per-warp kernel
    // busy; with thousands of them those kernels are the faster ones (no hand-over, no block-count rounding)
    const bool route_long = c.max_len > LONG_LEN && !c.no_cta &&
                            (c.force_cta || c.est_long_pairs < 4ull * (uint64_t)d.grid_fast * WARPS_PER_CTA);
    // pairs too long for plain 16-bit scores stay on the s16x2 kernel (floating-window variant) unless they go to the
    // CTA kernel; the edge rows then take two bbuf entries each
    // (a gap extension beyond -1024 could wrap a 16-bit half before the maximum with the opening term is taken)
    // sequences with IUPAC ambiguity codes (no gap character): a second launch, the 4-bit-set form of the s16x2 kernel
    // (pa_dp_sets.cuh), takes their items; it adds (match - mismatch) where two sets intersect, so that must be >= 0
    const bool amb = duo && c.kduo == 0 && c.any_sparse && !c.no_amb && p.match >= p.mismatch;
    // ... and the values a lane holds at one time (13 columns, two rows, what its neighbour hands over) must fit the
    // window around its right edge with the storage bias in place: about 4000 either side of the re-base band
    // (stored values stay in [-32768 + |ge|, -16]); neighbouring states differ by at most one of each penalty
    const long long win_spread = 13ll * (std::llabs((long long)p.match) + std::llabs((long long)p.mismatch) +
                                         std::llabs((long long)p.gap_open) + std::llabs((long long)p.gap_ext));
    const bool win = duo && c.kduo == 0 && !c.no_win && c.max_len > l16 && !route_long && p.gap_ext >= -1024 &&
                     win_spread <= 3500 && (uint64_t)d.bbuf_rows >= 2ull * ((uint64_t)c.max_len + 1);
    int rc = ensure_deferred(d, (size_t)count);
    if (rc) return rc;
    CU(cudaEventRecord(E.k[0], d.stream));
    bool stage2 = false, stage_cta = false, stage_gen = false;
    PairSource src2 = src;
    const unsigned int *count2 = nullptr;
    if (duo) {
        const ItemRange ir = item_range(c, first, count);
        if (ir.n_items > d.cap_duo_items) {       // only ever grows (cudaFree waits for the chunk in flight)
            cudaFree(d.duo_items); d.duo_items = nullptr; d.cap_duo_items = 0;
            const size_t want = (size_t)std::max<uint64_t>(ir.n_items, std::min<uint64_t>(CHUNK_PAIRS, c.tri.pairs()) / 2 + c.n_seq + 2);
            CU(cudaMalloc(&d.duo_items, want * sizeof(uint4)));
            d.cap_duo_items = want;
        }
        pa_duo_items_kernel<<<(unsigned)std::min<uint32_t>(ir.a_hi - ir.a_lo + 1, 8u * (uint32_t)d.n_sm), 256, 0, d.stream>>>(
            d.order, c.n_seq, ir.a_lo, ir.b_lo, ir.a_hi, ir.b_hi, d.row_items, ir.items_first, d.duo_items);
        CU(cudaGetLastError());
        d.launches += 1;
        const uint4 *items = d.duo_items;
        const uint64_t n_items = ir.n_items;
        if (c.kduo == 0 && p.gap_ext == -1)      // pairalign's own gap extension: the build with GE as an immediate
            pa_warp_duo_kernel<0, 1, -1><<<d.grid_duo_auto, threads, 0, d.stream>>>(
                S, sc, first, items, n_items, l16, d.counters, d.bbuf, d.bbuf_rows, d_out,
                d.deferred, d.n_deferred, c.kduo_mask, c.kduo_step_cost, win ? 1 : 0, amb ? 1 : 0);
        else if (c.kduo == 0)
            pa_warp_duo_kernel<0><<<d.grid_duo_auto, threads, 0, d.stream>>>(
                S, sc, first, items, n_items, l16, d.counters, d.bbuf, d.bbuf_rows, d_out,
                d.deferred, d.n_deferred, c.kduo_mask, c.kduo_step_cost, win ? 1 : 0, amb ? 1 : 0);
        else if (c.kduo == 8)
            pa_warp_duo_kernel<8><<<d.grid_duo8, threads, 0, d.stream>>>(
                S, sc, first, items, n_items, l16, d.counters, d.bbuf, d.bbuf_rows, d_out,
                d.deferred, d.n_deferred);
        else if (c.duo_minb == 3)
            pa_warp_duo_kernel<KDUO, 3><<<d.grid_duo3, threads, 0, d.stream>>>(
                S, sc, first, items, n_items, l16, d.counters, d.bbuf, d.bbuf_rows, d_out,
                d.deferred, d.n_deferred);
        else
            pa_warp_duo_kernel<KDUO><<<d.grid_duo, threads, 0, d.stream>>>(
                S, sc, first, items, n_items, l16, d.counters, d.bbuf, d.bbuf_rows, d_out,
                d.deferred, d.n_deferred);
        CU(cudaGetLastError());
        d.launches += 1;
        if (amb) {   // the items with an ambiguous sequence: same work items, 4-bit-set variant (its own work counter)
            if (p.gap_ext == -1 && p.match - p.mismatch == 12)
                pa_warp_sets_kernel<-1, 12><<<d.grid_sets, threads, 0, d.stream>>>(
                    S, sc, first, items, n_items, l16, d.counters + 4, d.bbuf, d.bbuf_rows, d_out, win ? 1 : 0);
            else
                pa_warp_sets_kernel<0><<<d.grid_sets, threads, 0, d.stream>>>(
                    S, sc, first, items, n_items, l16, d.counters + 4, d.bbuf, d.bbuf_rows, d_out, win ? 1 : 0);
            CU(cudaGetLastError());
            d.launches += 1;
        }
        E.duo = true;
        stage2 = !(amb ? c.all_fast : c.all_pure) || (c.max_len > l16 && !win);       // something may have been deferred
        src2.idx = d.deferred;
        count2 = d.n_deferred;
    } else if (fast) {
        stage2 = true;
    }
    CU(cudaEventRecord(E.k[1], d.stream));
    if (stage2) {
        pa_warp32_kernel<KFAST><<<d.grid_fast, threads, 0, d.stream>>>(
            S, sc, src2, count, count2, d.counters + 1, d.bbuf, d.bbuf_rows, d_out, d.deferred2, d.n_deferred + 1,
            route_long ? d.deferred3 : nullptr, d.n_deferred + 2, LONG_LEN);
        CU(cudaGetLastError());
        d.launches += 1;
        E.fast = true;
        stage_cta = route_long;
        stage_gen = !c.all_pure;
    }
    CU(cudaEventRecord(E.k[2], d.stream));
    if (stage_cta) {
        PairSource src3 = src;
        src3.idx = d.deferred3;
        pa_cta32_kernel<KFAST><<<d.grid_cta, CTA_WARPS * 32, 0, d.stream>>>(
            S, sc, src3, d.n_deferred + 2, d.counters + 3, d.bbuf, d.bbuf_rows, d_out);
        CU(cudaGetLastError());
        d.launches += 1;
        E.cta = true;
    }
    CU(cudaEventRecord(E.k[3], d.stream));
    if (!fast) {
        pa_warp_dp_kernel<KGEN, true><<<d.grid_gen, threads, 0, d.stream>>>(
            S, sc, src, count, nullptr, d.counters + 2, d.bbuf, d.bbuf_rows, d_out, nullptr, nullptr);
        CU(cudaGetLastError());
        d.launches += 1;
        E.gen = true;
    } else if (stage_gen) {
        PairSource src4 = src;
        src4.idx = d.deferred2;
        pa_warp_dp_kernel<KGEN, true><<<d.grid_gen, threads, 0, d.stream>>>(
            S, sc, src4, count, d.n_deferred + 1, d.counters + 2, d.bbuf, d.bbuf_rows, d_out, nullptr, nullptr);
        CU(cudaGetLastError());
        d.launches += 1;
        E.gen = true;
    }
    CU(cudaEventRecord(E.k[4], d.stream));
    return PA_OK;
}

// Kernel times of a finished chunk (its events are complete once `done` / k[4] has been waited for).
int collect_chunk_times(Device &d, const Device::ChunkEvents &E, bool with_d2h) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, E.k[0], E.k[1]));
    if (E.duo) d.duo_ms += ms;
    CU(cudaEventElapsedTime(&ms, E.k[1], E.k[2]));
    if (E.fast) d.fast_ms += ms;
    CU(cudaEventElapsedTime(&ms, E.k[2], E.k[3]));
    if (E.cta) d.cta_ms += ms;
    CU(cudaEventElapsedTime(&ms, E.k[3], E.k[4]));
    if (E.gen) d.gen_ms += ms;
    if (with_d2h) {
        CU(cudaEventElapsedTime(&ms, E.k[4], E.done));
        d.d2h_ms += ms;
    }
    return PA_OK;
}

// Run one device's share [first, first+count) of the triangle (or of the
// explicit lists) and deliver records to host memory `out` (out[0] is element
// `first`).  Two chunks are in flight: the kernels of chunk k+1 are enqueued
// right behind those of chunk k on d.stream (no host synchronisation in
// between), chunk k's records travel to pinned memory on copy_stream while
// chunk k+1 computes, and the host copies chunk k-1 into the caller's buffer
// meanwhile.  The host only ever waits for the chunk two launches back.
int run_range(Context &c, Device &d, const pa_params &p, uint64_t first, uint64_t count,
              const uint32_t *h_ia, const uint32_t *h_ib, pa_pair_result *out, pa_pair_result *d_resident) {
    CU(cudaSetDevice(d.id));
    d.duo_ms = d.fast_ms = d.cta_ms = d.gen_ms = d.d2h_ms = d.h2d_ms = 0;
    d.launches = 0;
    if (count == 0) return PA_OK;
    const uint64_t chunk = std::min<uint64_t>(CHUNK_PAIRS, count);
    int rc;
    if (!d_resident) { rc = ensure_out(d, (size_t)chunk); if (rc) return rc; }
    if (h_ia) {
        if (count > d.d_pairs_cap) {
            cudaFree(d.d_ia); cudaFree(d.d_ib); d.d_ia = d.d_ib = nullptr; d.d_pairs_cap = 0;
            CU(cudaMalloc(&d.d_ia, count * sizeof(uint32_t)));
            CU(cudaMalloc(&d.d_ib, count * sizeof(uint32_t)));
            d.d_pairs_cap = count;
        }
        CU(cudaEventRecord(d.ev[4], d.stream));
        CU(cudaMemcpyAsync(d.d_ia, h_ia, count * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.d_ib, h_ib, count * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
        CU(cudaEventRecord(d.ev[5], d.stream));
        CU(cudaStreamSynchronize(d.stream));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, d.ev[4], d.ev[5]));
        d.h2d_ms += ms;
    }
    uint64_t done = 0;
    int slot = 0;
    struct Pending { bool active = false; uint64_t off = 0, n = 0; } pend[2];
    // wait for the chunk in `s`, read its timings, hand its records to the caller
    auto drain = [&](int s) -> int {
        if (!pend[s].active) return PA_OK;
        Device::ChunkEvents &E = d.ce[s];
        CU(cudaEventSynchronize(d_resident ? E.k[4] : E.done));
        int r = collect_chunk_times(d, E, !d_resident);
        if (r) return r;
        if (!d_resident) memcpy(out + pend[s].off, d.h_stage[s], pend[s].n * sizeof(pa_pair_result));
        pend[s].active = false;
        return PA_OK;
    };
    while (done < count) {
        const uint64_t nthis = std::min<uint64_t>(chunk, count - done);
        pa_pair_result *dst = d_resident ? d_resident + done : d.d_out[slot];
        rc = drain(slot);            // the chunk two launches back: its events, d_out and h_stage are free again
        if (rc) return rc;
        Device::ChunkEvents &E = d.ce[slot];
        rc = launch_chunk(c, d, p, first + done, nthis, h_ia ? d.d_ia + done : nullptr, h_ia ? d.d_ib + done : nullptr, dst, E);
        if (rc) return rc;
        if (!d_resident) {
            CU(cudaStreamWaitEvent(d.copy_stream, E.k[4], 0));
            CU(cudaMemcpyAsync(d.h_stage[slot], dst, nthis * sizeof(pa_pair_result), cudaMemcpyDeviceToHost, d.copy_stream));
            CU(cudaEventRecord(E.done, d.copy_stream));
        }
        pend[slot].active = true; pend[slot].off = done; pend[slot].n = nthis;
        done += nthis;
        slot ^= 1;
    }
    rc = drain(slot); if (rc) return rc;        // older chunk first
    rc = drain(slot ^ 1); if (rc) return rc;
    CU(cudaStreamSynchronize(d.stream));
    return PA_OK;
}

int check_params(const pa_params *p) {
    if (!p) return fail(PA_EINVAL, "params is NULL");
    return PA_OK;
}

}  // namespace

// ============================================================================
// for the other translation units of the module (pa_nj.cu)
void pa_internal_set_error(const char *msg) { g_err = msg; }

extern "C" {

int pa_api_version(void) { return PA_API_VERSION; }
const char *pa_last_error(void) { return g_err.c_str(); }

// Tear a context down: its bring-up threads first (they write into c->dev), then the devices.
static void destroy_context(Context *c) {
    if (!c) return;
    for (auto &t : c->bringup) if (t.joinable()) t.join();
    for (auto &d : c->dev) free_device(d);
    if (c->host_masks) cudaFreeHost(c->host_masks);
    delete c;
}

// Resident CTAs per SM of every kernel: asked once, on the first device (the query makes the driver load the kernel for that
// device, a few hundred milliseconds for the whole module; the devices of a context are the same chip).
static cudaError_t query_occupancy(Context::Occ &o) {
    int occ = 0, occ_c = 0;
    cudaError_t e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp_duo_kernel<KDUO>, WARPS_PER_CTA * 32, 0);
    o.duo = std::max(1, occ);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp_duo_kernel<KDUO, 3>, WARPS_PER_CTA * 32, 0);
    o.duo3 = std::max(1, occ);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp_duo_kernel<8>, WARPS_PER_CTA * 32, 0);
    o.duo8 = std::max(1, occ);
    // the two builds of a kernel (gap extension at run time / as an immediate) share one grid size: the smaller
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp_duo_kernel<0>, WARPS_PER_CTA * 32, 0);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, pa_warp_duo_kernel<0, 1, -1>, WARPS_PER_CTA * 32, 0);
    o.duo_auto = std::max(1, std::min(occ, occ_c));
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp_sets_kernel<0>, WARPS_PER_CTA * 32, 0);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, pa_warp_sets_kernel<-1, 12>, WARPS_PER_CTA * 32, 0);
    o.sets = std::max(1, std::min(occ, occ_c));
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp_duo_moves_kernel<0>, WARPS_PER_CTA * 32, 0);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, pa_warp_duo_moves_kernel<-1>, WARPS_PER_CTA * 32, 0);
    o.moves_warp = std::max(1, std::min(occ, occ_c));
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp_duo_moves_kernel<0, true, 0>, WARPS_PER_CTA * 32, 0);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, pa_warp_duo_moves_kernel<-1, true, 12>, WARPS_PER_CTA * 32, 0);
    o.moves_warp_sets = std::max(1, std::min(occ, occ_c));
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_cta_duo_moves_kernel<0>, MOVES_CTA_WARPS * 32, moves_cta_smem(false));
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, pa_cta_duo_moves_kernel<-1>, MOVES_CTA_WARPS * 32, moves_cta_smem(false));
    o.moves_cta = std::max(1, std::min(occ, occ_c));
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp32_kernel<KFAST>, WARPS_PER_CTA * 32, 0);
    o.fast = std::max(1, occ);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_cta32_kernel<KFAST>, CTA_WARPS * 32, 0);
    o.cta = std::max(1, occ);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_warp_dp_kernel<KGEN, true>, WARPS_PER_CTA * 32, 0);
    o.gen = std::max(1, occ);
    if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pa_aligned_stats_kernel, WARPS_PER_CTA * 32, 0);
    o.stats = std::max(1, occ);
    return e2;
}

// One device from nothing to ready, on the calling thread: primary context (most of a second: the driver creates them one
// after the other whatever the number of host threads), attributes, the occupancy table (first device only; the others
// wait for it), streams, events and scratch.  Ends by publishing the device's state.
static void bring_up_device(Context *c, size_t k) {
    Device &d = c->dev[k];
    char msg[256] = "";
    int code = PA_ECUDA;
    auto done = [&](bool ok) {
        if (!ok) { c->dev_err[k] = msg; c->dev_code[k] = code; }
        if (k == 0) c->occ_ready.store(ok ? 1 : 2, std::memory_order_release);
        c->dev_state[k].store(ok ? 1 : 2, std::memory_order_release);
    };
    int major = 0, minor = 0, n_sm = 0;
    cudaError_t e = cudaSetDevice(d.id);
    if (e == cudaSuccess) e = cudaFree(nullptr);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d.id);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, d.id);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, d.id);
    if (e != cudaSuccess) { snprintf(msg, sizeof msg, "cannot select device %d: %s", d.id, cudaGetErrorString(e)); done(false); return; }
    if (major < 10) {
        snprintf(msg, sizeof msg, "device %d is sm_%d%d; this module is built for sm_100a only", d.id, major, minor);
        code = PA_ENODEVICE;
        done(false);
        return;
    }
    d.n_sm = n_sm;
    if (k == 0) {
        e = query_occupancy(c->occ);
        if (e != cudaSuccess) { snprintf(msg, sizeof msg, "occupancy query failed: %s", cudaGetErrorString(e)); done(false); return; }
        c->occ_ready.store(1, std::memory_order_release);
    } else {
        int st;
        while ((st = c->occ_ready.load(std::memory_order_acquire)) == 0) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        if (st != 1) { snprintf(msg, sizeof msg, "device %d given up: the first device failed", d.id); done(false); return; }
    }
    const Context::Occ &o = c->occ;
    e = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking);
    for (auto &ev : d.ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    for (auto &ce : d.ce) {
        for (auto &ev : ce.k) if (e == cudaSuccess) e = cudaEventCreate(&ev);
        if (e == cudaSuccess) e = cudaEventCreate(&ce.done);
    }
    if (e == cudaSuccess) e = cudaMalloc(&d.counters, 5 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&d.n_deferred, 3 * sizeof(unsigned int));
    d.grid_duo = o.duo * d.n_sm; d.grid_duo3 = o.duo3 * d.n_sm; d.grid_duo8 = o.duo8 * d.n_sm;
    d.grid_duo_auto = o.duo_auto * d.n_sm; d.grid_sets = o.sets * d.n_sm;
    d.grid_moves_warp = o.moves_warp * d.n_sm; d.grid_moves_warp_sets = o.moves_warp_sets * d.n_sm; d.grid_moves_cta = o.moves_cta * d.n_sm;
    d.grid_fast = o.fast * d.n_sm; d.grid_cta = o.cta * d.n_sm; d.grid_gen = o.gen * d.n_sm; d.grid_stats = o.stats * d.n_sm;
    d.n_warps = (uint32_t)std::max(std::max(std::max(std::max(std::max(std::max(d.grid_duo, d.grid_sets), d.grid_duo_auto), d.grid_duo3), d.grid_duo8), std::max(std::max(d.grid_fast, std::max(d.grid_moves_warp, d.grid_moves_warp_sets)), d.grid_gen)) * WARPS_PER_CTA, std::max(d.grid_cta, d.grid_moves_cta));
    if (e != cudaSuccess) { snprintf(msg, sizeof msg, "device %d setup failed: %s", d.id, cudaGetErrorString(e)); done(false); return; }
    done(true);
}

int pa_init(const int *devices, int n_dev) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx) { destroy_context(g_ctx); g_ctx = nullptr; }
    int avail = 0;
    cudaError_t e = cudaGetDeviceCount(&avail);
    if (e != cudaSuccess || avail <= 0)
        return fail(PA_ENODEVICE, "no CUDA device available (%s); pairalign_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    std::vector<int> ids;
    if (!devices || n_dev <= 0) ids.push_back(0);
    else for (int k = 0; k < n_dev; ++k) ids.push_back(devices[k]);
    for (int id : ids)
        if (id < 0 || id >= avail) return fail(PA_EINVAL, "device %d out of range (0..%d)", id, avail - 1);
    Context *c = new Context();
    c->dev.resize(ids.size());
    for (size_t k = 0; k < ids.size(); ++k) c->dev[k].id = ids[k];
    c->dev_state.reset(new std::atomic<int>[ids.size()]);
    for (size_t k = 0; k < ids.size(); ++k) c->dev_state[k].store(0);
    c->dev_err.assign(ids.size(), std::string());
    c->dev_code.assign(ids.size(), PA_OK);
    // Every device on its own host thread, side by side.  Measured on an 8-GPU box (profiles/r02_cli_init_8gpu.txt): the
    // first CUDA call of the process costs ~5 s there whatever follows, a context 0.1-0.4 s; bringing the other devices
    // up in the background while the first one already computes (tried) has nothing to overlap.
    for (size_t k = 1; k < ids.size(); ++k) c->bringup.emplace_back(bring_up_device, c, k);
    bring_up_device(c, 0);
    for (auto &t : c->bringup) if (t.joinable()) t.join();
    for (size_t k = 0; k < ids.size(); ++k) {
        if (c->dev_state[k].load(std::memory_order_acquire) == 1) continue;
        const std::string msg = c->dev_err[k];
        const int code = c->dev_code[k];
        destroy_context(c);
        return fail(code, "%s", msg.c_str());
    }
    if (const char *f = std::getenv("PAIRALIGN_FORCE_32BIT")) c->force_32bit = (f[0] == '1');
    if (const char *f = std::getenv("PAIRALIGN_KDUO")) c->kduo_forced = std::atoi(f);
    if (const char *f = std::getenv("PAIRALIGN_KDUO_SET")) c->kduo_mask = (uint32_t)std::strtoul(f, nullptr, 0);
    if (const char *f = std::getenv("PAIRALIGN_KDUO_A")) c->kduo_step_cost = std::atoi(f);
    if (const char *f = std::getenv("PAIRALIGN_NO_CTA")) c->no_cta = (f[0] == '1');
    if (const char *f = std::getenv("PAIRALIGN_NO_WIN")) c->no_win = (f[0] == '1');
    if (const char *f = std::getenv("PAIRALIGN_NO_AMB")) c->no_amb = (f[0] == '1');
    if (const char *f = std::getenv("PAIRALIGN_DUO_MINB")) c->duo_minb = std::atoi(f);
    if (const char *f = std::getenv("PAIRALIGN_FORCE_CTA")) c->force_cta = (f[0] == '1');
    if (const char *f = std::getenv("PAIRALIGN_ITEM_ORDER")) c->file_order = (f[0] == 'f');
    g_ctx = c;
    return PA_OK;
}

void pa_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) return;
    destroy_context(g_ctx);
    g_ctx = nullptr;
}

int pa_device_count(void) { return g_ctx ? (int)g_ctx->dev.size() : 0; }

int pa_visible_devices(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ---- front end --------------------------------------------------------------
int pa_char_to_mask(unsigned char ch) {
    // bit0 A, bit1 G, bit2 C, bit3 T; unions for the IUPAC codes; '-' empty; N and '.' everything
    static int8_t tab[256];
    static bool init = false;
    if (!init) {
        for (int k = 0; k < 256; ++k) tab[k] = -2;
        tab[' '] = tab['\n'] = tab['\r'] = tab['\t'] = -1;
        const char *codes = "-AGRCMSVTWKDYHBN";   // index = bit set
        for (int v = 0; v < 16; ++v) {
            tab[(unsigned char)codes[v]] = (int8_t)v;
            if (codes[v] >= 'A' && codes[v] <= 'Z') tab[(unsigned char)(codes[v] - 'A' + 'a')] = (int8_t)v;
        }
        tab['.'] = 15;
        init = true;
    }
    return tab[ch];
}

char pa_mask_to_char(uint8_t mask) {
    // ascending char order in the reference's map puts '-' before letters and '.' before 'N'
    static const char codes[17] = "-AGRCMSVTWKDYHB.";
    return codes[mask & 15];
}

size_t pa_encode_sequence(const char *text, size_t len, uint8_t *out, size_t *n_unknown,
                          char *unknown_chars, size_t unknown_cap) {
    size_t n = 0, unk = 0;
    for (size_t k = 1; k < len; ++k) {
        const int v = pa_char_to_mask((unsigned char)text[k]);
        if (v >= 0) out[n++] = (uint8_t)v;
        else if (v == -2) { if (unknown_chars && unk < unknown_cap) unknown_chars[unk] = text[k]; ++unk; }
    }
    if (n_unknown) *n_unknown = unk;
    return n;
}

// One device's copy of the current sequence set, from the host's copies in the context: raw sets + geometry, packing on the
// device (pa_pack_kernel); all asynchronous on d.stream.  upload_device_end adds what the host scan produced and waits.
static int upload_device_begin(Context &c, Device &d) {
    auto grow = [](void **ptr, size_t &cap, size_t bytes) -> cudaError_t {
        // device buffers only grow: re-uploading a set of the same size costs copies, not allocations
        if (bytes <= cap && *ptr) return cudaSuccess;
        cudaFree(*ptr);
        *ptr = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(ptr, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    };
    const uint32_t n_seq = c.n_seq;
    const size_t nidx = std::max<size_t>(n_seq, 1);
    CU(cudaSetDevice(d.id));
    CU(grow((void **)&d.row_items, d.cap_row_items, c.row_items.size() * sizeof(unsigned long long)));
    CU(grow((void **)&d.p2, d.cap_p2, c.words2 * 4));
    CU(grow((void **)&d.p4, d.cap_p4, c.words4 * 4));
    CU(grow((void **)&d.off2, d.cap_off2, nidx * 4));
    CU(grow((void **)&d.off4, d.cap_off4, nidx * 4));
    CU(grow((void **)&d.len, d.cap_len, nidx * 4));
    CU(grow((void **)&d.pure, d.cap_pure, nidx));
    CU(grow((void **)&d.fastok, d.cap_fastok, nidx));
    CU(grow((void **)&d.order, d.cap_order, nidx * 4));
    CU(grow((void **)&d.raw, d.cap_raw, (size_t)std::max<uint64_t>(c.n_bases, 16)));
    CU(grow((void **)&d.raw_off, d.cap_raw_off, (nidx + 1) * sizeof(unsigned long long)));
    d.bbuf_rows = std::max<uint32_t>(c.max_len, 1) + 1;   // + the virtual-column row of the s16x2 kernel
    if (c.max_len > 4096) d.bbuf_rows *= 2;               // floating-window s16x2 variant: (values, offsets) per row
    CU(grow((void **)&d.bbuf, d.cap_bbuf, (size_t)d.n_warps * d.bbuf_rows * sizeof(int4)));
    CU(cudaMemcpyAsync(d.row_items, c.row_items.data(), c.row_items.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, d.stream));
    if (n_seq) {
        if (c.n_bases) CU(cudaMemcpyAsync(d.raw, c.host_masks, (size_t)c.n_bases, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.raw_off, c.host_offsets.data(), ((size_t)n_seq + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.off2, c.off2.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.off4, c.off4.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.len, c.len.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.order, c.order.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, d.stream));
        pa_pack_kernel<<<(unsigned)std::min<uint32_t>(n_seq, 65535u * 16u), 128, 0, d.stream>>>(
            d.raw, d.raw_off, d.len, d.off2, d.off4, n_seq, d.p2, d.p4, d.pure);
        CU(cudaGetLastError());
    } else {
        CU(cudaMemsetAsync(d.p2, 0, 16, d.stream));
        CU(cudaMemsetAsync(d.p4, 0, 16, d.stream));
    }
    return PA_OK;
}

static int upload_device_end(Context &c, Device &d) {
    CU(cudaSetDevice(d.id));
    // d.pure was written by pa_pack_kernel from the same bytes; the host's copy is what launch decisions use
    if (c.n_seq) CU(cudaMemcpyAsync(d.fastok, c.host_fastok.data(), (size_t)c.n_seq, cudaMemcpyHostToDevice, d.stream));
    CU(cudaStreamSynchronize(d.stream));
    return PA_OK;
}

int pa_upload_sequences(const uint8_t *masks, const uint64_t *offsets, uint32_t n_seq) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) return fail(PA_ENODEVICE, "pa_init() has not succeeded");
    if (!offsets || (!masks && n_seq && offsets[n_seq] > 0)) return fail(PA_EINVAL, "NULL sequence buffer");
    Context &c = *g_ctx;
    std::vector<uint32_t> len(n_seq), off2(n_seq), off4(n_seq);
    uint64_t w2 = 0, w4 = 0;
    uint32_t max_len = 0;
    for (uint32_t s = 0; s < n_seq; ++s) {
        if (offsets[s + 1] < offsets[s]) return fail(PA_EINVAL, "offsets not ascending at %u", s);
        const uint64_t L = offsets[s + 1] - offsets[s];
        if (L > PA_MAX_SEQ_LEN) return fail(PA_ERANGE, "sequence %u has %llu bases (max %u)", s, (unsigned long long)L, PA_MAX_SEQ_LEN);
        len[s] = (uint32_t)L;
        max_len = std::max(max_len, len[s]);
        off2[s] = (uint32_t)w2; off4[s] = (uint32_t)w4;
        w2 += ((L + 15) / 16 + 3) / 4 * 4;     // whole 16-byte units
        w4 += ((L + 7) / 8 + 3) / 4 * 4;
        if (w4 > 0xffffffffull) return fail(PA_ERANGE, "sequence set too large");
    }
    const uint64_t n_bases = n_seq ? offsets[n_seq] - offsets[0] : 0;
    const uint64_t base0 = n_seq ? offsets[0] : 0;
    int rc;
    // The host keeps the sets (pa_align_pair_traceback rebuilds strings from them) in pinned memory, which is also
    // where the devices fetch them from.  Packing is the devices' job (pa_pack_kernel): the host only scans.
    if (n_bases > c.host_masks_cap) {
        if (c.host_masks) cudaFreeHost(c.host_masks);
        c.host_masks = nullptr; c.host_masks_cap = 0;
        CU(cudaSetDevice(c.dev[0].id));
        CU(cudaMallocHost(&c.host_masks, (size_t)std::max<uint64_t>(n_bases, 4096)));
        c.host_masks_cap = std::max<uint64_t>(n_bases, 4096);
    }
    if (n_bases) memcpy(c.host_masks, masks + base0, (size_t)n_bases);
    c.host_offsets.resize((size_t)n_seq + 1);
    for (uint32_t s = 0; s <= n_seq; ++s) c.host_offsets[s] = offsets[s] - base0;
    c.row_items.assign((size_t)n_seq + 1, 0);
    for (uint32_t r = 0; r < n_seq; ++r) c.row_items[r + 1] = c.row_items[r] + ((uint64_t)(n_seq - 1 - r) + 1) / 2;
    c.order.resize(n_seq);
    for (uint32_t s = 0; s < n_seq; ++s) c.order[s] = s;
    if (!c.file_order) std::stable_sort(c.order.begin(), c.order.end(), [&](uint32_t x, uint32_t y) { return len[x] > len[y]; });
    c.n_seq = n_seq;
    c.len = len;
    c.off2 = off2; c.off4 = off4;
    c.n_bases = n_bases;
    c.words2 = (size_t)std::max<uint64_t>(w2, 4); c.words4 = (size_t)std::max<uint64_t>(w4, 4);
    c.max_len = max_len;
    c.min_len = n_seq ? *std::min_element(len.begin(), len.end()) : 0;
    for (auto &d : c.dev) {
        rc = upload_device_begin(c, d);
        if (rc) return rc;
    }

    // Meanwhile on the host: which sequences are plain A/C/G/T and which of the others are free of gap characters
    // (branch-free scans the compiler vectorises).
    std::vector<uint8_t> pure(n_seq), fastok(n_seq);
    bool all_fast = true, all_pure = true, any_sparse = false;
    for (uint32_t s = 0; s < n_seq; ++s) {
        const uint8_t *src = c.host_masks + c.host_offsets[s];
        const uint32_t L = len[s];
        uint8_t bad = 0;
        for (uint32_t k = 0; k < L; ++k) {
            const uint8_t v = src[k] & 15u;
            bad |= (uint8_t)((v & (uint8_t)(v - 1u)) | (uint8_t)(v == 0));     // not exactly one bit set
        }
        const bool pr = bad == 0;
        pure[s] = pr ? 1 : 0;
        all_pure = all_pure && pr;
        // an ambiguous sequence stays on the s16x2 path (4-bit-set variant) unless it holds a gap character: '-' scores
        // INT_MIN with 32-bit wrap-around in the reference (src/seqpair.cpp:192-193), which only the int32 kernel reproduces
        bool ok = true;
        if (!pr) {
            uint8_t gap = 0;
            for (uint32_t k = 0; k < L; ++k) gap |= (uint8_t)((src[k] & 15u) == 0);
            ok = gap == 0;
            any_sparse = any_sparse || ok;
        }
        fastok[s] = ok ? 1 : 0;
        all_fast = all_fast && ok;
    }
    c.host_pure = pure;
    c.host_fastok = fastok;
    c.all_pure = all_pure;
    c.all_fast = all_fast;
    c.any_sparse = any_sparse;
    c.tri.build(len.data(), n_seq);
    // strip width of the s16x2 kernel: chosen per work item inside the kernel (0); PAIRALIGN_KDUO=8|12 forces one width
    c.kduo = (c.kduo_forced == 8 || c.kduo_forced == 12) ? c.kduo_forced : 0;
    {
        uint64_t n_long = 0;
        for (uint32_t s = 0; s < n_seq; ++s) n_long += len[s] > LONG_LEN ? 1 : 0;
        c.est_long_pairs = n_seq ? n_long * (uint64_t)(n_seq - 1) - n_long * (n_long ? n_long - 1 : 0) / 2 : 0;
    }
    for (auto &d : c.dev) {
        rc = upload_device_end(c, d);
        if (rc) return rc;
    }
    return PA_OK;
}

uint32_t pa_num_sequences(void) { return g_ctx ? g_ctx->n_seq : 0; }

uint64_t pa_num_pairs(void) {
    if (!g_ctx) return 0;
    const uint64_t n = g_ctx->n_seq;
    return n < 2 ? 0 : n * (n - 1) / 2;
}

int pa_pair_from_index(uint64_t k, uint32_t *a, uint32_t *b) {
    if (!g_ctx) return fail(PA_ENODEVICE, "pa_init() has not succeeded");
    if (k >= pa_num_pairs() || !a || !b) return fail(PA_EINVAL, "pair index out of range");
    tri_pair(k, g_ctx->n_seq, *a, *b);
    return PA_OK;
}

uint64_t pa_count_cells(uint64_t first, uint64_t count) {
    if (!g_ctx) return 0;
    const uint64_t total = pa_num_pairs();
    if (first > total) first = total;
    if (count > total - first) count = total - first;
    return (uint64_t)(cells_before(*g_ctx, first + count) - cells_before(*g_ctx, first));
}

int pa_partition_pairs(uint64_t first, uint64_t count, uint32_t n_parts, uint64_t *bounds) {
    if (!g_ctx) return fail(PA_ENODEVICE, "pa_init() has not succeeded");
    if (!bounds || n_parts == 0) return fail(PA_EINVAL, "bad partition arguments");
    const uint64_t total = pa_num_pairs();
    if (first > total || count > total - first) return fail(PA_ERANGE, "pair range outside the triangle");
    g_ctx->tri.partition(first, count, n_parts, bounds);
    return PA_OK;
}

int pa_partition_by_length(const uint32_t *lengths, uint32_t n_seq, uint64_t first, uint64_t count, uint32_t n_parts,
                           uint64_t *bounds, uint64_t *cells) {
    if (!lengths || !bounds || n_parts == 0) return fail(PA_EINVAL, "bad partition arguments");
    Triangle t;
    t.build(lengths, n_seq);
    const uint64_t total = t.pairs();
    if (first > total || count > total - first) return fail(PA_ERANGE, "pair range outside the triangle");
    t.partition(first, count, n_parts, bounds);
    if (cells)
        for (uint32_t p = 0; p < n_parts; ++p) cells[p] = (uint64_t)(t.cells_before(bounds[p + 1]) - t.cells_before(bounds[p]));
    return PA_OK;
}

static int ops_range(Context &c, Device &d, const pa_params &prm, const uint32_t *ia, const uint32_t *ib, uint64_t lo, uint64_t hi,
                     uint8_t *ops, const uint64_t *op_offsets, uint32_t *n_ops, pa_pair_result *res, double *kernel_ms,
                     pa_pair_result *d_res_out);

// A triangle range of LONG A/C/G/T pairs that is too small to fill the one-item-per-warp statistics kernel (config 5 cut
// over 4 or 8 GPUs: a few hundred 30 kb x 30 kb items per device, each 0.8 s of one warp) runs a CTA per item instead:
// the move-storing s16x2 kernel of pairalign -a (2.35 TCUPS) and the walk, which counts the statistics; no op strings.
static bool few_long_items(const Context &c, const pa_params &p, uint64_t count_per_device, const Device &d) {
    if (c.no_cta || c.force_32bit || p.aligned || !c.all_pure || c.min_len <= LONG_LEN || !fast_params_ok(p)) return false;
    const long long win_spread = 17ll * (std::llabs((long long)p.match) + std::llabs((long long)p.mismatch) +
                                         std::llabs((long long)p.gap_open) + std::llabs((long long)p.gap_ext));
    if (p.gap_ext < -1024 || win_spread > 3500 || max_len16(p) < 16) return false;
    // measured on config 5 (profiles/r02_c5_routes.txt): 4975 items per device are better off on the warp kernel (4.2 per warp), 2487 are not
    return c.force_cta || count_per_device / 2 < 4ull * (uint64_t)d.grid_duo_auto * WARPS_PER_CTA;
}

static int align_impl(const pa_params *params, uint64_t first, uint64_t count, const uint32_t *ia, const uint32_t *ib,
                      pa_pair_result *out, pa_pair_result *d_resident) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) return fail(PA_ENODEVICE, "pa_init() has not succeeded");
    int rc = check_params(params);
    if (rc) return rc;
    Context &c = *g_ctx;
    if (!c.dev[0].p4) return fail(PA_EINVAL, "no sequences uploaded");
    if (!out && !d_resident && count) return fail(PA_EINVAL, "NULL result buffer");
    if (!ia) {
        const uint64_t total = pa_num_pairs();
        if (first > total || count > total - first) return fail(PA_ERANGE, "pair range outside the triangle");
    } else {
        for (uint64_t k = 0; k < count; ++k)
            if (ia[k] >= c.n_seq || ib[k] >= c.n_seq) return fail(PA_ERANGE, "pair %llu names a sequence out of range", (unsigned long long)k);
    }
    const auto t0 = std::chrono::steady_clock::now();
    const size_t nd = d_resident ? 1 : c.dev.size();
    std::vector<uint64_t> bounds(nd + 1);
    if (!ia) {
        if (nd == 1) { bounds[0] = first; bounds[1] = first + count; }
        else {
            c.tri.partition(first, count, (uint32_t)nd, bounds.data());
        }
    } else {
        for (size_t p = 0; p <= nd; ++p) bounds[p] = count * p / nd;
    }
    std::vector<int> rcs(nd, PA_OK);
    std::vector<std::string> errs(nd);
    const bool via_moves = !ia && count && few_long_items(c, *params, (count + nd - 1) / nd, c.dev[0]);
    std::vector<double> kms(2 * nd, 0.0);
    auto work = [&](size_t p) {
        const uint64_t lo = bounds[p], n = bounds[p + 1] - bounds[p];
        Device &d = c.dev[p];
        if (via_moves) {
            d.duo_ms = d.fast_ms = d.cta_ms = d.gen_ms = d.d2h_ms = d.h2d_ms = 0;
            d.launches = 0;
            std::vector<uint32_t> la((size_t)n), lb((size_t)n);
            uint32_t a = 0, b = 0;
            if (n) tri_pair(lo, c.n_seq, a, b);
            for (uint64_t k = 0; k < n; ++k) { la[(size_t)k] = a; lb[(size_t)k] = b; if (++b == c.n_seq) { ++a; b = a + 1; } }
            rcs[p] = ops_range(c, d, *params, la.data(), lb.data(), 0, n, nullptr, nullptr, nullptr, out ? out + (lo - first) : nullptr,
                               &kms[2 * p], d_resident);
        }
        else if (!ia) rcs[p] = run_range(c, d, *params, lo, n, nullptr, nullptr, out ? out + (lo - first) : nullptr, d_resident);
        else rcs[p] = run_range(c, d, *params, 0, n, ia + lo, ib + lo, out + lo, nullptr);
        if (rcs[p]) errs[p] = g_err;
    };
    if (nd == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (size_t p = 0; p < nd; ++p) th.emplace_back(work, p);
        for (auto &t : th) t.join();
    }
    for (size_t p = 0; p < nd; ++p) if (rcs[p]) { g_err = errs[p]; return rcs[p]; }
    const auto t1 = std::chrono::steady_clock::now();
    pa_timing &tm = c.timing;
    tm = pa_timing();
    for (size_t p = 0; p < nd; ++p) {
        const Device &d = c.dev[p];
        tm.kernel_ms = std::max(tm.kernel_ms, d.duo_ms + d.fast_ms + d.cta_ms + d.gen_ms + kms[2 * p + 1]);
        tm.walk_ms = std::max(tm.walk_ms, kms[2 * p + 1]);
        tm.dp_cta_ms = std::max(tm.dp_cta_ms, d.cta_ms);
        tm.dp_duo_ms = std::max(tm.dp_duo_ms, d.duo_ms);
        tm.dp_fast_ms = std::max(tm.dp_fast_ms, d.fast_ms);
        tm.dp_general_ms = std::max(tm.dp_general_ms, d.gen_ms);
        tm.d2h_ms = std::max(tm.d2h_ms, d.d2h_ms);
        tm.h2d_ms = std::max(tm.h2d_ms, d.h2d_ms);
        tm.kernel_launches += d.launches;
    }
    tm.total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    tm.pairs = count;
    tm.n_devices = (uint32_t)nd;
    if (!ia) tm.cells = (uint64_t)(cells_before(c, first + count) - cells_before(c, first));
    else { uint64_t cells = 0; for (uint64_t k = 0; k < count; ++k) cells += (uint64_t)c.len[ia[k]] * c.len[ib[k]]; tm.cells = cells; }
    return PA_OK;
}

int pa_align_all_pairs(const pa_params *params, uint64_t first, uint64_t count, pa_pair_result *out) {
    return align_impl(params, first, count, nullptr, nullptr, out, nullptr);
}

int pa_align_pairs(const pa_params *params, const uint32_t *ia, const uint32_t *ib, uint64_t count, pa_pair_result *out) {
    if (count && (!ia || !ib)) return fail(PA_EINVAL, "NULL pair list");
    if (count == 0) return PA_OK;
    return align_impl(params, 0, count, ia, ib, out, nullptr);
}

int pa_align_all_pairs_device(const pa_params *params, uint64_t first, uint64_t count, void *d_out) {
    if (!d_out && count) return fail(PA_EINVAL, "NULL device buffer");
    return align_impl(params, first, count, nullptr, nullptr, nullptr, (pa_pair_result *)d_out);
}

// Bytes of the move store of one pair: n rows of P*W/4 bytes (W = 32*K slots per block), 128-byte aligned.
static uint64_t dirs_bytes(uint32_t n, uint32_t m, bool moves_kernel) {
    const uint64_t W = 32ull * (moves_kernel ? KMOV : KGEN);
    const uint64_t P = (m + W - 1) / W;
    const uint64_t bytes = (uint64_t)n * (P * W / 4);
    return (bytes + 127) / 128 * 128;
}

// pairalign -a on one device: pairs [lo, hi) of the caller's list in batches sized to the move store (2 bits per
// cell).  A/C/G/T pairs run on the s16x2 move-storing kernels (pa_dp_moves.cuh), two list entries that share their
// first sequence per work item; pairs longer than LONG_LEN take a CTA per item when a batch holds too few of them to
// fill a warp-per-item grid -- which the size of their move stores (226 MB for 30 kb x 30 kb) all but guarantees.
// Everything else (IUPAC codes, gap characters, scoring outside the byte tables) runs on the general int32 kernel.
// ops == nullptr: statistics only -- no op strings are made or copied, the records (score and end cell from the DP,
// compared / differing columns from the walk) go to res (host, indexed like ia) and / or d_res_out (device, element lo first).
static int ops_range(Context &c, Device &d, const pa_params &prm, const uint32_t *ia, const uint32_t *ib, uint64_t lo, uint64_t hi,
                     uint8_t *ops, const uint64_t *op_offsets, uint32_t *n_ops, pa_pair_result *res, double *kernel_ms,
                     pa_pair_result *d_res_out) {
    if (lo >= hi) return PA_OK;
    CU(cudaSetDevice(d.id));
    int bias16 = 0;
    const uint32_t l16 = fast_params_ok(prm) ? max_len16(prm, &bias16) : 0;
    const long long win_spread = 17ll * (std::llabs((long long)prm.match) + std::llabs((long long)prm.mismatch) +
                                         std::llabs((long long)prm.gap_open) + std::llabs((long long)prm.gap_ext));
    const bool win_ok = prm.gap_ext >= -1024 && win_spread <= 3500;
    // the s16x2 move kernels take the A/C/G/T pairs of this call (else: everything on the general kernel)
    const bool fast = l16 >= 16 && (c.max_len <= l16 || win_ok) && !c.force_32bit;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t budget = std::min<uint64_t>((uint64_t)((free_b + d.cap_dirs) / 2), 72ull << 30);
    Scoring sc{prm.match, prm.mismatch, prm.gap_open, prm.gap_ext};
    sc.bias16 = bias16;
    const SeqStore S = store_of(d, c.n_seq);
    const int threads = WARPS_PER_CTA * 32;
    std::vector<unsigned long long> h_dirs_off, h_ops_off;
    std::vector<uint2> h_items[2][2];            // [long][sets]: work items of the four s16x2 move kernels
    const uint64_t wave_items = (uint64_t)d.grid_moves_cta;              // one wave of the CTA-per-item grid
    // which kernel family takes a pair: 0 the general int32 kernel, 1 the s16x2 move kernels on 2-bit codes (both plain
    // A/C/G/T), 2 the same on 4-bit sets (IUPAC codes, no gap character)
    const bool sets_ok = prm.match >= prm.mismatch && !c.no_amb;
    auto klass = [&](uint32_t a, uint32_t b) -> int {
        if (!fast || !c.len[a] || !c.len[b]) return 0;
        if (c.host_pure[a] && c.host_pure[b]) return 1;
        return (sets_ok && c.host_fastok[a] && c.host_fastok[b]) ? 2 : 0;
    };
    uint64_t s0 = lo;
    while (s0 < hi) {
        // one batch: as many pairs as the move store holds
        uint64_t e0 = s0, dbytes = 0, obytes = 0, n_long = 0, n_long_items = 0;
        uint64_t ck_e0 = s0, ck_dbytes = 0, ck_obytes = 0;      // the batch as it was after the last whole wave of long items
        bool any_general = false;
        bool open = false, open_long = false;                   // the last entry began a work item that may still take a partner
        uint32_t open_a = 0;
        h_dirs_off.clear(); h_ops_off.clear();
        while (e0 < hi && e0 - s0 < (1ull << 20)) {
            const uint32_t a = ia[e0], b = ib[e0];
            const int kl = klass(a, b);
            const uint64_t need = (c.len[a] && c.len[b]) ? dirs_bytes(c.len[a], c.len[b], kl != 0) : 0;
            if (dbytes + need > budget) {
                if (e0 == s0) return fail(PA_ENOMEM, "the moves of pair %llu need %llu bytes; %llu available", (unsigned long long)e0,
                                          (unsigned long long)need, (unsigned long long)budget);
                // a batch of long pairs runs one item per CTA: end it on a whole number of waves of the persistent grid
                if (n_long == e0 - s0 && ck_e0 > s0 && !c.no_cta) {
                    e0 = ck_e0; dbytes = ck_dbytes; obytes = ck_obytes;
                    h_dirs_off.resize(e0 - s0); h_ops_off.resize(e0 - s0);
                }
                break;
            }
            h_dirs_off.push_back(dbytes);
            h_ops_off.push_back(obytes);
            dbytes += need;
            if (ops) obytes += (uint64_t)c.len[a] + c.len[b];
            if (kl) {      // same pairing rule as the item loop below: neighbours with the same first sequence
                const bool lng = std::max(c.len[a], c.len[b]) > LONG_LEN;
                if (lng) ++n_long;
                if (open && open_a == a && open_long == lng) open = false;
                else { open = true; open_a = a; open_long = lng; if (lng) ++n_long_items; }
            } else {
                open = false;
                if (need) any_general = true;
            }
            ++e0;
            if (n_long == e0 - s0 && n_long_items % wave_items == 0) { ck_e0 = e0; ck_dbytes = dbytes; ck_obytes = obytes; }
        }
        const uint64_t nb = e0 - s0;
        // work items of the s16x2 kernels: within a run of neighbouring entries with the same first sequence (and the same
        // side of LONG_LEN) the entries go together in twos in order of the second sequence's length -- both pairs of an
        // item sweep the passes of the longer one; an item with an ambiguous sequence runs on the set form
        for (auto &row : h_items) for (auto &v : row) v.clear();
        std::vector<uint32_t> run;
        for (uint64_t k = 0; k < nb;) {
            const uint32_t a = ia[s0 + k];
            if (!klass(a, ib[s0 + k])) { ++k; continue; }
            const bool lng = std::max(c.len[a], c.len[ib[s0 + k]]) > LONG_LEN;
            run.clear();
            uint64_t k1 = k;
            for (; k1 < nb && ia[s0 + k1] == a; ++k1) {
                const uint32_t b = ib[s0 + k1];
                if (!klass(a, b) || (std::max(c.len[a], c.len[b]) > LONG_LEN) != lng) break;
                run.push_back((uint32_t)k1);
            }
            if (!c.file_order)
                std::stable_sort(run.begin(), run.end(), [&](uint32_t x, uint32_t y) { return c.len[ib[s0 + x]] > c.len[ib[s0 + y]]; });
            for (size_t j = 0; j < run.size(); j += 2) {
                const bool two = j + 1 < run.size();
                int kl = klass(a, ib[s0 + run[j]]);
                if (two) kl = std::max(kl, klass(a, ib[s0 + run[j + 1]]));
                h_items[lng ? 1 : 0][kl == 2 ? 1 : 0].push_back(make_uint2(run[j], two ? run[j + 1] : 0xffffffffu));
            }
            k = k1;
        }
        const size_t n_long_it = h_items[1][0].size() + h_items[1][1].size();
        const bool route_long = n_long_it && !c.no_cta && (c.force_cta || n_long_it < 4ull * (uint64_t)d.grid_moves_warp * WARPS_PER_CTA);
        if (!route_long)
            for (int z = 0; z < 2; ++z) { h_items[0][z].insert(h_items[0][z].end(), h_items[1][z].begin(), h_items[1][z].end()); h_items[1][z].clear(); }
        size_t item_off[2][2], n_items_all = 0;
        for (int l = 0; l < 2; ++l) for (int z = 0; z < 2; ++z) { item_off[l][z] = n_items_all; n_items_all += h_items[l][z].size(); }
        auto grow = [](void **ptr, size_t &cap, size_t bytes) -> cudaError_t {
            if (bytes <= cap && *ptr) return cudaSuccess;
            cudaFree(*ptr); *ptr = nullptr; cap = 0;
            cudaError_t e = cudaMalloc(ptr, bytes);
            if (e == cudaSuccess) cap = bytes;
            return e;
        };
        CU(grow((void **)&d.d_dirs, d.cap_dirs, std::max<uint64_t>(dbytes, 128)));
        CU(grow((void **)&d.d_ops, d.cap_ops, std::max<uint64_t>(obytes, 128)));
        CU(grow((void **)&d.d_items, d.cap_items, std::max<size_t>(n_items_all, 1) * sizeof(uint2)));
        if (nb > d.cap_tb_pairs) {
            cudaFree(d.d_dirs_off); cudaFree(d.d_ops_off); cudaFree(d.d_nops); cudaFree(d.d_res);
            d.d_dirs_off = d.d_ops_off = nullptr; d.d_nops = nullptr; d.d_res = nullptr; d.cap_tb_pairs = 0;
            CU(cudaMalloc(&d.d_dirs_off, nb * sizeof(unsigned long long)));
            CU(cudaMalloc(&d.d_ops_off, nb * sizeof(unsigned long long)));
            CU(cudaMalloc(&d.d_nops, nb * sizeof(uint32_t)));
            CU(cudaMalloc(&d.d_res, nb * sizeof(pa_pair_result)));
            d.cap_tb_pairs = nb;
        }
        if (nb > d.d_pairs_cap) {
            cudaFree(d.d_ia); cudaFree(d.d_ib); d.d_ia = d.d_ib = nullptr; d.d_pairs_cap = 0;
            CU(cudaMalloc(&d.d_ia, nb * sizeof(uint32_t)));
            CU(cudaMalloc(&d.d_ib, nb * sizeof(uint32_t)));
            d.d_pairs_cap = nb;
        }
        CU(cudaMemcpyAsync(d.d_ia, ia + s0, nb * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.d_ib, ib + s0, nb * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.d_dirs_off, h_dirs_off.data(), nb * sizeof(unsigned long long), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemcpyAsync(d.d_ops_off, h_ops_off.data(), nb * sizeof(unsigned long long), cudaMemcpyHostToDevice, d.stream));
        for (int l = 0; l < 2; ++l) for (int z = 0; z < 2; ++z)
            if (!h_items[l][z].empty())
                CU(cudaMemcpyAsync(d.d_items + item_off[l][z], h_items[l][z].data(), h_items[l][z].size() * sizeof(uint2), cudaMemcpyHostToDevice, d.stream));
        CU(cudaMemsetAsync(d.counters, 0, 5 * sizeof(unsigned long long), d.stream));
        CU(cudaMemsetAsync(d.d_res, 0, nb * sizeof(pa_pair_result), d.stream));
        CU(cudaEventRecord(d.ev[0], d.stream));
        const bool gec = prm.gap_ext == -1, dc12 = gec && prm.match - prm.mismatch == 12;
        // warp per item: plain (work counter 0), sets (counter 2)
        if (!h_items[0][0].empty()) {
            const uint32_t ni = (uint32_t)h_items[0][0].size();
            if (gec) pa_warp_duo_moves_kernel<-1><<<d.grid_moves_warp, threads, 0, d.stream>>>(S, sc, d.d_ia, d.d_ib, d.d_items + item_off[0][0], ni, l16,
                                                                                      d.counters, d.bbuf, d.bbuf_rows, d.d_res, d.d_dirs, d.d_dirs_off);
            else pa_warp_duo_moves_kernel<0><<<d.grid_moves_warp, threads, 0, d.stream>>>(S, sc, d.d_ia, d.d_ib, d.d_items + item_off[0][0], ni, l16,
                                                                                 d.counters, d.bbuf, d.bbuf_rows, d.d_res, d.d_dirs, d.d_dirs_off);
            CU(cudaGetLastError());
            d.launches += 1;
        }
        if (!h_items[0][1].empty()) {
            const uint32_t ni = (uint32_t)h_items[0][1].size();
            if (dc12) pa_warp_duo_moves_kernel<-1, true, 12><<<d.grid_moves_warp_sets, threads, 0, d.stream>>>(S, sc, d.d_ia, d.d_ib, d.d_items + item_off[0][1], ni, l16,
                                                                                                       d.counters + 2, d.bbuf, d.bbuf_rows, d.d_res, d.d_dirs, d.d_dirs_off);
            else pa_warp_duo_moves_kernel<0, true, 0><<<d.grid_moves_warp_sets, threads, 0, d.stream>>>(S, sc, d.d_ia, d.d_ib, d.d_items + item_off[0][1], ni, l16,
                                                                                                d.counters + 2, d.bbuf, d.bbuf_rows, d.d_res, d.d_dirs, d.d_dirs_off);
            CU(cudaGetLastError());
            d.launches += 1;
        }
        CU(cudaEventRecord(d.ev[3], d.stream));
        // CTA per item: plain (counter 3), sets (counter 4); more than 48 KB of dynamic shared memory needs the opt-in, once per device
        if (n_long_it && route_long && !d.moves_smem_set) {
            CU(cudaFuncSetAttribute(pa_cta_duo_moves_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)moves_cta_smem(false)));
            CU(cudaFuncSetAttribute(pa_cta_duo_moves_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)moves_cta_smem(false)));
            CU(cudaFuncSetAttribute(pa_cta_duo_moves_kernel<0, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)moves_cta_smem(true)));
            CU(cudaFuncSetAttribute(pa_cta_duo_moves_kernel<-1, true, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)moves_cta_smem(true)));
            d.moves_smem_set = true;
        }
        if (!h_items[1][0].empty()) {
            const uint32_t ni = (uint32_t)h_items[1][0].size();
            if (gec) pa_cta_duo_moves_kernel<-1><<<d.grid_moves_cta, MOVES_CTA_WARPS * 32, moves_cta_smem(false), d.stream>>>(
                         S, sc, d.d_ia, d.d_ib, d.d_items + item_off[1][0], ni, d.counters + 3, d.bbuf, d.bbuf_rows, d.d_res, d.d_dirs, d.d_dirs_off);
            else pa_cta_duo_moves_kernel<0><<<d.grid_moves_cta, MOVES_CTA_WARPS * 32, moves_cta_smem(false), d.stream>>>(
                     S, sc, d.d_ia, d.d_ib, d.d_items + item_off[1][0], ni, d.counters + 3, d.bbuf, d.bbuf_rows, d.d_res, d.d_dirs, d.d_dirs_off);
            CU(cudaGetLastError());
            d.launches += 1;
        }
        if (!h_items[1][1].empty()) {
            const uint32_t ni = (uint32_t)h_items[1][1].size();
            if (dc12) pa_cta_duo_moves_kernel<-1, true, 12><<<d.grid_moves_cta, MOVES_CTA_WARPS * 32, moves_cta_smem(true), d.stream>>>(
                          S, sc, d.d_ia, d.d_ib, d.d_items + item_off[1][1], ni, d.counters + 4, d.bbuf, d.bbuf_rows, d.d_res, d.d_dirs, d.d_dirs_off);
            else pa_cta_duo_moves_kernel<0, true, 0><<<d.grid_moves_cta, MOVES_CTA_WARPS * 32, moves_cta_smem(true), d.stream>>>(
                     S, sc, d.d_ia, d.d_ib, d.d_items + item_off[1][1], ni, d.counters + 4, d.bbuf, d.bbuf_rows, d.d_res, d.d_dirs, d.d_dirs_off);
            CU(cudaGetLastError());
            d.launches += 1;
        }
        CU(cudaEventRecord(d.ev[4], d.stream));
        if (any_general) {
            pa_general_dirs_kernel<<<d.grid_gen, threads, 0, d.stream>>>(S, sc, d.d_ia, d.d_ib, nb, d.counters + 1, d.bbuf, d.bbuf_rows,
                                                                         d.d_res, d.d_dirs, d.d_dirs_off, fast ? 0 : 1);
            CU(cudaGetLastError());
            d.launches += 1;
        }
        CU(cudaEventRecord(d.ev[1], d.stream));
        pa_walk_kernel<<<(unsigned)((nb + WALK_WARPS - 1) / WALK_WARPS), WALK_WARPS * 32, 0, d.stream>>>(S, d.d_ia, d.d_ib, nb, d.d_res, d.d_dirs, d.d_dirs_off,
                                                                        ops ? d.d_ops : nullptr, d.d_ops_off, d.d_nops,
                                                                        fast ? KMOV : KGEN, (fast && sets_ok) ? KMOV : KGEN, KGEN);
        CU(cudaGetLastError());
        d.launches += 1;
        CU(cudaEventRecord(d.ev[2], d.stream));
        if (ops) {
            CU(cudaMemcpyAsync(ops + op_offsets[s0], d.d_ops, obytes, cudaMemcpyDeviceToHost, d.stream));
            CU(cudaMemcpyAsync(n_ops + s0, d.d_nops, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, d.stream));
        }
        if (res) CU(cudaMemcpyAsync(res + s0, d.d_res, nb * sizeof(pa_pair_result), cudaMemcpyDeviceToHost, d.stream));
        if (d_res_out) CU(cudaMemcpyAsync(d_res_out + (s0 - lo), d.d_res, nb * sizeof(pa_pair_result), cudaMemcpyDeviceToDevice, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        float dp_ms = 0, walk_ms = 0, warp_ms = 0, cta_ms = 0;
        CU(cudaEventElapsedTime(&dp_ms, d.ev[0], d.ev[1]));
        CU(cudaEventElapsedTime(&walk_ms, d.ev[1], d.ev[2]));
        CU(cudaEventElapsedTime(&warp_ms, d.ev[0], d.ev[3]));
        CU(cudaEventElapsedTime(&cta_ms, d.ev[3], d.ev[4]));
        kernel_ms[0] += dp_ms; kernel_ms[1] += walk_ms;
        if (!h_items[0][0].empty() || !h_items[0][1].empty()) d.duo_ms += warp_ms;
        if (n_long_it && route_long) d.cta_ms += cta_ms;
        if (any_general) d.gen_ms += dp_ms - warp_ms - cta_ms;
        // the walk wrote each op string backwards (the reference reverses at src/seqpair.cpp:183-188)
        if (ops) for (uint64_t k = s0; k < e0; ++k) std::reverse(ops + op_offsets[k], ops + op_offsets[k] + n_ops[k]);
        s0 = e0;
    }
    return PA_OK;
}

// pa_align_pairs_ops with g_mu already held by the caller
static int ops_impl(const pa_params *params, const uint32_t *ia, const uint32_t *ib, uint64_t count,
                    uint8_t *ops, uint64_t ops_cap, uint64_t *op_offsets, uint32_t *n_ops, pa_pair_result *res) {
    if (!g_ctx) return fail(PA_ENODEVICE, "pa_init() has not succeeded");
    int rc = check_params(params);
    if (rc) return rc;
    if (count == 0) return PA_OK;
    if (!ia || !ib || !ops || !op_offsets || !n_ops) return fail(PA_EINVAL, "NULL buffer");
    Context &c = *g_ctx;
    if (!c.dev[0].p4) return fail(PA_EINVAL, "no sequences uploaded");
    if (params->aligned) return fail(PA_EINVAL, "pairalign -A prints its input: there is nothing to trace back");
    uint64_t total_ops = 0;
    unsigned __int128 total_cells = 0;
    for (uint64_t k = 0; k < count; ++k) {
        if (ia[k] >= c.n_seq || ib[k] >= c.n_seq) return fail(PA_ERANGE, "pair %llu names a sequence out of range", (unsigned long long)k);
        op_offsets[k] = total_ops;
        total_ops += (uint64_t)c.len[ia[k]] + c.len[ib[k]];
        total_cells += (unsigned __int128)c.len[ia[k]] * c.len[ib[k]];
    }
    op_offsets[count] = total_ops;
    if (total_ops > ops_cap) return fail(PA_EINVAL, "op buffer holds %llu bytes, %llu needed (sum of both lengths over the pairs)",
                                         (unsigned long long)ops_cap, (unsigned long long)total_ops);
    // contiguous ranges of the list with nearly equal DP cells, one per device (each on its own host thread);
    // every range writes its own slice of ops / n_ops / res, so nothing is shared
    const auto t0 = std::chrono::steady_clock::now();
    const size_t nd = c.dev.size();
    std::vector<uint64_t> bounds(nd + 1, count);
    bounds[0] = 0;
    if (nd > 1) {
        unsigned __int128 acc = 0;
        size_t p = 1;
        for (uint64_t k = 0; k < count && p < nd; ++k) {
            while (p < nd && acc >= total_cells * p / nd) bounds[p++] = k;
            acc += (unsigned __int128)c.len[ia[k]] * c.len[ib[k]];
        }
    }
    std::vector<int> rcs(nd, PA_OK);
    std::vector<std::string> errs(nd);
    std::vector<double> kms(2 * nd, 0.0);
    for (auto &d : c.dev) { d.duo_ms = d.fast_ms = d.cta_ms = d.gen_ms = d.d2h_ms = d.h2d_ms = 0; d.launches = 0; }
    auto work = [&](size_t p) {
        Device &d = c.dev[p];
        rcs[p] = ops_range(c, d, *params, ia, ib, bounds[p], bounds[p + 1], ops, op_offsets, n_ops, res, &kms[2 * p], nullptr);
        if (rcs[p]) errs[p] = g_err;
    };
    if (nd == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (size_t p = 0; p < nd; ++p) th.emplace_back(work, p);
        for (auto &t : th) t.join();
    }
    for (size_t p = 0; p < nd; ++p) if (rcs[p]) { g_err = errs[p]; return rcs[p]; }
    const auto t1 = std::chrono::steady_clock::now();
    pa_timing &tm = c.timing;
    tm = pa_timing();
    for (size_t p = 0; p < nd; ++p) {
        const Device &d = c.dev[p];
        tm.kernel_ms = std::max(tm.kernel_ms, kms[2 * p] + kms[2 * p + 1]);
        tm.dp_cta_ms = std::max(tm.dp_cta_ms, d.cta_ms);
        tm.dp_duo_ms = std::max(tm.dp_duo_ms, d.duo_ms);
        tm.dp_general_ms = std::max(tm.dp_general_ms, d.gen_ms);
        tm.walk_ms = std::max(tm.walk_ms, kms[2 * p + 1]);
        tm.kernel_launches += d.launches;
    }
    tm.total_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    tm.pairs = count;
    tm.n_devices = (uint32_t)nd;
    tm.cells = (uint64_t)total_cells;
    return PA_OK;
}

int pa_align_pairs_ops(const pa_params *params, const uint32_t *ia, const uint32_t *ib, uint64_t count,
                       uint8_t *ops, uint64_t ops_cap, uint64_t *op_offsets, uint32_t *n_ops, pa_pair_result *res) {
    std::lock_guard<std::mutex> lk(g_mu);
    return ops_impl(params, ia, ib, count, ops, ops_cap, op_offsets, n_ops, res);
}

int pa_align_pair_traceback(const pa_params *params, uint32_t a, uint32_t b, uint8_t *ax, uint8_t *ay, uint32_t cap,
                            uint32_t *alen, pa_pair_result *res) {
    // one lock from the length check to the rendering: a concurrent pa_upload_sequences cannot swap the set in between
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) return fail(PA_ENODEVICE, "pa_init() has not succeeded");
    if (a >= g_ctx->n_seq || b >= g_ctx->n_seq) return fail(PA_ERANGE, "sequence index out of range");
    if (!ax || !ay || !alen) return fail(PA_EINVAL, "NULL output buffer");
    const uint32_t n = g_ctx->len[a], m = g_ctx->len[b];
    if (cap < n + m) return fail(PA_EINVAL, "alignment buffers need n+m = %u bytes", n + m);
    if (n == 0 || m == 0) return fail(PA_EINVAL, "empty sequence: the reference's behaviour is undefined here");
    std::vector<uint8_t> ops((size_t)n + m);
    uint64_t off[2];
    uint32_t nops = 0;
    int rc = ops_impl(params, &a, &b, 1, ops.data(), ops.size(), off, &nops, res);
    if (rc) return rc;
    const uint8_t *x = g_ctx->host_masks + g_ctx->host_offsets[a], *y = g_ctx->host_masks + g_ctx->host_offsets[b];
    uint32_t i = 0, j = 0;
    for (uint32_t k = 0; k < nops; ++k) {
        ax[k] = ops[k] == 2 ? 0 : x[i++];
        ay[k] = ops[k] == 1 ? 0 : y[j++];
    }
    *alen = nops;
    return PA_OK;
}

int pa_s16_limits(const pa_params *params, uint32_t *max_len, int32_t *bias) {
    int rc = check_params(params);
    if (rc) return rc;
    int b = 0;
    const uint32_t l = fast_params_ok(*params) ? max_len16(*params, &b) : 0;
    if (max_len) *max_len = l;
    if (bias) *bias = l ? b : 0;
    return PA_OK;
}

int pa_get_timing(pa_timing *t) {
    if (!g_ctx) return fail(PA_ENODEVICE, "pa_init() has not succeeded");
    if (!t) return fail(PA_EINVAL, "NULL timing");
    *t = g_ctx->timing;
    return PA_OK;
}

// ---- statistics: the reference's expressions, evaluated by the host's libm ----
double pa_similarity(uint32_t dist, uint32_t len) {
    if (len > 0) return 1.0 - ((int)dist / double((int)len));
    return 1.0;
}
double pa_pdistance(uint32_t dist, uint32_t len) { return 1 - pa_similarity(dist, len); }
double pa_jc_distance(uint32_t dist, uint32_t len) {
    const double p = 1 - pa_similarity(dist, len);
    return log(1.0 - (4.0 / 3.0) * p) * (-3.0 / 4.0);
}
double pa_jc_minus_p(uint32_t dist, uint32_t len) { return pa_jc_distance(dist, len) - (1.0 - pa_similarity(dist, len)); }

int pa_int32_peak(int which, double *gops, double *sm_mhz) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) return fail(PA_ENODEVICE, "pa_init() has not succeeded");
    if (!gops) return fail(PA_EINVAL, "NULL output");
    Device &d = g_ctx->dev[0];
    CU(cudaSetDevice(d.id));
    double g = 0, mhz = 0;
    cudaError_t e = pa::run_peak(which, d.n_sm, d.stream, &g, &mhz);
    if (e == cudaErrorInvalidValue) return fail(PA_EINVAL, "unknown instruction class %d", which);
    if (e != cudaSuccess) return fail(PA_ECUDA, "peak kernel failed: %s", cudaGetErrorString(e));
    *gops = g;
    if (sm_mhz) *sm_mhz = mhz;
    return PA_OK;
}

}  // extern "C"
