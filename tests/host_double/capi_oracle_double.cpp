// capi_oracle_double.cpp -- TEST DOUBLE, never part of the product.
//
// Answers the COMPUTE entry points of include/pairalign_b200.h from the oracle (oracle/pa_oracle.c) so that the
// reference's own pairalign.cpp, patched by integration/pairalign_b200.patch, can be run on a machine without a GPU
// (tests/test_integration_patch.py): what is under test there is the patch and integration/b200_batch.h, i.e. the
// collect / align / replay restructuring of cluster().  Linked in front of libpairalign_b200.so, whose host-only
// entry points (pa_encode_sequence, pa_mask_to_char, pa_similarity ...) stay in use.
#include <climits>
#include <cstring>
#include <vector>

#include "../../include/pairalign_b200.h"
#include "../../oracle/pa_oracle.h"

namespace {
std::vector<uint8_t> g_masks;
std::vector<uint64_t> g_off;

void one_pair(const pa_params &p, uint32_t a, uint32_t b, pa_pair_result *out) {
    pa_oracle_result r;
    std::memset(&r, 0, sizeof r);
    const uint8_t *x = g_masks.data() + g_off[a], *y = g_masks.data() + g_off[b];
    const int32_t n = (int32_t)(g_off[a + 1] - g_off[a]), m = (int32_t)(g_off[b + 1] - g_off[b]);
    if (p.aligned) pa_oracle_aligned_stats(x, n, y, m, &r);
    else if (n == 0 || m == 0) { r.score = INT_MIN; r.end_i = n - 1; r.end_j = m - 1; }     // as the CUDA module defines it
    else pa_oracle_align_forward(x, n, y, m, p.match, p.mismatch, p.gap_open, p.gap_ext, &r);
    out->score = r.score; out->dist = r.dist; out->len = r.len; out->end_i = r.end_i; out->end_j = r.end_j;
}
}  // namespace

extern "C" {
int pa_init(const int *, int) { return PA_OK; }
void pa_shutdown(void) {}
int pa_upload_sequences(const uint8_t *masks, const uint64_t *offsets, uint32_t n_seq) {
    g_off.assign(offsets, offsets + n_seq + 1);
    g_masks.assign(masks, masks + offsets[n_seq]);
    return PA_OK;
}
uint64_t pa_num_pairs(void) {
    const uint64_t n = g_off.empty() ? 0 : g_off.size() - 1;
    return n < 2 ? 0 : n * (n - 1) / 2;
}
int pa_align_all_pairs(const pa_params *p, uint64_t first, uint64_t count, pa_pair_result *out) {
    const uint32_t n = (uint32_t)(g_off.size() - 1);
    uint64_t k = 0;
    for (uint32_t a = 0; a + 1 < n; ++a)
        for (uint32_t b = a + 1; b < n; ++b, ++k)
            if (k >= first && k < first + count) one_pair(*p, a, b, out + (k - first));
    return PA_OK;
}
int pa_align_pairs(const pa_params *p, const uint32_t *ia, const uint32_t *ib, uint64_t count, pa_pair_result *out) {
    for (uint64_t k = 0; k < count; ++k) one_pair(*p, ia[k], ib[k], out + k);
    return PA_OK;
}
int pa_align_pairs_ops(const pa_params *p, const uint32_t *ia, const uint32_t *ib, uint64_t count, uint8_t *ops, uint64_t ops_cap,
                       uint64_t *op_offsets, uint32_t *n_ops, pa_pair_result *res) {
    uint64_t total = 0;
    for (uint64_t k = 0; k < count; ++k) {
        op_offsets[k] = total;
        total += (g_off[ia[k] + 1] - g_off[ia[k]]) + (g_off[ib[k] + 1] - g_off[ib[k]]);
    }
    op_offsets[count] = total;
    if (total > ops_cap) return PA_EINVAL;
    for (uint64_t k = 0; k < count; ++k) {
        const uint32_t a = ia[k], b = ib[k];
        const int32_t n = (int32_t)(g_off[a + 1] - g_off[a]), m = (int32_t)(g_off[b + 1] - g_off[b]);
        pa_oracle_result r;
        int32_t alen = 0;
        std::vector<uint8_t> buf((size_t)n + m + 1);
        if (pa_oracle_align_ops(g_masks.data() + g_off[a], n, g_masks.data() + g_off[b], m, p->match, p->mismatch, p->gap_open,
                                p->gap_ext, &r, buf.data(), &alen)) return PA_EINVAL;
        std::memcpy(ops + op_offsets[k], buf.data(), (size_t)alen);
        n_ops[k] = (uint32_t)alen;
        if (res) { res[k].score = r.score; res[k].dist = r.dist; res[k].len = r.len; res[k].end_i = r.end_i; res[k].end_j = r.end_j; }
    }
    return PA_OK;
}
}
