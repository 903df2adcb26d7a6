// seqpair_batch_oracle.cpp -- TEST DOUBLE, never part of the product.
//
// Stands in for phylommand_b200/host/seqpair_batch_device.cpp when the command line is built for the CPU test
// suite (tests/test_host_replay_cpu.py): the per-pair records and op strings come from the oracle
// (oracle/pa_oracle.c) instead of the CUDA module, everything else -- FASTA index and order, matrix framing,
// number formatting, single-link clusters, MAD groups, pair-FASTA input -- is the product's own host code,
// compared byte for byte with the reference's output in tests/golden/cli/.  The four C-ABI entry points the
// command line calls directly are answered here as well (the executable's definitions take precedence over the
// library's), so nothing in this binary needs a device.
#include <cstring>
#include <algorithm>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../../oracle/pa_oracle.h"
#include "../../phylommand_b200/host/seqpair_batch.h"

namespace {
uint32_t g_n_seq = 0;

void to_record(const pa_oracle_result &r, pa_pair_result *out) {
    out->score = r.score; out->dist = r.dist; out->len = r.len; out->end_i = r.end_i; out->end_j = r.end_j;
}

void pair_of(uint64_t k, uint32_t n_seq, uint32_t *a, uint32_t *b) {      // row-major upper triangle
    uint32_t r = 0;
    uint64_t row = n_seq - 1;
    while (k >= row) { k -= row; --row; ++r; }
    *a = r; *b = r + 1 + (uint32_t)k;
}
}  // namespace

extern "C" {
int pa_device_count(void) { return 1; }
int pa_get_timing(pa_timing *) { return PA_ENODEVICE; }
void pa_shutdown(void) {}
int pa_pair_from_index(uint64_t k, uint32_t *a, uint32_t *b) {
    if (g_n_seq < 2 || k >= (uint64_t)g_n_seq * (g_n_seq - 1) / 2) return PA_EINVAL;
    pair_of(k, g_n_seq, a, b);
    return PA_OK;
}
}

namespace pab {

void init_devices(double) {}

void SeqpairBatch::upload() { g_n_seq = (uint32_t)size(); }

static void one_pair(const SeqpairBatch &sb, const pa_params &p, uint32_t a, uint32_t b, pa_pair_result *out) {
    pa_oracle_result r;
    std::memset(&r, 0, sizeof r);
    const int32_t n = (int32_t)sb.length(a), m = (int32_t)sb.length(b);
    if (p.aligned) pa_oracle_aligned_stats(sb.masks(a), n, sb.masks(b), m, &r);
    else if (n == 0 || m == 0) { r.score = INT32_MIN; r.end_i = n - 1; r.end_j = m - 1; }      // as the CUDA module defines it
    else if (pa_oracle_align_forward(sb.masks(a), n, sb.masks(b), m, p.match, p.mismatch, p.gap_open, p.gap_ext, &r))
        throw std::runtime_error("oracle failed");
    to_record(r, out);
}

void SeqpairBatch::align_range(const pa_params &p, uint64_t first, uint64_t count, pa_pair_result *out) {
    bool any_empty = false;
    for (size_t s = 0; s < size(); ++s) any_empty = any_empty || length(s) == 0;
    if (!p.aligned && !any_empty && count > 64) {      // the example files: the oracle's own threaded all-pairs driver
        static_assert(sizeof(pa_oracle_result) == sizeof(pa_pair_result), "same five 32-bit fields");
        std::vector<pa_oracle_result> tmp((size_t)count);
        unsigned hw = std::thread::hardware_concurrency();
        if (pa_oracle_all_pairs(masks_.data(), offsets_.data(), (uint32_t)size(), p.match, p.mismatch, p.gap_open, p.gap_ext,
                                first, first + count, (int)std::max(1u, std::min(16u, hw)), tmp.data()))
            throw std::runtime_error("oracle failed");
        for (uint64_t k = 0; k < count; ++k) to_record(tmp[(size_t)k], out + k);
        return;
    }
    for (uint64_t k = 0; k < count; ++k) {
        uint32_t a = 0, b = 0;
        pair_of(first + k, (uint32_t)size(), &a, &b);
        one_pair(*this, p, a, b, out + k);
    }
}

void SeqpairBatch::align_list(const pa_params &p, const std::vector<uint32_t> &ia, const std::vector<uint32_t> &ib,
                              pa_pair_result *out) {
    for (size_t k = 0; k < ia.size(); ++k) one_pair(*this, p, ia[k], ib[k], out + k);
}

void SeqpairBatch::alignments(const pa_params &p, const std::vector<uint32_t> &ia, const std::vector<uint32_t> &ib, OpBatch &out) {
    out.offsets.assign(ia.size() + 1, 0);
    out.n_ops.assign(ia.size(), 0);
    uint64_t cap = 0;
    for (size_t k = 0; k < ia.size(); ++k) { out.offsets[k] = cap; cap += (uint64_t)length(ia[k]) + length(ib[k]); }
    out.offsets[ia.size()] = cap;
    out.ops.assign(cap ? cap : 1, 0);
    for (size_t k = 0; k < ia.size(); ++k) {
        const int32_t n = (int32_t)length(ia[k]), m = (int32_t)length(ib[k]);
        uint8_t *ops = out.ops.data() + out.offsets[k];
        if (n == 0 || m == 0) {                      // nothing to align: the other sequence against gaps
            std::memset(ops, 1, (size_t)n); std::memset(ops + n, 2, (size_t)m);
            out.n_ops[k] = (uint32_t)(n + m);
            continue;
        }
        pa_oracle_result r;
        int32_t alen = 0;
        if (pa_oracle_align_ops(masks(ia[k]), n, masks(ib[k]), m, p.match, p.mismatch, p.gap_open, p.gap_ext, &r, ops, &alen))
            throw std::runtime_error("oracle failed");
        out.n_ops[k] = (uint32_t)alen;
    }
}

void SeqpairBatch::alignment(const pa_params &p, uint32_t a, uint32_t b, std::string &x, std::string &y) {
    if (p.aligned) { x = text(a); y = text(b); return; }       // pairalign -A -a prints the input back
    OpBatch ob;
    alignments(p, std::vector<uint32_t>(1, a), std::vector<uint32_t>(1, b), ob);
    render(a, b, ob.ops.data(), ob.n_ops[0], x, y);
}

}  // namespace pab
