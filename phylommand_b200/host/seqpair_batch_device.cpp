// seqpair_batch_device.cpp -- the half of SeqpairBatch that calls the CUDA module (include/pairalign_b200.h).
// There is no CPU path here: every call either reaches a GPU or throws.
#include "seqpair_batch.h"

#include <cmath>
#include <cstdlib>
#include <stdexcept>

namespace pab {

static void check(int rc, const char *what) {
    if (rc != PA_OK) throw std::runtime_error(std::string(what) + ": " + pa_last_error());
}

void init_devices(double est_cells) {
    std::vector<int> ids;
    if (const char *env = std::getenv("PAIRALIGN_DEVICES")) {
        std::string s(env), tok;
        for (size_t k = 0; k <= s.size(); ++k) {
            if (k == s.size() || s[k] == ',') { if (!tok.empty()) ids.push_back(std::atoi(tok.c_str())); tok.clear(); }
            else tok += s[k];
        }
    } else {
        // A CUDA context costs ~0.3 s per extra device: take one device per ~0.5 s of single-GPU work (1.2e12 DP cells)
        int n = pa_visible_devices();
        if (est_cells >= 0) {
            const double want = std::ceil(est_cells / 1.2e12);
            if (want < n) n = want < 1 ? 1 : (int)want;
        }
        for (int k = 0; k < n; ++k) ids.push_back(k);
    }
    check(pa_init(ids.empty() ? nullptr : ids.data(), (int)ids.size()), "pa_init");
}

void SeqpairBatch::upload() {
    check(pa_upload_sequences(masks_.data(), offsets_.data(), (uint32_t)size()), "pa_upload_sequences");
}

void SeqpairBatch::align_range(const pa_params &p, uint64_t first, uint64_t count, pa_pair_result *out) {
    check(pa_align_all_pairs(&p, first, count, out), "pa_align_all_pairs");
}

void SeqpairBatch::align_list(const pa_params &p, const std::vector<uint32_t> &ia, const std::vector<uint32_t> &ib,
                              pa_pair_result *out) {
    check(pa_align_pairs(&p, ia.data(), ib.data(), ia.size(), out), "pa_align_pairs");
}

void SeqpairBatch::alignment(const pa_params &p, uint32_t a, uint32_t b, std::string &x, std::string &y) {
    const uint32_t n = length(a), m = length(b);
    if (p.aligned) { x = text(a); y = text(b); return; }       // pairalign -A -a prints the input back
    if (n == 0 || m == 0) {
        // the reference indexes outside its (empty) matrices here; what it prints in practice is the
        // non-empty sequence against gaps
        x = n ? text(a) : std::string(m, '-');
        y = m ? text(b) : std::string(n, '-');
        return;
    }
    std::vector<uint8_t> ax(n + m), ay(n + m);
    uint32_t alen = 0;
    check(pa_align_pair_traceback(&p, a, b, ax.data(), ay.data(), n + m, &alen, nullptr), "pa_align_pair_traceback");
    x.assign(alen, '-');
    y.assign(alen, '-');
    for (uint32_t k = 0; k < alen; ++k) { x[k] = pa_mask_to_char(ax[k]); y[k] = pa_mask_to_char(ay[k]); }
}

void SeqpairBatch::alignments(const pa_params &p, const std::vector<uint32_t> &ia, const std::vector<uint32_t> &ib, OpBatch &out) {
    uint64_t cap = 0;
    for (size_t k = 0; k < ia.size(); ++k) cap += (uint64_t)length(ia[k]) + length(ib[k]);
    out.ops.resize(std::max<uint64_t>(cap, 1));
    out.offsets.resize(ia.size() + 1);
    out.n_ops.resize(ia.size());
    if (ia.empty()) return;
    check(pa_align_pairs_ops(&p, ia.data(), ib.data(), ia.size(), out.ops.data(), cap, out.offsets.data(), out.n_ops.data(), nullptr),
          "pa_align_pairs_ops");
}

}  // namespace pab
