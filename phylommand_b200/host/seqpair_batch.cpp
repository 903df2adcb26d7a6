#include "seqpair_batch.h"

#include <cmath>
#include <cstdlib>
#include <iostream>
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

size_t SeqpairBatch::add_sequence(const std::string &raw) {
    const size_t start = masks_.size();
    masks_.resize(start + raw.size() + 1);
    size_t n_unknown = 0;
    std::string unk(raw.size(), '\0');
    const size_t n = pa_encode_sequence(raw.data(), raw.size(), masks_.data() + start, &n_unknown,
                                        unk.empty() ? nullptr : &unk[0], unk.size());
    masks_.resize(start + n);
    offsets_.push_back(masks_.size());
    unk.resize(n_unknown);
    if (n_unknown) any_unknown_ = true;
    unknown_.push_back(unk);
    return offsets_.size() - 2;
}

void SeqpairBatch::warn_unknown(size_t s) const {
    for (char c : unknown_[s]) std::cerr << "Can not interpret '" << c << "'. Not in alphabet." << std::endl;
}

std::string SeqpairBatch::text(size_t s) const {
    std::string out(length(s), '-');
    const uint8_t *m = masks(s);
    for (size_t k = 0; k < out.size(); ++k) out[k] = pa_mask_to_char(m[k]);
    return out;
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

void SeqpairBatch::render(uint32_t a, uint32_t b, const uint8_t *ops, uint32_t n_ops, std::string &x, std::string &y) const {
    // translate_to_string (src/seqpair.cpp:62-72) through a 16-entry table: this loop writes every character of
    // pairalign -a's output (1.2 GB for 200 x 30 kb)
    static const struct Lut { char c[16]; Lut() { for (int m = 0; m < 16; ++m) c[m] = pa_mask_to_char((uint8_t)m); } } lut;
    const uint8_t *ma = masks(a), *mb = masks(b);
    x.resize(n_ops);
    y.resize(n_ops);
    char *px = &x[0], *py = &y[0];
    size_t i = 0, j = 0;
    for (uint32_t k = 0; k < n_ops; ++k) {
        const uint8_t o = ops[k];
        const bool tx = (o != 2), ty = (o != 1);
        px[k] = tx ? lut.c[ma[i] & 15u] : '-';
        py[k] = ty ? lut.c[mb[j] & 15u] : '-';
        i += tx; j += ty;
    }
}

}  // namespace pab
