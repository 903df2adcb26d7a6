// seqpair_batch.cpp -- the host-only half of SeqpairBatch: encoding, text, rendering of op strings.
// The half that talks to the CUDA module is seqpair_batch_device.cpp (tests/host_double/ swaps that file for one that
// asks the oracle, so that the whole command line -- order, framing, formatting, clustering, MAD -- is also checked
// against the reference's output on a machine without a GPU; the product links the device half and nothing else).
#include "seqpair_batch.h"

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

namespace pab {

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

void SeqpairBatch::render(uint32_t a, uint32_t b, const uint8_t *ops, uint32_t n_ops, std::string &x, std::string &y) const {
    x.resize(n_ops);
    y.resize(n_ops);
    if (n_ops) render_into(a, b, ops, n_ops, &x[0], &y[0]);
}

void SeqpairBatch::render_into(uint32_t a, uint32_t b, const uint8_t *ops, uint32_t n_ops, char *px, char *py) const {
    // translate_to_string (src/seqpair.cpp:62-72) through a 16-entry table: this loop writes every character of
    // pairalign -a's output (1.2 GB for 200 x 30 kb)
    static const struct Lut { char c[16]; Lut() { for (int m = 0; m < 16; ++m) c[m] = pa_mask_to_char((uint8_t)m); } } lut;
    const uint8_t *ma = masks(a), *mb = masks(b);
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
