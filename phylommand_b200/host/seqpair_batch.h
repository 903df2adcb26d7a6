// seqpair_batch.h -- the reworked seqpair front end: what the reference does per
// pair in seqpair's constructor and align() (src/seqpair.h:53-60,
// src/seqpair.cpp:74-190) is done here once per sequence (encoding) and once per
// batch of pairs (the CUDA module behind include/pairalign_b200.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "pairalign_b200.h"

namespace pab {

// The per-pair view the emitters use; mirrors the seqpair getters
// (src/seqpair.h:83-99) on top of one pa_pair_result record.
struct PairStats {
    pa_pair_result r;
    double similarity() const { return pa_similarity(r.dist, r.len); }
    double proportion_different() const { return pa_pdistance(r.dist, r.len); }
    double jc_distance() const { return pa_jc_distance(r.dist, r.len); }
    double jc_minus_p() const { return pa_jc_minus_p(r.dist, r.len); }
};

class SeqpairBatch {
public:
    // translate_to_binary of one raw sequence (first character dropped, blanks and
    // unknown characters skipped); returns its index
    size_t add_sequence(const std::string &raw);
    size_t size() const { return offsets_.size() - 1; }
    uint32_t length(size_t s) const { return (uint32_t)(offsets_[s + 1] - offsets_[s]); }
    const uint8_t *masks(size_t s) const { return masks_.data() + offsets_[s]; }
    // the reference warns "Can not interpret 'c'. Not in alphabet." every time a
    // sequence is encoded, i.e. once per pair it takes part in (src/seqpair.cpp:86)
    void warn_unknown(size_t s) const;
    bool any_unknown() const { return any_unknown_; }
    std::string text(size_t s) const;          // translate_to_string (src/seqpair.cpp:62-72)

    // push everything to the devices; throws std::runtime_error on failure
    void upload();
    // all pairs [first, first+count) of the upper triangle in reference order
    void align_range(const pa_params &p, uint64_t first, uint64_t count, pa_pair_result *out);
    void align_list(const pa_params &p, const std::vector<uint32_t> &ia, const std::vector<uint32_t> &ib, pa_pair_result *out);
    // get_x()/get_y() after align() (pairalign -a), one pair
    void alignment(const pa_params &p, uint32_t a, uint32_t b, std::string &x, std::string &y);
    // the same for a batch: op strings (0 both, 1 base over gap, 2 gap over base) of pairs (ia[k], ib[k])
    struct OpBatch {
        std::vector<uint8_t> ops;
        std::vector<uint64_t> offsets;
        std::vector<uint32_t> n_ops;
    };
    void alignments(const pa_params &p, const std::vector<uint32_t> &ia, const std::vector<uint32_t> &ib, OpBatch &out);
    // turn one op string into the two printed lines
    void render(uint32_t a, uint32_t b, const uint8_t *ops, uint32_t n_ops, std::string &x, std::string &y) const;
    // the same into caller-owned memory (n_ops characters each): lets several host threads fill one output buffer
    void render_into(uint32_t a, uint32_t b, const uint8_t *ops, uint32_t n_ops, char *px, char *py) const;

private:
    std::vector<uint8_t> masks_;
    std::vector<uint64_t> offsets_{0};
    std::vector<std::string> unknown_;
    bool any_unknown_ = false;
};

// pa_init on the devices named by PAIRALIGN_DEVICES (comma separated) or on every
// visible device; throws std::runtime_error when there is none (no CPU fallback)
// devices from PAIRALIGN_DEVICES, else as many of the visible ones as the work justifies (est_cells < 0: all)
void init_devices(double est_cells = -1.0);

}  // namespace pab
