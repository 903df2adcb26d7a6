// b200_batch.h -- glue between the reference's own pairalign.cpp (as patched by
// integration/pairalign_b200.patch, compiled with -DPAIRALIGN_B200) and the
// C-ABI of include/pairalign_b200.h.  C++11, header only, no CUDA types.
//
// The reference aligns one pair at a time: cluster() pulls a pair from the
// seqdatabase iterator and align_pair() builds a `seqpair`, calls align() and
// asks it for similarity()/jc_distance()/get_x()/get_y()
// (src/pairalign.cpp:527-628, :675-856).  The B200 module is batched, so the
// patched cluster() runs in two sweeps over the same loop:
//   1. collect: every pair the iterator visits is noted (b200::batch::add), each
//      distinct sequence encoded once with pa_encode_sequence;
//   2. align: one upload, one pa_align_all_pairs (when the visited pairs are the
//      row-major triangle, which is what the fasta iterator produces) or
//      pa_align_pairs (anything else), op strings in chunks for -a;
//   3. replay: the original loop body runs per noted pair with a b200::pair in
//      place of the seqpair -- same member names, so align_pair() is untouched
//      below its first five lines.
// Nothing here computes an alignment: without a CUDA device pa_init fails and the
// program exits with status 2 (there is no CPU fallback).
#ifndef B200_BATCH_H
#define B200_BATCH_H

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "pairalign_b200.h"

namespace b200 {

inline void check(int rc) {
    if (rc != PA_OK) {
        std::cerr << "pairalign_b200: " << pa_last_error() << std::endl;
        std::exit(2);
    }
}

struct sequence {
    std::string accno, text;
    std::vector<uint8_t> masks;        // translate_to_binary (src/seqpair.cpp:74-92): first character dropped
    std::string unknown;               // characters outside the alphabet, in order (the reference warns per pair)
};

struct noted_pair { uint32_t a, b; bool new_row; };

class batch;

// What align_pair() asks of its `sequences` object (src/seqpair.h:83-99).
class pair {
public:
    pair() : owner(0), k(0) {}
    pair(batch *o, size_t index) : owner(o), k(index) {}
    void set_cost_matrix(int, int) {}          // 7 / -5 went to the device with pa_params
    int align();                               // returns the score of the record
    double similarity();
    double jc_distance();
    std::string get_x();
    std::string get_y();
private:
    batch *owner;
    size_t k;
};

class batch {
public:
    batch() : aligned(false), want_ops(false), ops_first(0), ops_count(0) {}
    size_t size() const { return pairs.size(); }
    void add(const std::string &accno1, const std::string &seq1, const std::string &accno2, const std::string &seq2, bool new_row) {
        noted_pair p;
        p.a = intern(accno1, seq1);
        p.b = intern(accno2, seq2);
        p.new_row = new_row;
        pairs.push_back(p);
    }
    // one upload, one call: replaces seqpair::align() for every pair of the loop
    void run(bool aligned_input, bool alignments) {
        aligned = aligned_input;
        want_ops = alignments && !aligned_input;
        if (pairs.empty()) return;
        std::vector<uint8_t> masks;
        std::vector<uint64_t> offsets(1, 0);
        for (size_t s = 0; s < seqs.size(); ++s) {
            masks.insert(masks.end(), seqs[s].masks.begin(), seqs[s].masks.end());
            offsets.push_back(masks.size());
        }
        check(pa_init(0, 0));
        check(pa_upload_sequences(masks.empty() ? 0 : &masks[0], &offsets[0], (uint32_t)seqs.size()));
        pa_params prm = {7, -5, -15, -1, aligned ? 1 : 0};      // src/pairalign.cpp:682, src/seqpair.h:57-58
        params = prm;
        rec.resize(pairs.size());
        bool triangle = pairs.size() == pa_num_pairs();
        uint32_t a = 0, b = 1;
        for (size_t k = 0; triangle && k < pairs.size(); ++k) {
            triangle = pairs[k].a == a && pairs[k].b == b;
            if (++b == seqs.size()) { ++a; b = a + 1; }
        }
        if (triangle) check(pa_align_all_pairs(&params, 0, rec.size(), &rec[0]));
        else {
            std::vector<uint32_t> ia(pairs.size()), ib(pairs.size());
            for (size_t k = 0; k < pairs.size(); ++k) { ia[k] = pairs[k].a; ib[k] = pairs[k].b; }
            check(pa_align_pairs(&params, &ia[0], &ib[0], ia.size(), &rec[0]));
        }
    }
    const noted_pair &at(size_t k) const { return pairs[k]; }
    const sequence &seq(uint32_t s) const { return seqs[s]; }
    // the reference constructs a seqpair per pair and warns about unknown characters each time (src/seqpair.cpp:86)
    pair view(size_t k) {
        const noted_pair &p = pairs[k];
        for (int side = 0; side < 2; ++side) {
            const std::string &u = seqs[side ? p.b : p.a].unknown;
            for (size_t c = 0; c < u.size(); ++c) std::cerr << "Can not interpret '" << u[c] << "'. Not in alphabet." << std::endl;
        }
        return pair(this, k);
    }

private:
    friend class pair;
    uint32_t intern(const std::string &accno, const std::string &text) {
        std::map<std::string, uint32_t>::iterator it = index.find(accno);
        if (it != index.end() && seqs[it->second].text == text) return it->second;
        sequence s;
        s.accno = accno; s.text = text;
        s.masks.resize(text.size() + 1);
        size_t unk = 0;
        std::vector<char> unknown(text.size() + 1);
        const size_t n = pa_encode_sequence(text.data(), text.size(), &s.masks[0], &unk, &unknown[0], unknown.size());
        s.masks.resize(n);
        s.unknown.assign(unknown.begin(), unknown.begin() + (unk < unknown.size() ? unk : unknown.size()));
        seqs.push_back(s);
        index[accno] = (uint32_t)(seqs.size() - 1);
        return (uint32_t)(seqs.size() - 1);
    }
    static std::string decode(const uint8_t *m, size_t n) {
        std::string out(n, '-');
        for (size_t k = 0; k < n; ++k) out[k] = pa_mask_to_char(m[k]);
        return out;
    }
    // op strings for pairs [k, k + chunk): 0 = x over y, 1 = x over a gap, 2 = a gap over y
    void ensure_ops(size_t k) {
        if (k >= ops_first && k < ops_first + ops_count) return;
        size_t n = 0, bytes = 0;
        while (k + n < pairs.size() && n < 65536 && bytes < (256u << 20)) {
            bytes += seqs[pairs[k + n].a].masks.size() + seqs[pairs[k + n].b].masks.size();
            ++n;
        }
        std::vector<uint32_t> ia(n), ib(n);
        for (size_t q = 0; q < n; ++q) { ia[q] = pairs[k + q].a; ib[q] = pairs[k + q].b; }
        ops.resize(bytes ? bytes : 1);
        op_off.resize(n + 1);
        n_ops.resize(n);
        check(pa_align_pairs_ops(&params, &ia[0], &ib[0], n, &ops[0], ops.size(), &op_off[0], &n_ops[0], 0));
        ops_first = k; ops_count = n;
    }
    void render(size_t k, std::string &x, std::string &y) {
        const sequence &sx = seqs[pairs[k].a], &sy = seqs[pairs[k].b];
        if (!want_ops || sx.masks.empty() || sy.masks.empty()) {   // -A prints its input; an empty side has nothing to align
            x = decode(sx.masks.empty() ? 0 : &sx.masks[0], sx.masks.size());
            y = decode(sy.masks.empty() ? 0 : &sy.masks[0], sy.masks.size());
            return;
        }
        ensure_ops(k);
        const uint8_t *o = &ops[op_off[k - ops_first]];
        const uint32_t n = n_ops[k - ops_first];
        x.assign(n, '-'); y.assign(n, '-');
        size_t i = 0, j = 0;
        for (uint32_t c = 0; c < n; ++c) {
            if (o[c] != 2) x[c] = pa_mask_to_char(sx.masks[i++]);
            if (o[c] != 1) y[c] = pa_mask_to_char(sy.masks[j++]);
        }
    }

    std::vector<sequence> seqs;
    std::map<std::string, uint32_t> index;
    std::vector<noted_pair> pairs;
    std::vector<pa_pair_result> rec;
    pa_params params;
    bool aligned, want_ops;
    std::vector<uint8_t> ops;
    std::vector<uint64_t> op_off;
    std::vector<uint32_t> n_ops;
    size_t ops_first, ops_count;
};

inline int pair::align() { return owner->rec[k].score; }
inline double pair::similarity() { return pa_similarity(owner->rec[k].dist, owner->rec[k].len); }
inline double pair::jc_distance() { return pa_jc_distance(owner->rec[k].dist, owner->rec[k].len); }
inline std::string pair::get_x() { std::string x, y; owner->render(k, x, y); return x; }
inline std::string pair::get_y() { std::string x, y; owner->render(k, x, y); return y; }

}  // namespace b200
#endif
