#include "mad_groups.h"

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <unordered_map>

namespace pab {

struct MadGroups::NameIndex { std::unordered_map<std::string, int> ids; };

MadGroups::MadGroups() : root_(new Node()), index_(new NameIndex()) {}

int MadGroups::intern(const std::string &s) {
    auto it = index_->ids.find(s);
    if (it != index_->ids.end()) return it->second;
    const int id = (int)index_->ids.size();
    index_->ids.emplace(s, id);
    return id;
}

bool MadGroups::insert_value(int id1, const std::string &s1, int id2, const std::string &s2, float value) {
    if (id1 < 0 || id2 < 0 || std::isnan(value)) return insert_value(s1, s2, value);
    const size_t need = (size_t)(id1 > id2 ? id1 : id2) + 1;
    if (memo_.size() < need) memo_.resize(need);
    std::vector<Node *> &row = memo_[(size_t)id1];
    if (row.size() <= (size_t)id2) row.resize((size_t)id2 + 1, nullptr);
    if (Node *leaf = row[(size_t)id2]) {
        // find_node_insert_value's last step (src/align_group.cpp:84-91)
        value *= kPrecision;
        if (value > kBins - 1) value = kBins - 1;
        int bin = int(value);
        if (bin < 0) bin = 0;
        leaf->hist[bin] += 1;
        return true;
    }
    const unsigned long w0 = warnings_;
    last_leaf_ = nullptr;
    const bool ok = insert_value(s1, s2, value);
    if (ok && warnings_ == w0 && last_leaf_) row[(size_t)id2] = last_leaf_;
    return ok;
}

bool MadGroups::insert_value(const std::string &s1, const std::string &s2, float value) {
    // common leading taxa of "a; b; c" strings; the separator is ';' plus ONE skipped
    // character (src/align_group.cpp:41,45: i += 2)
    std::string inclusive;
    size_t i = 0, j = 0;
    const size_t l1 = s1.size(), l2 = s2.size();
    while (i < l1 && j < l2) {
        std::string t1, t2;
        while (i < l1 && s1[i] != ';') t1 += s1[i++];
        i += 2;
        while (j < l2 && s2[j] != ';') t2 += s2[j++];
        j += 2;
        if (t1 == t2) { inclusive += t1; inclusive += ';'; }
        else break;
    }
    if (inclusive.empty()) inclusive = root_->taxon;
    if (std::isnan(value)) {   // int(NaN) indexes outside the histogram in the reference (undefined behaviour)
        ++warnings_;
        std::cerr << "WARNING!!! JC distance undefined (more than 75% different sites); value not inserted." << std::endl;
        return false;
    }
    find_node_insert_value(inclusive, value, root_.get());
    return true;
}

void MadGroups::find_node_insert_value(std::string taxon, float value, Node *leaf) {
    // src/align_group.cpp:68-118
    std::string highest, next, rest;
    size_t i = 0;
    while (i < taxon.size() && taxon[i] != ';') highest += taxon[i++];
    ++i;
    while (i < taxon.size() && taxon[i] != ';') next += taxon[i++];
    while (i < taxon.size()) rest += taxon[i++];
    if (leaf->taxon.empty()) leaf->taxon = highest;
    if (highest != leaf->taxon) {
        ++warnings_;
        std::cerr << "WARNING!!! Error in align_group::insert_value!!! Reached unexpected node!!! Value not inserted!!!" << std::endl;
        return;
    }
    if (next.empty()) {
        value *= kPrecision;
        if (value > kBins - 1) value = kBins - 1;
        int bin = int(value);
        if (bin < 0) bin = 0;      // only reachable through -0.0 / rounding; the reference would index below the array
        leaf->hist[bin] += 1;
        last_leaf_ = leaf;
        return;
    }
    bool found = false;
    const size_t n_children = leaf->children.size();
    for (size_t c = 0; c < n_children; ++c) {
        if (!found && next == leaf->children[c]->taxon) {
            found = true;
            find_node_insert_value(next + rest, value, leaf->children[c].get());
        }
    }
    if (!found) {
        leaf->children.emplace_back(new Node());
        leaf->children.back()->taxon = next;
        find_node_insert_value(next + rest, value, leaf->children.back().get());
    }
}

void MadGroups::add_values(std::vector<int> &values, const Node *leaf) {
    for (const auto &c : leaf->children) add_values(values, c.get());
    for (int k = 0; k < kBins; ++k) values[k] += leaf->hist[k];
}

float MadGroups::calc_approx_mad(const std::vector<int> &values) {
    // src/align_group.cpp:196-217: median bin, then median of |bin - median|
    int sum = 0;
    for (int k = 0; k < kBins; ++k) sum += values[k];
    std::vector<int> deviation(kBins, 0);
    int median = 0;
    int target = (sum / 2) + sum % 2;
    for (int k = 0; k < kBins; ++k) {
        median = k;
        target -= values[k];
        if (target <= 0) break;
    }
    for (int k = 0; k < kBins; ++k) deviation[std::abs(k - median)] += values[k];
    target = (sum / 2) + sum % 2;
    for (int k = 0; k < kBins; ++k) {
        median = k;
        target -= deviation[k];
        if (target <= 0) break;
    }
    return 1.4826 * float(median) / kPrecision;
}

std::string MadGroups::get_levels(const Node *leaf) const {
    std::vector<int> values(kBins, 0);
    add_values(values, leaf);
    const float mad = calc_approx_mad(values);
    if (mad < 0.01) {
        if (leaf->taxon.empty()) return "empty";
        return leaf->taxon + "_A;";
    }
    if (!leaf->children.empty()) {
        std::string out;
        for (const auto &c : leaf->children) out += get_levels(c.get());
        return out;
    }
    return leaf->taxon + "_T;";
}

std::string MadGroups::get_levels() const { return get_levels(root_.get()); }

float MadGroups::approx_mad() const {
    std::vector<int> values(kBins, 0);
    add_values(values, root_.get());
    return calc_approx_mad(values);
}

}  // namespace pab
