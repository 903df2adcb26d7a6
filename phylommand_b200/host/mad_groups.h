// mad_groups.h -- taxonomy tree of (JC - p) histograms and the approximate MAD
// test, with the behaviour of the reference's align_group class
// (src/align_group.h, src/align_group.cpp).
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace pab {

class MadGroups {
public:
    static constexpr int kPrecision = 1000;               // src/align_group.h:33
    static constexpr int kHighest = 2;                    // src/align_group.h:34
    static constexpr int kBins = kPrecision * kHighest;   // src/align_group.h:35

    MadGroups();
    // insert_value (src/align_group.cpp:23-63): the value lands in the deepest taxon
    // the two "a; b; c" strings share.  Returns false (and stores nothing) for a NaN,
    // which indexes out of bounds in the reference.
    bool insert_value(const std::string &taxon_string1, const std::string &taxon_string2, float value);
    // The same with the two strings named by small integers the caller keeps stable (one per distinct string, e.g.
    // from intern()): the histogram a pair of strings leads to is looked up once and remembered, so the 12 million
    // inserts of a 5 000-sequence run cost an array access each instead of string splitting and a tree walk.  The
    // first insert of every (id1, id2) goes through insert_value(), in the caller's order, so nodes are created in
    // the reference's order; outcomes that print a warning are not remembered (they warn every time).
    bool insert_value(int id1, const std::string &taxon_string1, int id2, const std::string &taxon_string2, float value);
    int intern(const std::string &taxon_string);
    // get_levels (src/align_group.cpp:133-166)
    std::string get_levels() const;
    // aprox_mad (src/align_group.h:62-67)
    float approx_mad() const;
    unsigned long warnings() const { return warnings_; }

private:
    struct Node {
        std::string taxon;
        std::vector<std::unique_ptr<Node>> children;   // insertion order, as the reference's linked list
        std::vector<int> hist;
        Node() : hist(kBins, 0) {}
    };
    void find_node_insert_value(std::string taxon, float value, Node *leaf);
    static void add_values(std::vector<int> &values, const Node *leaf);
    static float calc_approx_mad(const std::vector<int> &values);
    std::string get_levels(const Node *leaf) const;
    std::unique_ptr<Node> root_;
    unsigned long warnings_ = 0;
    Node *last_leaf_ = nullptr;                       // histogram the last successful insert went to
    std::vector<std::vector<Node *>> memo_;           // [id1][id2] -> histogram node, nullptr = not known
    struct NameIndex;                                 // string -> id
    std::shared_ptr<NameIndex> index_;
};

}  // namespace pab
