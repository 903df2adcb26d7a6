// cluster_store.h -- single-link cluster bookkeeping with the observable
// behaviour of seqdatabase::single_link_clusters (src/seqdatabase.h:161-221),
// keyed by sequence index instead of accession string.
//
// The reference keeps std::map<lead accession, std::set<member accession>> and
// answers get_cluster() with a linear scan over all clusters (two scans per
// pair).  Sequences here are numbered in ascending accession order, so index
// order is the reference's map/set order and the printed clusters are
// identical; a reverse index makes get_cluster() O(1).
#pragma once
#include <ostream>
#include <set>
#include <string>
#include <vector>

namespace pab {

class ClusterStore {
public:
    enum : long { EMPTY = -1, LEAD = -2 };   // the reference's "empty" / "lead" strings

    explicit ClusterStore(size_t n = 0) { reset(n); }
    void reset(size_t n) {
        has_key_.assign(n, 0);
        members_.assign(n, std::set<long>());
        member_of_.assign(n, std::vector<long>());
        ghost_.assign(n, 0);
    }
    // A file with ONE sequence: the reference still visits a pair whose second iterator is end() -- empty accession,
    // empty sequence, similarity 1 -- and with a cut-off below 1 the only sequence becomes a lead with the empty name
    // as its member: its cluster line ends in a blank ("only \n"; -O0 and -O2 builds agree).
    void add_ghost_member(long lead) { ghost_[lead] = 1; }
    // get_cluster (src/seqdatabase.h:203-211): LEAD if it heads a cluster, else the
    // first (lowest) lead whose member set holds it, else EMPTY
    long get_cluster(long accno) const {
        if (has_key_[accno]) return LEAD;
        const std::vector<long> &in = member_of_[accno];
        if (in.empty()) return EMPTY;
        long best = in[0];
        for (long l : in) if (l < best) best = l;
        return best;
    }
    // update (src/seqdatabase.h:163-202).  cluster may be LEAD.  Returns false for the
    // "accession / cluster name collision" case, which the reference only warns about.
    bool update(long accno, long cluster, bool where_accno) {
        if (accno == cluster) return false;
        if (where_accno) {
            if (cluster == LEAD) {              // clusters[accno] = set<string>()
                drop_members(accno);
                has_key_[accno] = 1;
            } else if (cluster >= 0 && has_key_[cluster]) {
                add_member(cluster, accno);
            }
        } else if (has_key_[accno] && cluster >= 0) {
            if (!has_key_[cluster]) has_key_[cluster] = 1;      // clusters[cluster] = clusters[accno]
            std::set<long> moved;
            moved.swap(members_[accno]);
            for (long m : moved) { remove_ref(m, accno); add_member(cluster, m); }
            has_key_[accno] = 0;
        }
        return true;
    }
    // print_clusters (src/seqdatabase.h:212-218)
    template <class NameOf>
    void print(std::ostream &out, NameOf name_of) const {
        for (size_t l = 0; l < has_key_.size(); ++l) {
            if (!has_key_[l]) continue;
            out << name_of((long)l);
            if (ghost_[l]) out << ' ';                      // the empty accession sorts first in the reference's set
            for (long m : members_[l]) out << ' ' << name_of(m);
            out << '\n';
        }
    }
    size_t n_clusters() const { size_t c = 0; for (char k : has_key_) c += k ? 1 : 0; return c; }

private:
    void add_member(long lead, long m) {
        if (members_[lead].insert(m).second) member_of_[m].push_back(lead);
    }
    void remove_ref(long m, long lead) {
        std::vector<long> &v = member_of_[m];
        for (size_t k = 0; k < v.size(); ++k) if (v[k] == lead) { v[k] = v.back(); v.pop_back(); break; }
    }
    void drop_members(long lead) {
        for (long m : members_[lead]) remove_ref(m, lead);
        members_[lead].clear();
    }
    std::vector<char> has_key_;
    std::vector<std::set<long>> members_;
    std::vector<std::vector<long>> member_of_;
    std::vector<char> ghost_;
};

}  // namespace pab
