#include "fasta_index.h"

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>

namespace pab {

bool FastaIndex::open(const std::string &name) {
    records_.clear();
    if (name.empty()) {
        parse(std::cin);
        opened_ = true;
    } else {
        std::ifstream f(name.c_str());
        opened_ = f.good();
        if (opened_) parse(f);
    }
    return good();
}

void FastaIndex::parse(std::istream &in) {
    std::map<std::string, FastaRecord> index;
    std::string line, accno, preamble;
    bool seen_header = false;
    unsigned int seqno = 0;
    while (std::getline(in, line)) {
        if (!line.empty() && line[0] == '>') {
            seen_header = true;
            ++seqno;
            accno.clear();
            std::string taxon;
            char mode = 'a';
            for (size_t i = 1; i < line.size(); ++i) {
                const char c = line[i];
                if (c == '|' && (mode == 't' || mode == 'T')) break;
                else if (c == '|' && mode == 'a') mode = 't';
                else if (mode == 'a' && c != ' ') accno += c;
                else if (mode == 't' && c != ' ') { taxon += c; mode = 'T'; }
                else if (mode == 'T') taxon += c;
            }
            if (accno.empty()) accno = std::to_string(seqno);
            FastaRecord r;
            r.accno = accno;
            r.taxon = taxon;
            index[accno] = r;            // a repeated accession starts over
        } else {
            if (!seen_header) preamble += line;
            bool counted = false;
            unsigned int n = 0, non_n = 0;
            for (char c : line) {
                if (c != ' ' && c != '\n' && c != '\r' && c != '\t') {
                    counted = true;
                    if (c == 'N' || c == 'n') ++n; else ++non_n;
                }
            }
            auto it = index.find(accno);
            if (it == index.end()) {
                if (!counted) continue;           // the reference only creates the entry when it counts a character
                FastaRecord r;
                r.accno = accno;                  // "" : text before the first header
                it = index.insert(std::make_pair(accno, r)).first;
            }
            it->second.n_count += n;
            it->second.non_n_count += non_n;
            if (seen_header) it->second.text += line;
        }
    }
    auto pre = index.find("");
    if (pre != index.end()) pre->second.text = preamble;   // seekg(0): everything up to the first header line
    records_.reserve(index.size());
    for (auto &kv : index) records_.push_back(std::move(kv.second));
}

long FastaIndex::find(const std::string &accno) const {
    auto it = std::lower_bound(records_.begin(), records_.end(), accno,
                               [](const FastaRecord &r, const std::string &k) { return r.accno < k; });
    if (it == records_.end() || it->accno != accno) return -1;
    return (long)(it - records_.begin());
}

}  // namespace pab
