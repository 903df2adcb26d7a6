// fasta_index.h -- load-once replacement for the reference's indexedfasta
// (src/indexedfasta.h, src/indexedfasta.cpp) and the pair iterator built on it
// (seqdatabase::move_to_next_pair_fst, src/seqdatabase.cpp:69-115).
//
// The reference keeps a std::map<accession, {taxon_string, N, nonN, file position}>
// and re-reads both sequences from disk for every pair.  Here the file is read
// once; the observable semantics are kept:
//   * records are ordered by accession (std::less<std::string>), not file order;
//   * a repeated accession replaces the earlier record (src/indexedfasta.cpp:61);
//   * accession = header text before the first '|' with blanks removed; taxon
//     string = text after it (leading blanks skipped) up to a second '|'
//     (src/indexedfasta.cpp:44-54); an empty accession becomes the 1-based
//     ordinal of the record (src/indexedfasta.cpp:55-59);
//   * non-blank characters on sequence lines are counted as N / non-N for the
//     clustering "comp value" (src/indexedfasta.cpp:66-73, src/indexedfasta.h:34);
//   * text before the first header belongs to an accession "" (a side effect
//     of `++index[accno].N` with an empty accno, src/indexedfasta.cpp:70).
// One deliberate deviation: input WITHOUT a final newline.  The reference's
// `cin >> noskipws` loop then duplicates the last character (stdin: the last
// sequence gains a base), and from a file its stream is left failed after the
// first read of the last record, so every later read returns an empty sequence.
// Here such input reads like the same text with the newline
// (tests/test_host_fuzz_vs_reference.py::test_input_without_final_newline...).
#pragma once
#include <istream>
#include <map>
#include <string>
#include <vector>

namespace pab {

struct FastaRecord {
    std::string accno;
    std::string taxon;
    unsigned int n_count = 0;      // 'N' / 'n'
    unsigned int non_n_count = 0;
    std::string text;              // sequence lines concatenated (as get_sequence returns them)
    // float(nonN - N) with the reference's unsigned subtraction (src/indexedfasta.h:34)
    float comp_value() const { return float(non_n_count - n_count); }
};

class FastaIndex {
public:
    // name empty: read standard input (src/indexedfasta.cpp:23-33)
    bool open(const std::string &name);
    bool good() const { return opened_ && !records_.empty(); }
    size_t size() const { return records_.size(); }
    const FastaRecord &operator[](size_t k) const { return records_[k]; }
    FastaRecord &at(size_t k) { return records_[k]; }
    // index of an accession or -1
    long find(const std::string &accno) const;
    // taxonomy file support (seqdatabase::add_taxonomy, src/seqdatabase.h:158-160)
private:
    void parse(std::istream &in);
    bool opened_ = false;
    std::vector<FastaRecord> records_;   // ascending accession
};

}  // namespace pab
