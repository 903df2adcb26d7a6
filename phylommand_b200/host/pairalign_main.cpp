// pairalign_main.cpp -- the pairalign command line on top of the B200 module.
//
// Same flags, same stdout (byte for byte) and the same files as the reference's
// pairalign (src/pairalign.cpp), but the all-pairs loop is turned inside out:
// the reference pulls one pair at a time through seqdatabase and aligns it on
// the spot (cluster(), src/pairalign.cpp:495-665 -> align_pair(), :675-861);
// here every sequence is read and encoded once, the pairs are aligned in large
// batches by the CUDA module (include/pairalign_b200.h) while a second thread
// replays the finished batch -- matrix framing, per-pair text, the single-link
// clustering state machine and the MAD histograms -- strictly in the
// reference's pair order, because that order is observable.
//
// There is no CPU alignment path: without a CUDA device the program stops with
// an error.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <unistd.h>
#include <algorithm>
#include <charconv>
#include <chrono>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "cluster_store.h"
#include "fasta_index.h"
#include "mad_groups.h"
#include "pairalign_b200.h"
#include "seqpair_batch.h"

namespace {

using namespace pab;

const char *kVersion = "1.1";   // src/constants.h:26
const char *kYear = "2019";     // src/constants.h:27

struct Options {
    char output_mode = 'a';
    bool matrix = false;
    bool quiet = true;
    bool output_names = false;
    bool aligned = false;
    std::string file_name;
    std::string taxonomy_file;
    std::string cut_off = "all,0.999";     // src/pairalign.cpp:126
    int min_length = 100;
    bool only_lead = false;
    std::string format = "fasta";
};

// argv_parser::pars_sub_args (src/argv_parser.cpp:53-72): split on `sep`, backslash escapes
std::vector<std::string> split_sub_args(const char *arg, char sep) {
    std::vector<std::string> out(1);
    bool escape = false;
    for (const char *p = arg; *p; ++p) {
        char c = *p;
        if (c == '\\' && !escape) { escape = true; continue; }
        if (escape) { if (c == 'n') c = '\n'; else if (c == 'r') c = '\r'; else if (c == 't') c = '\t'; }
        if (c == sep && !escape) out.emplace_back();
        else out.back() += c;
        escape = false;
    }
    return out;
}

void print_help() {
    // src/pairalign.cpp:312-403 (default build: no DATABASE, no PTHREAD, so -T / --threads is 'not recognized' here as it is there)
    std::cout << "Pairalign " << kVersion << " will perform pairwise alignment of DNA sequences given in fasta\n"
              << "format through standard in.\n"
              << "(c) Martin Ryberg " << kYear << ".\n\n"
              << "Usage:\npairalign [arguments] < inputfile.fasta\npairalign [arguments] inputfile.fasta\n\n"
              << "Arguments:\n"
              << "--aligned / -A                  input file is already aligned.\n"
              << "--alignments / -a               output aligned sequences pairwise.\n"
              << "--difference / -i               output difference between the Jukes-Cantor (JC)\n"
              << "                                distance and proportion different sites.\n"
              << "--distances / -d                output proportion different sites, JC distance,\n"
              << "                                and diference between the two.\n"
              << "--format [fasta/pairfst]        set the format of the input to fasta or fasta\n"
              << "                                with sequences pairwise (as output given the -a\n"
              << "                                -n option). If sequences are aligned give the -A\n"
              << "                                switch.\n"
              << "--group / -g                    this option will cluster sequences that are\n"
              << "                                similar and/or find the most inclusive taxa in a\n"
              << "                                hierarchy that are alignable according to MAD\n"
              << "                                (Smith et al. 2009, BMC evol. Biol. 9:37). It\n"
              << "                                need the taxonomy given after a (the first) | in\n"
              << "                                the sequence name or in a separate file. Each\n"
              << "                                taxa in the hierarchy should be separated by a\n"
              << "                                semicolon, with the highest rank first and then\n"
              << "                                increasingly nested levels until the lowest\n"
              << "                                known level for the sequence. The groups that\n"
              << "                                can be aligned are put in a file with the ending\n"
              << "                                .alignment_groups and printed to the screen\n"
              << "                                preceded by #. Clusters are printed to the\n"
              << "                                screen after a heading, preceded by ###. To get\n"
              << "                                alignable groups give 'alignment_groups' as\n"
              << "                                extra argument, to cluster give 'cluster', and\n"
              << "                                to do both give 'both'. Cut off value for\n"
              << "                                pairwise similarity can be given after colon (:)\n"
              << "                                by cut-off= followed value, e.g. -g both:\n"
              << "                                cut-off=0.97. A file with taxonomy can be given\n"
              << "                                with taxonomy=. The taxonomy file should have\n"
              << "                                the taxonomy (as above) first on each row\n"
              << "                                followed by a |, and the sequence name with that\n"
              << "                                taxonomy as a comma (,) and/or space ( )\n"
              << "                                separated string. The same taxon can be repeated\n"
              << "                                several times.\n"
              << "--help / -h                     print this help.\n"
              << "--jc_distance / -j              output Jukes-Cantor (JC) distance.\n"
              << "--matrix / -m                   output in the form of a space separated\n"
              << "                                left-upper triangular matrix.\n"
              << "--names / -n                    output sequence names (if outputting alignments\n"
              << "                                then in fasta format).\n"
              << "--proportion_difference / -p    output proportion sites that are different.\n"
              << "--similarity / -s               output similarity between sequences (1-proportion\n"
              << "                                different).\n"
              << "--verbose / -v                  get additional output.\n";
    std::cout.flush();
}

// PAIRALIGN_TIMING=1: wall-clock phases on stderr (not part of the reference's surface)
struct PhaseTimer {
    bool on = std::getenv("PAIRALIGN_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::cerr << "[pairalign_b200] " << what << ": " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " ms" << std::endl;
        t0 = t1;
    }
};

// ---- buffered stdout with the reference's number formatting -------------------
// `ostream << double` at default precision is printf("%g").
class Out {
public:
    ~Out() { flush(); }
    void put(char c) { buf_ += c; maybe_flush(); }
    void put(const std::string &s) { buf_ += s; maybe_flush(); }
    void put(const char *s) { buf_ += s; maybe_flush(); }
    // std::to_chars(general, 6) is printf("%g") -- compared on 6.6 M values incl. -0, inf, nan and denormals --
    // at 2.5 x the speed of snprintf
    static int format(double v, char *dst) {
        const auto r = std::to_chars(dst, dst + 32, v, std::chars_format::general, 6);
        return (int)(r.ptr - dst);
    }
    void num(double v) {
        char tmp[40];
        buf_.append(tmp, (size_t)format(v, tmp));
        maybe_flush();
    }
    void raw(const char *p, size_t n) {
        if (n >= (1u << 20)) {      // a batch of rendered alignments (up to gigabytes): straight to stdout, no second copy
            if (!buf_.empty()) { std::fwrite(buf_.data(), 1, buf_.size(), stdout); buf_.clear(); }
            std::fwrite(p, 1, n, stdout);
            return;
        }
        buf_.append(p, n);
        maybe_flush();
    }
    void fill(char c, size_t n) { buf_.append(n, c); maybe_flush(); }
    void flush() {
        if (!buf_.empty()) { std::fwrite(buf_.data(), 1, buf_.size(), stdout); buf_.clear(); }
        std::fflush(stdout);
    }
private:
    void maybe_flush() { if (buf_.size() > (1u << 20)) { std::fwrite(buf_.data(), 1, buf_.size(), stdout); buf_.clear(); } }
    std::string buf_;
};

// ---- the text of one pair in the statistic modes (src/pairalign.cpp:815-851) -----------
// Depends on the pair alone: its two names, its place in the triangle and its (mismatches, columns) record.
// Measured on the 16-core GPU host: 43 M values/s on one thread (49 995 000 values of a 10 000-sequence -A -j -n -m run in
// 1.17 s, the reader of the pipe being the limit); formatting batches on several threads was tried and is not faster.
class TextEmitter {
public:
    TextEmitter(char mode, bool matrix, bool names) : mode_(mode), matrix_(matrix), names_(names) {}
    static bool handles(char mode) { return mode == 'd' || mode == 'c' || mode == 'j' || mode == 'p' || mode == 's'; }

    // src/pairalign.cpp:594-602; row = number of rows begun before this one
    template <class Sink>
    void begin_row(Sink &out, const std::string &accno1, unsigned int row) const {
        if (row > 0) out.put('\n');
        if (names_) { out.put(accno1); out.put(' '); }
        if (row > 0) out.fill(' ', row);
    }
    template <class Sink>
    void pair(Sink &out, const std::string &accno1, const std::string &accno2, const PairStats &st) {
        if (mode_ == 'd') {
            if (!matrix_) {
                names(out, accno1, accno2);
                out.put("Proportion sites that are different (and similarity): ");
                out.num(st.proportion_different()); out.put(" ("); out.num(st.similarity());
                out.put("), Jukes-Cantor distance: "); out.num(st.jc_distance());
                out.put(", difference between the two: "); out.num(st.jc_minus_p()); out.put(".\n");
            } else {
                out.num(st.proportion_different()); out.put('/'); out.num(st.similarity()); out.put('/');
                out.num(st.jc_distance()); out.put('/'); out.num(st.jc_minus_p()); out.put('.');
            }
            return;
        }
        // The printed figure depends on (mismatches, columns) only, and an all-pairs run repeats the same few
        // thousand combinations millions of times: keep the formatted text in a direct-mapped table.
        const uint64_t key = (((uint64_t)st.r.dist << 32) | st.r.len) + 1;
        Formatted &f = memo_[(size_t)((key * 0x9E3779B97F4A7C15ull) >> 48)];
        if (f.key != key) {
            f.key = key;
            const double v = mode_ == 'c' ? st.jc_minus_p() : mode_ == 'j' ? st.jc_distance()
                           : mode_ == 'p' ? st.proportion_different() : st.similarity();
            f.n = (uint8_t)Out::format(v, f.txt);
        }
        if (!matrix_) { names(out, accno1, accno2); out.raw(f.txt, f.n); out.put('\n'); }
        else { out.raw(f.txt, f.n); out.put(' '); }
    }
private:
    template <class Sink>
    void names(Sink &out, const std::string &accno1, const std::string &accno2) const {
        if (names_) { out.put(accno1); out.put(" - "); out.put(accno2); out.put(" | "); }
    }
    struct Formatted { uint64_t key = 0; uint8_t n = 0; char txt[23]; };
    std::vector<Formatted> memo_ = std::vector<Formatted>(1u << 16);
    const char mode_;
    const bool matrix_, names_;
};

// ---- what one replayed pair needs ------------------------------------------------
struct PairView {
    const std::string *accno1, *accno2;
    long seq1, seq2;                 // indices into the SeqpairBatch (seq2 < 0: the reference's empty second sequence)
    long cid1, cid2;                 // indices into the cluster store (ascending accession order)
    bool new_row;                    // seqdatabase::at_new_first() (src/seqdatabase.h:135-138)
};

// Everything align_pair() does after align() (src/pairalign.cpp:690-856).
class Replayer {
public:
    Replayer(const Options &o, float cut_off, SeqpairBatch &batch, Out &out)
        : opt_(o), cut_off_(cut_off), batch_(batch), out_(out), text_(o.output_mode, o.matrix, o.output_names) {}

    std::function<std::string(const PairView &, int)> taxon_of;   // get_taxon_string of side 1 / 2 (src/seqdatabase.h:106-114)
    bool taxon_per_sequence = false;   // taxon_of depends on the sequence only: strings and their ids are kept per sequence
    std::function<float(long)> comp_of;                           // get_comp_value(accession or lead) (src/seqdatabase.h:97-105)
    // get_x()/get_y() of the pair being replayed, when a batch of alignments is at hand (mode 'a')
    std::function<bool(const PairView &, std::string &, std::string &)> alignment_of;
    ClusterStore clusters;
    MadGroups deviations;

    void begin_row(const std::string &accno1) {      // src/pairalign.cpp:594-602
        text_.begin_row(out_, accno1, n_seq_);
        ++n_seq_;
    }

    void pair(const PairView &v, const PairStats &st, const pa_params &params) {
        if (opt_.matrix && v.new_row) begin_row(*v.accno1);
        if (batch_.any_unknown()) {           // the reference re-encodes (and re-warns) per pair
            batch_.warn_unknown((size_t)v.seq1);
            if (v.seq2 >= 0) batch_.warn_unknown((size_t)v.seq2);
        }
        const char mode = opt_.output_mode;
        if (mode == 'A' || mode == 'B') group(v, st);
        else if (mode == 'a') {
            std::string x, y;
            if (v.seq2 < 0) { x = batch_.text((size_t)v.seq1); y.assign(x.size(), '-'); }
            else if (!alignment_of || !alignment_of(v, x, y)) batch_.alignment(params, (uint32_t)v.seq1, (uint32_t)v.seq2, x, y);
            if (opt_.output_names) { out_.put('>'); out_.put(*v.accno1); out_.put('\n'); }
            out_.put(x); out_.put('\n');
            if (opt_.output_names) { out_.put('>'); out_.put(*v.accno2); out_.put('\n'); }
            out_.put(y); out_.put('\n');
        } else if (TextEmitter::handles(mode)) text_.pair(out_, *v.accno1, *v.accno2, st);
        // mode 'C' (-g cluster) falls through every branch of the reference's align_pair: nothing happens
        if (!opt_.quiet) std::cerr << '.';
    }

private:
    // src/pairalign.cpp:690-805
    void group(const PairView &v, const PairStats &st) {
        const long a1 = v.cid1, a2 = v.cid2;
        if (taxon_per_sequence && v.seq2 >= 0) { group_cached(v, st); return; }
        if (v.seq2 < 0 && cut_off_ > 0.000000001 && st.similarity() > cut_off_) {
            // one sequence in the file: it absorbs the reference's empty second accession (ClusterStore::add_ghost_member)
            upd(a1, ClusterStore::LEAD, true);
            clusters.add_ghost_member(a1);
            return;
        }
        const std::string tax1 = taxon_of(v, 1), tax2 = v.seq2 >= 0 ? taxon_of(v, 2) : std::string();
        if (cut_off_ > 0.000000001 && v.seq2 >= 0) {
            const long c1 = clusters.get_cluster(a1), c2 = clusters.get_cluster(a2);
            const bool free1 = (c1 == ClusterStore::EMPTY || c1 == ClusterStore::LEAD);
            const bool free2 = (c2 == ClusterStore::EMPTY || c2 == ClusterStore::LEAD);
            if (st.similarity() > cut_off_) {
                const float comp1 = comp_of(free1 ? a1 : c1);
                const float comp2 = comp_of(free2 ? a2 : c2);
                // the better (longer, fewer N) sequence or cluster absorbs the other
                if (comp1 >= comp2) absorb(a1, c1, a2, c2, free2);
                else absorb(a2, c2, a1, c1, free1);
            } else {
                if (c1 == ClusterStore::EMPTY) upd(a1, ClusterStore::LEAD, true);
                if (c2 == ClusterStore::EMPTY) upd(a2, ClusterStore::LEAD, true);
                if (!tax1.empty() && tax1 != "empty" && !tax2.empty() && tax2 != "empty" && free1 && free2)
                    deviations.insert_value(tax1, tax2, (float)st.jc_minus_p());
            }
        } else {
            deviations.insert_value(tax1, tax2, (float)st.jc_minus_p());
        }
    }
    // group() for the all-pairs loop over one FASTA file: the taxon string of a sequence is fetched once and interned,
    // so the MAD insert of a pair is an array access (MadGroups::insert_value with ids).  Same decisions, same order.
    int cached_taxon_id(const PairView &v, int side) {
        const size_t s = (size_t)(side == 1 ? v.seq1 : v.seq2);
        if (tax_id_[s] == -2) { tax_cache_[s] = taxon_of(v, side); tax_id_[s] = deviations.intern(tax_cache_[s]); }
        return tax_id_[s];
    }
    void group_cached(const PairView &v, const PairStats &st) {
        const long a1 = v.cid1, a2 = v.cid2;
        const size_t need = (size_t)std::max(v.seq1, v.seq2) + 1;      // grow first: the references below must stay valid
        if (tax_cache_.size() < need) { tax_cache_.resize(need); tax_id_.resize(need, -2); }
        const int id1 = cached_taxon_id(v, 1), id2 = cached_taxon_id(v, 2);
        const std::string &tax1 = tax_cache_[(size_t)v.seq1], &tax2 = tax_cache_[(size_t)v.seq2];
        if (cut_off_ > 0.000000001) {
            const long c1 = clusters.get_cluster(a1), c2 = clusters.get_cluster(a2);
            const bool free1 = (c1 == ClusterStore::EMPTY || c1 == ClusterStore::LEAD);
            const bool free2 = (c2 == ClusterStore::EMPTY || c2 == ClusterStore::LEAD);
            if (st.similarity() > cut_off_) {
                const float comp1 = comp_of(free1 ? a1 : c1);
                const float comp2 = comp_of(free2 ? a2 : c2);
                if (comp1 >= comp2) absorb(a1, c1, a2, c2, free2);
                else absorb(a2, c2, a1, c1, free1);
            } else {
                if (c1 == ClusterStore::EMPTY) upd(a1, ClusterStore::LEAD, true);
                if (c2 == ClusterStore::EMPTY) upd(a2, ClusterStore::LEAD, true);
                if (!tax1.empty() && tax1 != "empty" && !tax2.empty() && tax2 != "empty" && free1 && free2)
                    deviations.insert_value(id1, tax1, id2, tax2, (float)st.jc_minus_p());
            }
        } else {
            deviations.insert_value(id1, tax1, id2, tax2, (float)st.jc_minus_p());
        }
    }
    // winner w (cluster state cw) takes in loser l (cluster state cl); src/pairalign.cpp:717-775
    void absorb(long w, long cw, long l, long cl, bool l_free) {
        long target;
        if (cw == ClusterStore::EMPTY) { upd(w, ClusterStore::LEAD, true); target = w; }
        else if (cw == ClusterStore::LEAD) target = w;
        else { if (cw == cl) return; target = cw; }        // already in the same cluster
        upd(l, target, true);
        if (l_free) upd(l, target, false);
        else { upd(cl, target, true); upd(cl, target, false); }
    }
    void upd(long accno, long cluster, bool where_accno) {
        if (!clusters.update(accno, cluster, where_accno))
            std::cerr << "Warning!!! Accession number cluster name collision when updating clusters." << std::endl;
    }

    const Options &opt_;
    const float cut_off_;
    SeqpairBatch &batch_;
    Out &out_;
    TextEmitter text_;
    unsigned int n_seq_ = 0;
    std::vector<std::string> tax_cache_;   // taxon string per sequence (taxon_per_sequence)
    std::vector<int> tax_id_;              // its MadGroups::intern() id, -2: not fetched yet
};

// cut-off string "gene,value,gene,value" or a bare number (src/pairalign.cpp:456-479)
float parse_cut_off(const std::string &cut_off, const std::string &table) {
    float present = 0.0f;
    const int length = (int)cut_off.length();
    std::string gene;
    int i = 0;
    for (; i < length; ++i) {
        if (cut_off[i] == ',') {
            if (gene == table || gene == "all") {
                std::string number;
                ++i;
                while (i < length && cut_off[i] != ',') { number += cut_off[i]; ++i; }
                present = (float)atof(number.c_str());
                gene.clear();
                break;
            }
        } else gene += cut_off[i];
    }
    if (i > 0 && i >= length && !gene.empty() && (std::isdigit((unsigned char)gene[0]) || gene[0] == '.'))
        present = (float)atof(gene.c_str());
    return present;
}

// taxonomy file: "tax1; tax2|acc1,acc2 acc3" per row (src/pairalign.cpp:409-439)
bool read_taxonomy(const std::string &path, std::map<std::string, std::string> &tax) {
    std::ifstream f(path.c_str());
    if (!f.good()) { std::cerr << "Was not able to open " << path << ". No taxonomy read." << std::endl; return false; }
    std::string taxa, accno;
    bool taxonomy = true;
    while (f) {
        const char c = (char)f.get();
        if (c == '|' && taxonomy) taxonomy = false;
        else if (!taxonomy && (c == ' ' || c == ',' || c == '\n' || c == '\r') && !accno.empty() && !taxa.empty()) {
            tax[accno] = taxa;
            accno.clear();
            if (c == '\r' || c == '\n') { taxonomy = true; taxa.clear(); }
        } else if (c == '\r' || c == '\n') { taxonomy = true; taxa.clear(); }
        else if (taxonomy) taxa += c;
        else accno += c;
    }
    if (tax.empty()) { std::cerr << "Was not able to pars taxonomy from " << path << "." << std::endl; return false; }
    return true;
}

std::string lookup_taxonomy(const std::map<std::string, std::string> &tax, const std::string &accno) {
    auto it = tax.find(accno);            // get_taxon_string_from_map (src/seqdatabase.h:304-311)
    if (it != tax.end()) return it->second;
    it = tax.find("default");
    if (it != tax.end()) return it->second;
    return std::string();
}

// Double-buffered batches: the CUDA module fills one buffer while the caller replays the other.  chunk_at(first) is the
// size of the batch that starts at `first`; produce and consume are told which of the two buffers (slot) a batch uses.
template <class ChunkAt, class Produce, class Consume>
void pipeline(uint64_t total, ChunkAt chunk_at, Produce produce, Consume consume) {
    if (total == 0) return;
    std::vector<pa_pair_result> buf[2];
    std::string error;
    auto fill = [&](int slot, uint64_t first, uint64_t n) {
        buf[slot].resize((size_t)n);
        try { produce(first, n, buf[slot].data(), slot); } catch (const std::exception &e) { error = e.what(); }
    };
    auto size_at = [&](uint64_t first) { return std::min<uint64_t>(std::max<uint64_t>(1, chunk_at(first)), total - first); };
    int slot = 0;
    uint64_t n = size_at(0);
    fill(0, 0, n);
    for (uint64_t first = 0; first < total;) {
        if (!error.empty()) throw std::runtime_error(error);
        const uint64_t next = first + n;
        const uint64_t n_next = next < total ? size_at(next) : 0;
        std::thread worker;
        if (n_next) worker = std::thread(fill, slot ^ 1, next, n_next);
        // a throwing consumer must not leave the worker joinable (its destructor would call std::terminate)
        try { consume(first, buf[slot], slot); } catch (...) { if (worker.joinable()) worker.join(); throw; }
        if (worker.joinable()) worker.join();
        slot ^= 1;
        first = next;
        n = n_next;
    }
}

constexpr uint64_t kChunkPairs = 1ull << 21;

// ---- fasta input: all pairs of the file -------------------------------------------
void run_fasta(const Options &opt, Out &out) {
    PhaseTimer timer;
    FastaIndex index;
    index.open(opt.file_name);
    timer.mark("read + index FASTA");
    if (!opt.quiet) std::cerr << "Opened " << opt.format << " database." << std::endl;
    std::map<std::string, std::string> taxonomy;
    if (!opt.taxonomy_file.empty()) {
        if (!opt.quiet) std::cerr << "Parsing taxonomy from " << opt.taxonomy_file << "." << std::endl;
        if (read_taxonomy(opt.taxonomy_file, taxonomy) && !opt.quiet) std::cerr << "Added taxonomy." << std::endl;
    }
    const char mode = opt.output_mode;
    std::ofstream groups_file;
    if (mode == 'A' || mode == 'B') {
        if (!opt.quiet) std::cerr << "No alignment_groups file/table present. Trying to create it." << std::endl;
        const std::string name = opt.file_name.empty() ? "sequence.alignment_groups" : opt.file_name + ".alignment_groups";
        groups_file.open(name.c_str());
        if (!groups_file.is_open()) { std::cerr << "Failed to create alignment_groups." << std::endl; return; }
    }
    const std::string table = opt.file_name;     // the single pseudo-table (src/seqdatabase.h:254-258)
    float cut_off = 0.0f;
    if (mode == 'B' || mode == 'C') {
        cut_off = parse_cut_off(opt.cut_off, table);
        if (cut_off < 0.000000001) {
            std::cerr << "Could not find appropriate cut off (" << cut_off << ") for " << table
                      << ". Will only define alignment groups and not cluster." << std::endl;
            return;
        }
        if (!opt.quiet) std::cerr << "Using the cut off: " << cut_off << "." << std::endl;
    }
    if (!opt.quiet) std::cerr << "Checking " << table << std::endl;
    if (!opt.quiet) std::cerr << "Starting pairwise alignment." << std::endl;

    const size_t N = index.size();
    SeqpairBatch batch;
    Replayer rp(opt, cut_off, batch, out);
    if (N == 0) {
        std::cerr << "Could not initiate sequence retrieval. No aligning done for " << table << "." << std::endl;
    } else {
        for (size_t s = 0; s < N; ++s) batch.add_sequence(index[s].text);
        timer.mark("encode sequences");
        rp.clusters.reset(N);
        rp.taxon_of = [&](const PairView &v, int side) {
            const FastaRecord &r = index[(size_t)(side == 1 ? v.seq1 : v.seq2)];
            return taxonomy.empty() ? r.taxon : lookup_taxonomy(taxonomy, r.accno);
        };
        rp.comp_of = [&](long s) { return index[(size_t)s].comp_value(); };
        rp.taxon_per_sequence = true;
        if (opt.matrix && mode == 'd')
            out.put("Proportion different/Similarity/Jukes-Cantor distance/Difference between JC and similarity\n");
        pa_params params{7, -5, -15, -1, opt.aligned ? 1 : 0};      // src/pairalign.cpp:682, src/seqpair.h:57-58
        if (N == 1) {
            // the reference still visits one pair whose second sequence is empty (src/seqdatabase.cpp:108);
            // nothing can be compared: similarity 1, JC -0
            static const std::string none;
            PairView v{&index[0].accno, &none, 0, -1, 0, -1, true};
            PairStats st{{0, 0, 0, (int)batch.length(0) - 1, -1}};
            rp.pair(v, st, params);
        } else {
            const bool need_stats = (mode != 'a' && mode != 'C');
            const uint64_t total = (uint64_t)N * (N - 1) / 2;
            if (need_stats || mode == 'a') {
                double sum = 0, sum2 = 0;                       // DP cells of all pairs: ((sum len)^2 - sum len^2) / 2
                for (size_t s = 0; s < N; ++s) { const double l = batch.length(s); sum += l; sum2 += l * l; }
                init_devices(opt.aligned ? 0.0 : (sum * sum - sum2) / 2);
                timer.mark("pa_init (CUDA contexts)");
                batch.upload();
                timer.mark("pa_upload_sequences");
            }
            // mode 'a': batches of alignments (op strings) travel beside the record buffers
            SeqpairBatch::OpBatch opb[2];
            std::vector<uint32_t> ia_b, ib_b;
            uint32_t max_len = 1;
            for (size_t s = 0; s < N; ++s) max_len = std::max(max_len, batch.length(s));
            const bool want_ops = (mode == 'a' && !opt.aligned);
            // op strings of one batch: 256 MB per device in use (every device takes a contiguous share of the batch)
            const uint64_t op_bytes = (256ull << 20) * (uint64_t)std::max(1, pa_device_count());
            const uint64_t chunk_pairs = want_ops ? std::max<uint64_t>(1, std::min<uint64_t>(kChunkPairs, op_bytes / (2ull * max_len)))
                                                  : kChunkPairs;
            auto chunk_at = [&](uint64_t) { return chunk_pairs; };
            size_t cur_k = 0;
            int cur_slot = 0;
            if (want_ops)
                rp.alignment_of = [&](const PairView &v, std::string &x, std::string &y) {
                    const SeqpairBatch::OpBatch &ob = opb[cur_slot];
                    batch.render((uint32_t)v.seq1, (uint32_t)v.seq2, ob.ops.data() + ob.offsets[cur_k], ob.n_ops[cur_k], x, y);
                    return true;
                };
            uint32_t a = 0, b = 0;       // pair cursor in reference order: (0,1),(0,2)..(1,2)..
            // -a without per-pair side effects (warnings about unknown characters, progress dots): the text of a whole
            // batch is rendered straight into one buffer by several host threads -- with the GPU at 2.6 TCUPS on 1.5 kb
            // pairs one thread's 0.3 GB/s of formatting was five times slower than the alignments it prints
            const bool bulk_text = want_ops && opt.quiet && !opt.matrix && !batch.any_unknown();   // '-m -a': rows are framed per pair
            std::string bulk;
            auto consume_bulk = [&](uint64_t first, size_t n, int slot) {
                const SeqpairBatch::OpBatch &ob = opb[slot];
                std::vector<uint32_t> pa_(n), pb_(n);
                std::vector<uint64_t> at(n + 1, 0);
                uint32_t ca = 0, cb = 0;
                pa_pair_from_index(first, &ca, &cb);
                for (size_t k = 0; k < n; ++k) {
                    pa_[k] = ca; pb_[k] = cb;
                    uint64_t bytes = 2ull * ((uint64_t)ob.n_ops[k] + 1);
                    if (opt.output_names) bytes += 4 + index[ca].accno.size() + index[cb].accno.size();
                    at[k + 1] = at[k] + bytes;
                    if (++cb == N) { ++ca; cb = ca + 1; }
                }
                bulk.resize((size_t)at[n]);
                char *base = bulk.empty() ? nullptr : &bulk[0];
                auto fill = [&](size_t k0, size_t k1) {
                    for (size_t k = k0; k < k1; ++k) {
                        char *p = base + at[k];
                        const uint32_t no = ob.n_ops[k];
                        if (opt.output_names) { *p++ = '>'; const std::string &s1 = index[pa_[k]].accno; std::memcpy(p, s1.data(), s1.size()); p += s1.size(); *p++ = '\n'; }
                        char *px = p; p += no; *p++ = '\n';
                        if (opt.output_names) { *p++ = '>'; const std::string &s2 = index[pb_[k]].accno; std::memcpy(p, s2.data(), s2.size()); p += s2.size(); *p++ = '\n'; }
                        char *py = p; p += no; *p++ = '\n';
                        if (no) batch.render_into(pa_[k], pb_[k], ob.ops.data() + ob.offsets[k], no, px, py);
                    }
                };
                const size_t n_thr = std::max<size_t>(1, std::min<size_t>({(size_t)std::thread::hardware_concurrency(), (size_t)16, (size_t)(at[n] >> 20) + 1}));
                if (n_thr == 1) fill(0, n);
                else {      // contiguous shares with about the same number of bytes
                    std::vector<std::thread> th;
                    size_t k0 = 0;
                    for (size_t t = 0; t < n_thr; ++t) {
                        size_t k1 = k0;
                        const uint64_t target = at[n] * (t + 1) / n_thr;
                        while (k1 < n && at[k1 + 1] <= target) ++k1;
                        if (t + 1 == n_thr) k1 = n;
                        if (k1 > k0) th.emplace_back(fill, k0, k1);
                        k0 = k1;
                    }
                    for (auto &t : th) t.join();
                }
                out.raw(bulk.data(), bulk.size());
            };
            auto consume = [&](uint64_t first, const std::vector<pa_pair_result> &recs, int slot) {
                cur_slot = slot;
                if (bulk_text) { consume_bulk(first, recs.size(), slot); return; }
                for (size_t k = 0; k < recs.size(); ++k) {
                    cur_k = k;
                    if (b == 0) { a = 0; b = 1; }
                    PairView v{&index[a].accno, &index[b].accno, (long)a, (long)b, (long)a, (long)b, b == a + 1};
                    PairStats st{recs[k]};
                    rp.pair(v, st, params);
                    if (++b == N) { ++a; b = a + 1; }
                }
            };
            auto produce = [&](uint64_t first, uint64_t n, pa_pair_result *dst, int slot) {
                if (need_stats) batch.align_range(params, first, n, dst);
                else std::memset(dst, 0, (size_t)n * sizeof(pa_pair_result));
                if (want_ops) {
                    ia_b.resize((size_t)n); ib_b.resize((size_t)n);
                    uint32_t pa = 0, pb = 0;
                    pa_pair_from_index(first, &pa, &pb);
                    for (uint64_t k = 0; k < n; ++k) {
                        ia_b[(size_t)k] = pa; ib_b[(size_t)k] = pb;
                        if (++pb == N) { ++pa; pb = pa + 1; }
                    }
                    batch.alignments(params, ia_b, ib_b, opb[slot]);
                }
            };
            double replay_ms = 0, produce_ms = 0;       // busy time of the two sides of the pipeline (PAIRALIGN_TIMING)
            auto timed = [&](double &acc, auto &&fn) {
                const auto t0 = std::chrono::steady_clock::now();
                fn();
                acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            };
            pipeline(total, chunk_at,
                     [&](uint64_t first, uint64_t n, pa_pair_result *dst, int slot) { timed(produce_ms, [&] { produce(first, n, dst, slot); }); },
                     [&](uint64_t first, const std::vector<pa_pair_result> &recs, int slot) { timed(replay_ms, [&] { consume(first, recs, slot); }); });
            if (timer.on)
                std::cerr << "[pairalign_b200] pipeline busy: align " << produce_ms << " ms, replay + print " << replay_ms << " ms" << std::endl;
            if (timer.on) {
                pa_timing tm;
                if (pa_get_timing(&tm) == PA_OK)
                    std::cerr << "[pairalign_b200] last batch: kernels " << tm.kernel_ms << " ms, d2h " << tm.d2h_ms << " ms, call "
                              << tm.total_ms << " ms on " << tm.n_devices << " device(s)" << std::endl;
            }
            timer.mark("align + replay (pipelined)");
            // src/pairalign.cpp:629-630 -- printed with or without -n
            if (opt.matrix) { out.put('\n'); out.put(index[N - 1].accno); out.put('\n'); }
        }
    }
    if (!opt.quiet) std::cerr << std::endl;
    if (mode == 'A' || mode == 'B') {
        if (!opt.quiet)
            std::cerr << "Finished aligning. Calculating mad to determine taxonomic level suitable for alignment." << std::endl;
        const std::string levels = rp.deviations.get_levels();
        out.put("# Alignment groups for "); out.put(table); out.put('\n');
        out.put("#    Alignment groups: "); out.put(levels); out.put('\n');
        if (!opt.quiet) std::cerr << "    Aprox. mad. entire group = " << rp.deviations.approx_mad() << std::endl;
        if (groups_file.good()) groups_file << table << "\t" << levels << std::endl;
    }
    if (mode == 'B' || mode == 'C') {
        out.flush();
        std::cout << "### Clusters " << table << ", cut-off: " << cut_off << " ###" << std::endl;
        rp.clusters.print(std::cout, [&](long s) -> const std::string & { return index[(size_t)s].accno; });
        std::cout.flush();
    }
}

// ---- pair-fasta input: consecutive record pairs (src/seqdatabase.cpp:34-67) -------
struct PairRecord {
    std::string accno1, accno2, seq1, seq2, tax1, tax2;
    bool new_row;
};

void run_pairfasta(const Options &opt, Out &out) {
    std::ifstream file;
    std::istream *input = &std::cin;
    if (!opt.file_name.empty()) { file.open(opt.file_name.c_str()); input = &file; }
    if (!opt.quiet) std::cerr << "Opened pairfa database." << std::endl;
    const char mode = opt.output_mode;
    std::map<std::string, std::string> taxon_strings;
    if (!opt.taxonomy_file.empty()) {                          // src/pairalign.cpp:409-439
        if (!opt.quiet) std::cerr << "Parsing taxonomy from " << opt.taxonomy_file << "." << std::endl;
        if (read_taxonomy(opt.taxonomy_file, taxon_strings) && !opt.quiet) std::cerr << "Added taxonomy." << std::endl;
    } else {
        taxon_strings["default"] = "all";                      // src/pairalign.cpp:440-445
        if (!opt.quiet && (mode == 'A' || mode == 'B')) std::cerr << "All sequences will be treated as from same taxon." << std::endl;
    }
    std::ofstream groups_file;
    if (mode == 'A' || mode == 'B') {
        if (!opt.quiet) std::cerr << "No alignment_groups file/table present. Trying to create it." << std::endl;
        const std::string name = opt.file_name.empty() ? "sequence.alignment_groups" : opt.file_name + ".alignment_groups";
        groups_file.open(name.c_str());
        if (!groups_file.is_open()) { std::cerr << "Failed to create alignment_groups." << std::endl; return; }
    }
    const std::string table = opt.file_name;
    float cut_off = 0.0f;
    if (mode == 'B' || mode == 'C') {
        cut_off = parse_cut_off(opt.cut_off, table);
        if (cut_off < 0.000000001) {
            std::cerr << "Could not find appropriate cut off (" << cut_off << ") for " << table
                      << ". Will only define alignment groups and not cluster." << std::endl;
            return;
        }
        if (!opt.quiet) std::cerr << "Using the cut off: " << cut_off << "." << std::endl;
    }
    if (!opt.quiet) std::cerr << "Checking " << table << std::endl;
    // read every pair with the reference's character state machine; taxon strings found in the
    // headers are APPENDED to the map each time an accession is seen (src/seqdatabase.cpp:55-56)
    std::vector<PairRecord> pairs;
    std::string previous = "empty", last_accno2;
    bool ok = opt.file_name.empty() || file.good();
    while (ok) {
        PairRecord r;
        char read_mode = '0';
        while (*input) {
            const char c = (char)input->get();
            if (c == '>') { if (read_mode == '0') read_mode = 'A'; else if (read_mode == 'S') read_mode = 'a'; }
            else if (c == '|') { if (read_mode == 'A') read_mode = 'T'; else if (read_mode == 'a') read_mode = 't'; }
            else if (c == '\n' || c == '\r') {
                if (read_mode == 'A' || read_mode == 'T') read_mode = 'S';
                else if (read_mode == 'a' || read_mode == 't') read_mode = 's';
            }
            else if (read_mode == 'A') r.accno1 += c;
            else if (read_mode == 'a') r.accno2 += c;
            else if (read_mode == 'T') taxon_strings[r.accno1] += c;
            else if (read_mode == 't') taxon_strings[r.accno2] += c;
            else if (read_mode == 'S') r.seq1 += c;
            else if (read_mode == 's') r.seq2 += c;
            if (read_mode == 's' && (input->peek() == '>' || input->peek() == EOF)) break;
        }
        r.new_row = (previous == "empty" || previous != r.accno1);
        const bool last = input->bad() || input->peek() == EOF;
        if (last) r.new_row = true;                              // mode '9' also counts as a new first
        r.tax1 = lookup_taxonomy(taxon_strings, r.accno1);
        r.tax2 = lookup_taxonomy(taxon_strings, r.accno2);
        if (!r.accno1.empty()) previous = r.accno1;
        last_accno2 = r.accno2;
        pairs.push_back(std::move(r));
        if (last) break;
    }
    if (!opt.quiet) std::cerr << "Starting pairwise alignment." << std::endl;
    if (!ok) std::cerr << "Could not initiate sequence retrieval. No aligning done for " << table << "." << std::endl;   // src/pairalign.cpp:632
    SeqpairBatch batch;
    Replayer rp(opt, cut_off, batch, out);
    // accession dictionary in ascending order = the reference's map order for printed clusters
    std::map<std::string, long> dict;
    for (const auto &r : pairs) { dict[r.accno1] = 0; dict[r.accno2] = 0; }
    std::vector<const std::string *> name_of;
    for (auto &kv : dict) { kv.second = (long)name_of.size(); name_of.push_back(&kv.first); }
    rp.clusters.reset(name_of.size());
    std::vector<uint32_t> ia, ib;
    for (const auto &r : pairs) {
        ia.push_back((uint32_t)batch.add_sequence(r.seq1));
        ib.push_back((uint32_t)batch.add_sequence(r.seq2));
    }
    if (ok && opt.matrix && mode == 'd')
        out.put("Proportion different/Similarity/Jukes-Cantor distance/Difference between JC and similarity\n");
    pa_params params{7, -5, -15, -1, opt.aligned ? 1 : 0};
    std::vector<pa_pair_result> recs(pairs.size());
    if (!pairs.empty() && mode != 'C') {
        double cells = 0;
        for (size_t k = 0; k < ia.size(); ++k) cells += (double)batch.length(ia[k]) * (double)batch.length(ib[k]);
        init_devices(opt.aligned ? 0.0 : cells);
        batch.upload();
        if (mode != 'a') batch.align_list(params, ia, ib, recs.data());
    }
    size_t cur = 0;
    rp.taxon_of = [&](const PairView &, int side) { return side == 1 ? pairs[cur].tax1 : pairs[cur].tax2; };
    // get_comp_value_pair (src/seqdatabase.cpp:22-32): non-N characters of the first sequence of the
    // current pair whose name DIFFERS from the one asked for (the reference's test is inverted)
    auto non_n = [](const std::string &s) { float v = 0; for (char c : s) if (c != 'n' && c != 'N') v += 1.0f; return v; };
    rp.comp_of = [&](long which) {
        const PairRecord &r = pairs[cur];
        const std::string &accno = *name_of[(size_t)which];
        if (accno != r.accno1) return non_n(r.seq1);
        if (accno != r.accno2) return non_n(r.seq2);
        return 0.0f;
    };
    for (cur = 0; cur < pairs.size(); ++cur) {
        const PairRecord &r = pairs[cur];
        PairView v{&r.accno1, &r.accno2, (long)ia[cur], (long)ib[cur], dict[r.accno1], dict[r.accno2], r.new_row};
        rp.pair(v, PairStats{recs[cur]}, params);
    }
    if (opt.matrix && !pairs.empty() && !last_accno2.empty()) { out.put('\n'); out.put(last_accno2); out.put('\n'); }
    if (!opt.quiet) std::cerr << std::endl;
    if (mode == 'A' || mode == 'B') {
        if (!opt.quiet)
            std::cerr << "Finished aligning. Calculating mad to determine taxonomic level suitable for alignment." << std::endl;
        const std::string levels = rp.deviations.get_levels();
        out.put("# Alignment groups for "); out.put(table); out.put('\n');
        out.put("#    Alignment groups: "); out.put(levels); out.put('\n');
        if (!opt.quiet) std::cerr << "    Aprox. mad. entire group = " << rp.deviations.approx_mad() << std::endl;
        if (groups_file.good()) groups_file << table << "\t" << levels << std::endl;
    }
    if (mode == 'B' || mode == 'C') {
        out.flush();
        std::cout << "### Clusters " << table << ", cut-off: " << cut_off << " ###" << std::endl;
        rp.clusters.print(std::cout, [&](long s) -> const std::string & { return *name_of[(size_t)s]; });
        std::cout.flush();
    }
}

}  // namespace

int main(int argc, char *argv[]) {
    Options opt;
    // flag handling and return codes as src/pairalign.cpp:133-268 (most errors return 0)
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        auto is = [&](const char *s, const char *l) { return !strcmp(a, s) || !strcmp(a, l); };
        if (is("-a", "--alignments")) opt.output_mode = 'a';
        else if (is("-d", "--distances")) opt.output_mode = 'd';
        else if (is("-p", "--proportion_difference")) opt.output_mode = 'p';
        else if (is("-s", "--similarity")) opt.output_mode = 's';
        else if (is("-j", "--jc_distance")) opt.output_mode = 'j';
        else if (is("-i", "--difference")) opt.output_mode = 'c';
        else if (is("-m", "--matrix")) { opt.matrix = true; if (opt.output_mode == 'a') opt.output_mode = 'p'; }
        else if (is("-n", "--names")) opt.output_names = true;
        else if (is("-A", "--aligned")) opt.aligned = true;
        else if (is("-v", "--verbose")) opt.quiet = false;
        else if (is("-f", "--file")) {
            if (i + 1 < argc && argv[i + 1][0] != '-') opt.file_name = argv[++i];
            else { std::cerr << "-f/--file needs to be followed by a file name." << std::endl; return 1; }
        }
        else if (!strcmp(a, "--format")) {
            if (i + 1 < argc && argv[i + 1][0] != '-') {
                ++i;
                if (!strcmp(argv[i], "fasta")) opt.format = "fasta";
                else if (!strcmp(argv[i], "pairfst") || !strcmp(argv[i], "pairfa")) opt.format = "pairfa";
                else { std::cerr << argv[i] << " is not a valid option for --format. The options are pairfst, or fasta." << std::endl; return 1; }
            } else { std::cerr << "--format require fasta or pairwise as extra argument. Use -h for more help." << std::endl; return 1; }
        }
        else if (is("-g", "--group")) {
            opt.output_mode = 'A';
            if (i + 1 < argc && argv[i + 1][0] != '-') {
                ++i;
                const std::vector<std::string> args = split_sub_args(argv[i], ':');
                if (args[0] == "alignment_groups") opt.output_mode = 'A';
                else if (args[0] == "cluster") opt.output_mode = 'C';
                else if (args[0] == "both") opt.output_mode = 'B';
                else {
                    std::cerr << "Do not recognize argument '" << args[0] << "'. Options are 'alignment_groups', 'cluster', 'both'." << std::endl;
                    return 1;
                }
                for (size_t k = 1; k < args.size(); ++k) {
                    std::vector<std::string> sub = split_sub_args(args[k].c_str(), '=');
                    const std::string &key = sub[0];
                    const bool has_value = sub.size() > 1;
                    auto value = [&]() { return has_value ? sub[1] : std::string(); };
                    if ((key == "cut-off" || key == "cut_off" || key == "cutoff" || key == "cut off") && key.size() > 1) {
                        if (!has_value) { std::cerr << "Unrecognized argument or missing value for argument given to -g/--group: " << key << "." << std::endl; return 1; }
                        opt.cut_off = value();
                    } else if ((key == "min-length" || key == "min_length" || key == "minlength" || key == "min length") && key.size() > 1) {
                        if (!has_value) { std::cerr << "Unrecognized argument or missing value for argument given to -g/--group: " << key << "." << std::endl; return 1; }
                        opt.min_length = atoi(value().c_str());
                    } else if ((key == "taxon" || key == "taxonomy" || key == "taxon_file" || key == "taxonfile") && key.size() > 1) {
                        if (!has_value) { std::cerr << "Unrecognized argument or missing value for argument given to -g/--group: " << key << "." << std::endl; return 1; }
                        opt.taxonomy_file = value();
                    } else if (key == "only_lead" || key == "only-lead" || key == "only lead" || key == "previous_clusters") {
                        opt.only_lead = true;
                    } else {
                        std::cerr << "Unrecognized argument or missing value for argument given to -g/--group: " << key << "." << std::endl;
                        return 1;
                    }
                }
            }
        }
        else if (is("-h", "--help")) { print_help(); return 0; }
        else if (i == argc - 1 && a[0] != '-' && opt.file_name.empty()) opt.file_name = a;
        else {
            std::cerr << "The program was called with the following command:" << std::endl;
            for (int j = 0; j < argc; ++j) std::cerr << argv[j] << ' ';
            std::cerr << std::endl;
            std::cerr << "Argument " << a << " not recognized. For available arguments give -h or --help." << std::endl;
            return 0;
        }
    }
    if (!opt.quiet) {
        std::cerr << "The program was called with the following command:" << std::endl;
        for (int i = 0; i < argc; ++i) std::cerr << argv[i] << ' ';
        std::cerr << std::endl << std::endl;
    }
    bool error_flag = false;
    if (opt.min_length < 0) {
        std::cerr << "Minimum sequence length to consider for clustering must be positive integer, e.g. 100." << std::endl;
        error_flag = true;
    }
    if (error_flag) { std::cerr << "Quitting quietly." << std::endl; return 0; }
    time_t rawtime = time(0);
    if (!opt.quiet) std::cerr << "Started at:" << std::endl << asctime(localtime(&rawtime));
    int rc = 0;
    try {
        Out out;
        if (opt.format == "pairfa") run_pairfasta(opt, out);
        else run_fasta(opt, out);
        out.flush();
    } catch (const std::exception &e) {
        std::fflush(stdout);
        std::cerr << "pairalign_b200: " << e.what() << std::endl;
        rc = 2;
    }
    rawtime = time(0);
    if (!opt.quiet) std::cerr << "Ended at:" << std::endl << asctime(localtime(&rawtime));
    // Everything is written; tearing the CUDA contexts down in order costs seconds on a multi-GPU box
    // and buys nothing, so leave without it (PAIRALIGN_CLEAN_EXIT=1 keeps the orderly path for sanitizers).
    std::cout.flush();
    std::cerr.flush();
    std::fflush(nullptr);
    if (std::getenv("PAIRALIGN_CLEAN_EXIT")) { pa_shutdown(); return rc; }
    _exit(rc);
}
