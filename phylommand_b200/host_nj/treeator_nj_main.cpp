// treeator_nj_main.cpp -- `treeator -n` on the B200: reads the distance matrix pairalign -m prints,
// joins neighbours on the GPU (pa_nj_build, csrc/pa_nj.cu) and prints the tree as the reference does.
//
// Drop-in for the neighbour-joining entry of treeator only (reference src/treeator.cpp:378-393):
//   -n/--neighbour_joining, -L/--no_label, -0/--no_branch_length, -f/--file <matrix> or a trailing file name,
//   -d/--data_file, --output newick|new|w|no, -v/--verbose, -h/--help; the matrix comes from stdin otherwise.
// Everything else treeator does (parsimony, likelihood, simulation, nexus output) is out of scope and is
// refused with a message.  There is no CPU fallback: without a GPU the program fails.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

#include "pairalign_b200.h"

namespace {

struct Matrix {
    std::vector<std::string> names;
    std::vector<std::vector<float>> rows;
};

// njtree::read_distance_matrix (src/nj_tree.cpp:252-352): values end at blank, tab, CR, LF -- or at the LAST
// character of the input, which is never part of a value (infile.peek() == EOF); the first value of a line is the
// label (or, with -L, a distance and the label is the row number); a label alone on its line keeps the row open for
// the next line; a last taxon named by its number is added when the final row still has distances.
Matrix read_distance_matrix(const std::string &data, bool labels) {
    Matrix m;
    bool new_row = true;
    int n_taxa = 0;
    std::string value;
    for (size_t k = 0; k < data.size(); ++k) {
        const char ch = data[k];
        const bool at_end = k + 1 == data.size();
        if (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t' || at_end) {
            if (!value.empty()) {
                if (new_row) {
                    // the reference reuses the last row when it has no node yet; that never happens after the first push
                    m.rows.emplace_back();
                    if (labels) m.names.push_back(value);
                    else {
                        m.names.push_back(std::to_string(n_taxa));
                        m.rows.back().push_back((float)atof(value.c_str()));
                    }
                    ++n_taxa;
                    if (ch != '\n' && ch != '\r') new_row = false;
                } else m.rows.back().push_back((float)atof(value.c_str()));
                value.clear();
            }
        } else value += ch;
        if (ch == '\n' || ch == '\r') new_row = true;
    }
    if (!m.rows.empty() && !m.rows.back().empty()) {
        m.rows.emplace_back();
        m.names.push_back(std::to_string(n_taxa));
    }
    return m;
}

// njtree::matrix_good (src/nj_tree.cpp:22-30)
bool matrix_good(const Matrix &m) {
    size_t n = m.rows.size();
    for (const auto &row : m.rows) {
        if (row.size() != n - 1) return false;
        --n;
    }
    return n == 0;
}

struct Printer {
    const Matrix &m;
    const std::vector<pa_nj_join> &joins;
    bool br;
    std::string out;
    void length(double v) {
        if (!br) return;
        char buf[400];
        const int k = std::snprintf(buf, sizeof buf, ":%f", v);     // ':' << fixed << branchlength (src/tree.cpp:278)
        out.append(buf, (size_t)k);
    }
    // tree::print_newick_subtree (src/tree.cpp:239-279) without recursion: a caterpillar of 60 000 taxa is fine
    void subtree(uint32_t id, double len) {
        struct Frame { uint32_t id; double len; int phase; };
        const uint32_t n = (uint32_t)m.names.size();
        std::vector<Frame> st{{id, len, 0}};
        while (!st.empty()) {
            Frame f = st.back();
            st.pop_back();
            if (f.id < n) { out += m.names[f.id]; length(f.len); continue; }
            const pa_nj_join &j = joins[f.id - n];
            if (f.phase == 0) {
                out += '(';
                st.push_back({f.id, f.len, 1});
                st.push_back({j.left, j.left_len, 0});
            } else if (f.phase == 1) {
                out += ',';
                st.push_back({f.id, f.len, 2});
                st.push_back({j.right, j.right_len, 0});
            } else { out += ')'; length(f.len); }
        }
    }
};

void help() {
    std::cout << "treeator_b200 -n: neighbour joining of a pairalign -m distance matrix on the GPU.\n"
                 "Usage: treeator_b200 -n [-L] [-0] [--output newick|no] [-v] [matrix_file | < matrix]\n";
}

}  // namespace

int main(int argc, char **argv) {
    bool labels = true, quiet = true, print_br_length = true;
    char method = 'p', print_tree = 'w';
    std::string data_file_name, tree_file_name;
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (!strcmp(a, "-L") || !strcmp(a, "--no_label")) labels = false;
        else if (!strcmp(a, "-d") || !strcmp(a, "--data_file")) {
            if (i < argc - 1 && argv[i + 1][0] != '-') {
                ++i;
                if (data_file_name.empty()) data_file_name = argv[i];
                else std::cerr << "Data file already given (" << data_file_name << "). Will ignore " << argv[i] << "." << std::endl;
            } else { std::cerr << "-d/--data_file require a file name as next argument" << std::endl; return 1; }
        }
        else if (!strcmp(a, "-n") || !strcmp(a, "--neighbour_joining")) method = 'n';
        else if (!strcmp(a, "-v") || !strcmp(a, "--verbose")) quiet = false;
        else if (!strcmp(a, "-h") || !strcmp(a, "--help")) { help(); return 0; }
        else if (!strcmp(a, "-0") || !strcmp(a, "--no_branch_length")) print_br_length = false;
        else if (!strcmp(a, "--output")) {
            if (i < argc - 1 && argv[i + 1][0] != '-') {
                ++i;
                if (!strcmp(argv[i], "newick") || !strcmp(argv[i], "new") || !strcmp(argv[i], "w")) print_tree = 'w';
                else if (!strcmp(argv[i], "no")) print_tree = 'N';
                else if (!strcmp(argv[i], "nexus") || !strcmp(argv[i], "nex") || !strcmp(argv[i], "x")) {
                    std::cerr << "treeator_b200: nexus output is outside the neighbour-joining drop-in (use the reference treeator)." << std::endl;
                    return 1;
                } else { std::cerr << "Do not recognize format " << argv[i] << "." << std::endl; return 1; }
            } else std::cerr << "--output require nexus(nex or x) or newick (new or w) as additional argument" << std::endl;
        }
        else if ((i == argc - 1 && a[0] != '-') ||
                 ((!strcmp(a, "-f") || !strcmp(a, "--file")) && i < argc - 1 && argv[i + 1][0] != '-' && ++i)) {
            if (data_file_name.empty()) data_file_name = argv[i];
            else if (tree_file_name.empty()) tree_file_name = argv[i];
            else { std::cerr << "Have no use for argument " << argv[i] << ". Already have file names for tree and data." << std::endl; return 1; }
        }
        else { std::cerr << "Unrecognized argument " << a << ". Quitting quietly." << std::endl; return 1; }
    }
    if (!quiet) {
        std::cerr << "The program was called with the following command:" << std::endl;
        for (int i = 0; i < argc; ++i) std::cerr << argv[i] << ' ';
        std::cerr << std::endl << std::endl;
    }
    if (method != 'n') {
        std::cerr << "treeator_b200 only does neighbour joining (-n); parsimony and likelihood are outside this drop-in." << std::endl;
        return 1;
    }
    std::string data;
    if (!data_file_name.empty()) {
        if (!quiet) std::cerr << "Reading data from: " << data_file_name << std::endl;
        std::ifstream in(data_file_name.c_str(), std::ios::in | std::ios::binary);
        if (!in.good()) { std::cerr << "Could not open file: " << data_file_name << std::endl; return 1; }
        data.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
    } else data.assign(std::istreambuf_iterator<char>(std::cin), std::istreambuf_iterator<char>());
    if (!quiet) std::cerr << "Reading distance matrix." << std::endl;
    const Matrix m = read_distance_matrix(data, labels);
    if (m.rows.empty() || !matrix_good(m)) {
        std::cerr << "Error in distance matrix. Check distance matrix format." << std::endl;
        return 1;
    }
    const uint32_t n = (uint32_t)m.rows.size();
    if (!quiet) std::cerr << "Read distances for " << n << " taxa." << std::endl;
    if (n < 2 || n > PA_NJ_MAX_TAXA) {
        std::cerr << "treeator_b200: neighbour joining needs 2 to " << PA_NJ_MAX_TAXA << " taxa, got " << n << "." << std::endl;
        return 1;
    }
    std::vector<float> tri;
    tri.reserve((size_t)n * (n - 1) / 2);
    for (const auto &row : m.rows) tri.insert(tri.end(), row.begin(), row.end());
    if (!quiet) std::cerr << "Creating NJ tree." << std::endl;
    std::vector<pa_nj_join> joins(n > 2 ? n - 2 : 0);
    uint32_t root_left = 0, root_right = 1;
    double root_right_len = 0.0, ms = 0.0;
    const int rc = pa_nj_build(tri.data(), n, joins.data(), &root_left, &root_right, &root_right_len, &ms);
    if (rc != PA_OK) {
        std::cerr << "treeator_b200: " << pa_last_error() << std::endl;
        return 2;
    }
    if (getenv("PAIRALIGN_TIMING")) std::cerr << "timing: neighbour joining " << ms << " ms on the device" << std::endl;
    if (print_tree == 'w') {
        Printer p{m, joins, print_br_length, std::string()};
        p.out += '(';
        p.subtree(root_left, 0.0);
        p.out += ',';
        p.subtree(root_right, root_right_len);
        p.out += ");\n";
        std::fwrite(p.out.data(), 1, p.out.size(), stdout);
    }
    return 0;
}
