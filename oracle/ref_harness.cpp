// ref_harness.cpp -- thin C entry points over the UNMODIFIED reference seqpair
// class, compiled in place from /root/reference/src (see oracle/Makefile; the
// reference sources are never copied into this repository).  TEST
// INFRASTRUCTURE ONLY.  It exposes what the pairalign CLI throws away: the
// score returned by seqpair::align() (src/pairalign.cpp:683-684) and
// full-precision doubles.
#include "seqpair.h"   // -I/root/reference/src
#include <cstring>

extern "C" {

// x,y: raw sequence text exactly as pairalign would pass it (the reference
// drops x[0] and y[0], src/seqpair.cpp:80).  aligned!=0 mimics -A
// (src/pairalign.cpp:681).  ax/ay receive get_x()/get_y() (cap bytes each).
int ref_seqpair_run(const char *x, const char *y, int aligned,
                    int *score, int *hamming, double *sim, double *jc,
                    char *ax, char *ay, int cap) {
    seqpair sp((std::string(x)), (std::string(y)));
    int sc = 0;
    if (!aligned) {
        sp.set_cost_matrix(7, -5);          // src/pairalign.cpp:682
        sc = sp.align();                    // src/pairalign.cpp:683
    }
    if (score) *score = sc;
    if (hamming) *hamming = sp.hamming_distance();
    if (sim) *sim = sp.similarity();
    if (jc) *jc = sp.jc_distance();
    std::string gx = sp.get_x(), gy = sp.get_y();
    if (ax) { if ((int)gx.size() + 1 > cap) return -1; std::memcpy(ax, gx.c_str(), gx.size() + 1); }
    if (ay) { if ((int)gy.size() + 1 > cap) return -1; std::memcpy(ay, gy.c_str(), gy.size() + 1); }
    return 0;
}

}
