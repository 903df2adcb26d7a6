// nj_build_oracle.cpp -- TEST DOUBLE, never part of the product.
//
// Answers pa_nj_build with the oracle's neighbour joining (oracle/nj_oracle.c) so that the host side of
// build/treeator_b200 -- the reference's matrix reader, option handling and newick printer as restated in
// phylommand_b200/host_nj/treeator_nj_main.cpp -- can be compared with the reference's output on a machine
// without a GPU (tests/test_host_replay_cpu.py).  The executable's definition takes precedence over the library's.
#include <cstdint>

#include "../../include/pairalign_b200.h"

extern "C" {
typedef struct { uint32_t left, right; double left_len, right_len; } nj_oracle_join;
int nj_oracle_build(const float *dist, uint32_t n, nj_oracle_join *joins, uint32_t *root_left, uint32_t *root_right,
                    double *root_right_len);

int pa_nj_build(const float *dist, uint32_t n, pa_nj_join *joins, uint32_t *root_left, uint32_t *root_right,
                double *root_right_len, double *kernel_ms) {
    static_assert(sizeof(nj_oracle_join) == sizeof(pa_nj_join), "same record");
    if (kernel_ms) *kernel_ms = 0.0;
    return nj_oracle_build(dist, n, reinterpret_cast<nj_oracle_join *>(joins), root_left, root_right, root_right_len) ? PA_EINVAL : PA_OK;
}
}
