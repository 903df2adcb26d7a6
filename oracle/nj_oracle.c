/* nj_oracle.c -- CPU restatement of phylommand's neighbour joining (treeator -n).
 *
 * TEST INFRASTRUCTURE ONLY (see pa_oracle.h).  Follows njtree::build_nj_tree
 * (src/nj_tree.cpp:32-205) literally in its arithmetic -- float sums in the
 * reference's order, float Q values, first strict minimum -- but keeps the
 * distances in a square matrix with an order list instead of ragged vectors,
 * and lets a few threads share the rows of one round (every row's sum and
 * first minimum is computed exactly as one thread would; the rows are then
 * scanned in order, so the winner is the one a single scan finds).
 * Parity status: PINNED against the unmodified reference treeator built from
 * /root/reference/src (tests/golden/nj/, made by oracle/make_nj_golden.py;
 * checked by tests/test_nj_oracle.py).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef struct {
    uint32_t left, right;      /* node ids: 0..n-1 are the taxa in matrix order, n.. the joins in creation order */
    double left_len, right_len;
} nj_oracle_join;

typedef struct {
    const float *D;
    const uint32_t *slot;
    float *S, *rowmin;
    int *rowarg;
    uint32_t n;
    int r, phase, n_threads, quit;
    pthread_barrier_t bar;
} nj_shared;

typedef struct { nj_shared *sh; int tid; } nj_worker;

static void nj_rows(nj_shared *sh, int tid) {
    const float *D = sh->D;
    const uint32_t *slot = sh->slot;
    const uint32_t n = sh->n;
    const int r = sh->r;
    for (int p = tid; p < r; p += sh->n_threads) {
        if (sh->phase == 0) {
            /* step 1: S of taxon p; its partners are met in ascending order (:39-47) */
            float s = 0;
            for (int q = 0; q < r; ++q) if (q != p) s += D[(size_t)slot[p] * n + slot[q]];
            sh->S[p] = s;
        } else {
            /* step 2, one row: first strictly smallest (r-2)*d - S_p - S_q, starting from 100000 (:53-74) */
            float m = 100000;
            int bq = -1;
            for (int q = p + 1; q < r; ++q) {
                float value = (r - 2) * D[(size_t)slot[p] * n + slot[q]] - sh->S[p] - sh->S[q];
                if (value < m) { m = value; bq = q; }
            }
            sh->rowmin[p] = m; sh->rowarg[p] = bq;
        }
    }
}

static void *nj_thread(void *arg) {
    nj_worker *w = (nj_worker *)arg;
    for (;;) {
        pthread_barrier_wait(&w->sh->bar);
        if (w->sh->quit) return 0;
        nj_rows(w->sh, w->tid);
        pthread_barrier_wait(&w->sh->bar);
    }
}

static void nj_run_phase(nj_shared *sh, int phase) {
    sh->phase = phase;
    if (sh->n_threads == 1) { nj_rows(sh, 0); return; }
    pthread_barrier_wait(&sh->bar);
    nj_rows(sh, 0);
    pthread_barrier_wait(&sh->bar);
}

/* dist: upper triangle, row-major, n(n-1)/2 floats.  joins: n-2 records.
 * The tree is ((... joins ...)root_left:0, root_right:root_right_len).  Returns 0, -1 on bad arguments. */
int nj_oracle_build(const float *dist, uint32_t n, nj_oracle_join *joins,
                    uint32_t *root_left, uint32_t *root_right, double *root_right_len) {
    if (n < 2 || !dist) return -1;
    float *D = (float *)calloc((size_t)n * n, sizeof(float));
    float *S = (float *)malloc(n * sizeof(float));
    float *rowmin = (float *)malloc(n * sizeof(float));
    int *rowarg = (int *)malloc(n * sizeof(int));
    uint32_t *slot = (uint32_t *)malloc(n * sizeof(uint32_t));   /* order position -> matrix slot */
    uint32_t *node = (uint32_t *)malloc(n * sizeof(uint32_t));   /* order position -> node id     */
    if (!D || !S || !slot || !node || !rowmin || !rowarg) {
        free(D); free(S); free(slot); free(node); free(rowmin); free(rowarg);
        return -1;
    }
    size_t k = 0;
    for (uint32_t a = 0; a < n; ++a)
        for (uint32_t b = a + 1; b < n; ++b) { D[(size_t)a * n + b] = dist[k]; D[(size_t)b * n + a] = dist[k]; ++k; }
    for (uint32_t a = 0; a < n; ++a) { slot[a] = a; node[a] = a; }
    nj_shared sh;
    memset(&sh, 0, sizeof sh);
    sh.D = D; sh.slot = slot; sh.S = S; sh.rowmin = rowmin; sh.rowarg = rowarg; sh.n = n;
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    sh.n_threads = n < 600 ? 1 : (int)(cores < 1 ? 1 : cores > 16 ? 16 : cores);
    pthread_t tids[16];
    nj_worker workers[16];
    if (sh.n_threads > 1) {
        pthread_barrier_init(&sh.bar, 0, (unsigned)sh.n_threads);
        for (int t = 1; t < sh.n_threads; ++t) {
            workers[t].sh = &sh; workers[t].tid = t;
            pthread_create(&tids[t], 0, nj_thread, &workers[t]);
        }
    }
    int r = (int)n;
    uint32_t next_id = n, nj = 0;
    while (r > 2) {
        sh.r = r;
        nj_run_phase(&sh, 0);
        nj_run_phase(&sh, 1);
        float M = 100000;
        int bi = 0, bj = 0;
        for (int p = 0; p < r; ++p)
            if (rowarg[p] >= 0 && rowmin[p] < M) { M = rowmin[p]; bi = p; bj = rowarg[p] - p - 1; }
        const int i = bi, jp = bi + bj + 1;
        /* steps 3-4 (:79-103) */
        const float length = D[(size_t)slot[i] * n + slot[jp]];
        const double left_len = (length / 2) + (S[i] - S[jp]) / (2 * (r - 2));
        const double right_len = length - left_len;
        joins[nj].left = node[i]; joins[nj].right = node[jp];
        joins[nj].left_len = left_len; joins[nj].right_len = right_len;
        ++nj;
        /* step 5 (:108-176): the new node goes to the FRONT; its distances follow the old order of the others */
        const uint32_t ns = slot[i];                 /* reuse taxon i's slot for the new node (after reading its row) */
        float *newd = (float *)malloc((size_t)r * sizeof(float));
        int cnt = 0;
        for (int p = 0; p < r; ++p) {
            if (p == i || p == jp) continue;
            newd[cnt++] = (D[(size_t)slot[p] * n + slot[i]] + D[(size_t)slot[p] * n + slot[jp]] - length) / 2;
        }
        uint32_t *slot2 = (uint32_t *)malloc((size_t)r * sizeof(uint32_t)), *node2 = (uint32_t *)malloc((size_t)r * sizeof(uint32_t));
        slot2[0] = ns; node2[0] = next_id++;
        cnt = 1;
        for (int p = 0; p < r; ++p) if (p != i && p != jp) { slot2[cnt] = slot[p]; node2[cnt] = node[p]; ++cnt; }
        for (int c = 1; c < cnt; ++c) { D[(size_t)ns * n + slot2[c]] = newd[c - 1]; D[(size_t)slot2[c] * n + ns] = newd[c - 1]; }
        D[(size_t)ns * n + ns] = 0;
        memcpy(slot, slot2, (size_t)cnt * sizeof(uint32_t));
        memcpy(node, node2, (size_t)cnt * sizeof(uint32_t));
        free(newd); free(slot2); free(node2);
        r = cnt;
    }
    if (sh.n_threads > 1) {
        sh.quit = 1;
        pthread_barrier_wait(&sh.bar);
        for (int t = 1; t < sh.n_threads; ++t) pthread_join(tids[t], 0);
        pthread_barrier_destroy(&sh.bar);
    }
    /* :193-201 */
    *root_left = node[0];
    *root_right = node[1];
    *root_right_len = D[(size_t)slot[0] * n + slot[1]];
    free(D); free(S); free(slot); free(node); free(rowmin); free(rowarg);
    return 0;
}
