/*
 * CPU oracle for the local-greedy MWIS half of the hot path.  TEST INFRASTRUCTURE ONLY:
 * built into oracle/liblgs_oracle.so and loaded solely by tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline / reference arm.  Nothing under distgcn_b200/ links or calls it.
 *
 * Plain-C restatement of the reference's synchronous-round distributed greedy heuristic
 * (reference heuristics.py, paths relative to the reference root):
 *   local_greedy_search            heuristics.py:77-116
 *   local_greedy_search_count      heuristics.py:119-160   (+ number of rounds)
 *   local_greedy_search_stats      heuristics.py:163-209   (+ p2p and bst message counts)
 *   local_greedy_search_overhead   heuristics.py:212-263   (+ per-vertex overhead vector)
 *   local_greedy_search_nstep      heuristics.py:266-305   (stop after nstep rounds, return nb_is)
 *   greedy_search                  heuristics.py:13-35     (centralised greedy, distinct weights)
 *   dist_greedy_search             heuristics.py:38-74     (threshold greedy; scan order canonicalised, see dgs_oracle_run)
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imports the reference's own heuristics.py
 * (unmodified; three unused third-party imports stubbed) and stores its outputs; this file is
 * checked against those vectors in tests/test_oracle_golden.py.
 *
 * Graph input is CSR of the 0/1 symmetric zero-diagonal adjacency pattern (what
 * np.nonzero(adj[v]) enumerates at heuristics.py:94); weights are IEEE doubles compared with the
 * ordinary C operators, as numpy compares them.
 */
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * One call covers every variant.
 *   init_remain : NULL = every vertex starts in `remain` (heuristics.py:87); otherwise 0/1 per vertex
 *                 (used to restate the zero-weight removal of mwis_dqn_call.py:202-207 as a mask).
 *   nstep       : < 0 = run until `remain` is empty; otherwise at most nstep rounds (heuristics.py:279).
 *   member      : out, 1 for vertices in the returned set `mwis`.
 *   nb_is       : out (may be NULL), 1 for vertices in the reference's `nb_is` set.
 *   steps/p2p/bst/oh_vec : out (may be NULL), as returned by the _count/_stats/_overhead variants.
 * Returns the number of rounds executed, or -1 on allocation failure, or -2 when `max_rounds`
 * rounds passed without emptying `remain` (the reference would loop forever: NaN weights or
 * self-loops, see SURVEY.md section 7 "hard parts").
 */
long long lgs_oracle_run(int n, const long long *row_ptr, const int *col_idx, const double *wts,
                         const unsigned char *init_remain, int nstep, long long max_rounds,
                         unsigned char *member, unsigned char *nb_is, long long *steps,
                         long long *p2p, long long *bst, double *oh_vec)
{
    unsigned char *remain = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));
    unsigned char *nbis = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    if (!remain || !nbis) { free(remain); free(nbis); return -1; }
    long long n_remain = 0, n_member = 0;
    for (int v = 0; v < n; ++v) {
        remain[v] = init_remain ? (init_remain[v] != 0) : 1;
        n_remain += remain[v];
        member[v] = 0;
        if (oh_vec) oh_vec[v] = 0.0;
    }
    long long rounds = 0, c_p2p = 0, c_bst = 0;
    long long budget = nstep;
    while (n_remain > 0 && (nstep < 0 || budget > 0)) {
        if (max_rounds >= 0 && rounds >= max_rounds) { free(remain); free(nbis); return -2; }
        c_bst += n_remain;                                   /* heuristics.py:179 */
        /* decisions of this round read only the round-start `remain` (it is rebound after the
           vertex loop, heuristics.py:114) */
        for (int v = 0; v < n; ++v) {
            if (!remain[v]) continue;
            long long cnt = 0;
            double wbar = 0.0;
            int have = 0;
            for (long long e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
                int u = col_idx[e];
                if (!remain[u]) continue;                    /* nb_set ∩ remain, heuristics.py:95 */
                ++cnt;
                /* np.max propagates NaN (heuristics.py:102) */
                if (!have || wts[u] > wbar || wts[u] != wts[u]) { if (!(wbar != wbar)) wbar = wts[u]; have = 1; }
            }
            c_p2p += cnt;                                    /* heuristics.py:185 */
            if (oh_vec) oh_vec[v] += (double)cnt;            /* heuristics.py:238 */
            int join = 0, mute = 0;
            if (cnt == 0) {
                join = 1;                                    /* heuristics.py:96-98: no nb_is update */
            } else if (wts[v] > wbar) {
                join = 1; mute = 1;                          /* heuristics.py:103-105 */
            } else if (wts[v] == wbar) {
                /* smallest-index remaining neighbour whose weight equals wts[v], heuristics.py:107-109 */
                int nbv = -1;
                for (long long e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
                    int u = col_idx[e];
                    if (remain[u] && wts[u] == wts[v] && (nbv < 0 || u < nbv)) nbv = u;
                }
                if (v < nbv) { join = 1; mute = 1; }
            }
            if (join && !member[v]) { member[v] = 1; ++n_member; }
            if (mute) {
                for (long long e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
                    int u = col_idx[e];
                    if (remain[u]) nbis[u] = 1;              /* nb_is ∪= nb_set */
                }
                if (oh_vec) oh_vec[v] += 1.0;                /* "mute signaling", heuristics.py:249,256 */
            }
        }
        n_remain = 0;
        for (int v = 0; v < n; ++v) {                        /* remain - mwis - nb_is */
            if (remain[v] && (member[v] || nbis[v])) remain[v] = 0;
            n_remain += remain[v];
        }
        ++rounds;
        if (nstep >= 0) --budget;
    }
    c_bst += n_member;                                       /* heuristics.py:208 */
    if (nb_is) memcpy(nb_is, nbis, (size_t)n);
    if (steps) *steps = rounds;
    if (p2p) *p2p = c_p2p;
    if (bst) *bst = c_bst;
    free(remain);
    free(nbis);
    return rounds;
}

/* Batched form over a packed CSR (graph g owns vertices graph_ptr[g] .. graph_ptr[g+1]-1, column
 * indices are batch-global).  Per-graph outputs are arrays of length n_graphs (may be NULL). */
long long lgs_oracle_run_batch(int n_graphs, const long long *graph_ptr, const long long *row_ptr,
                               const int *col_idx, const double *wts, const unsigned char *init_remain,
                               int nstep, long long max_rounds, unsigned char *member,
                               unsigned char *nb_is, long long *steps, long long *p2p, long long *bst,
                               double *oh_vec)
{
    long long worst = 0;
    for (int g = 0; g < n_graphs; ++g) {
        long long v0 = graph_ptr[g], v1 = graph_ptr[g + 1];
        int n = (int)(v1 - v0);
        long long e0 = row_ptr[v0];
        long long nnz = row_ptr[v1] - e0;
        long long *rp = (long long *)malloc(sizeof(long long) * (size_t)(n + 1));
        int *ci = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
        if (!rp || !ci) { free(rp); free(ci); return -1; }
        for (int i = 0; i <= n; ++i) rp[i] = row_ptr[v0 + i] - e0;
        for (long long e = 0; e < nnz; ++e) ci[e] = (int)(col_idx[e0 + e] - v0);
        long long r = lgs_oracle_run(n, rp, ci, wts + v0, init_remain ? init_remain + v0 : NULL, nstep,
                                     max_rounds, member + v0, nb_is ? nb_is + v0 : NULL,
                                     steps ? steps + g : NULL, p2p ? p2p + g : NULL,
                                     bst ? bst + g : NULL, oh_vec ? oh_vec + v0 : NULL);
        free(rp);
        free(ci);
        if (r < 0) return r;
        if (r > worst) worst = r;
    }
    return worst;
}

/*
 * Centralised greedy (heuristics.py:13-35): visit vertices by descending weight; take a vertex
 * unless an already-taken vertex is its neighbour.  `order` is the visiting order supplied by the
 * caller (the reference uses np.argsort(-wts), whose tie order is unspecified - pass that same
 * permutation to compare like with like).
 */
void greedy_oracle_run(int n, const long long *row_ptr, const int *col_idx, const int *order,
                       unsigned char *member)
{
    unsigned char *blocked = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int v = 0; v < n; ++v) member[v] = 0;
    for (int k = 0; k < n; ++k) {
        int v = order[k];
        if (blocked[v]) continue;                            /* heuristics.py:28-29 */
        member[v] = 1;
        for (long long e = row_ptr[v]; e < row_ptr[v + 1]; ++e) blocked[col_idx[e]] = 1;
    }
    free(blocked);
}

/*
 * Threshold ("epsilon") distributed greedy, heuristics.py:38-74.  Per round over `remain`:
 *   seta  = remaining vertices without remaining neighbours, or with wts[v] >= max(wts[N(v) & remain]) / alpha
 *           (heuristics.py:54-63; alpha = 1 + epsilon / 3 is computed by the caller exactly as :46 does);
 *   mis_i = vertices of seta taken one after the other unless a neighbour was taken before (:64-69);
 *   mis_i joins the result, mis_i and all its neighbours leave `remain` (:70-71).
 * The reference walks seta in the iteration order of a CPython set, an implementation accident that
 * this restatement does not model: it walks seta in ASCENDING VERTEX ID.  The two agree whenever
 * no two vertices of a round's seta are adjacent (then mis_i = seta whatever the order);
 * `*order_free` reports whether that held in every round, and the parity fixtures compare with
 * the reference's output only on such instances (tests/golden/make_golden.py, "dgs_ref").
 * np.max / >= semantics: a NaN among the remaining neighbours' weights or in wts[v] keeps v out.
 * Returns the number of rounds, -1 on allocation failure, -2 when max_rounds passed (negative or
 * NaN weights can leave seta empty for ever; the reference then never returns).
 */
long long dgs_oracle_run(int n, const long long *row_ptr, const int *col_idx, const double *wts,
                         const unsigned char *init_remain, double alpha, long long max_rounds,
                         unsigned char *member, int *order_free)
{
    unsigned char *remain = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));
    unsigned char *seta = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));
    unsigned char *mis = (unsigned char *)malloc((size_t)(n > 0 ? n : 1));
    if (!remain || !seta || !mis) { free(remain); free(seta); free(mis); return -1; }
    long long n_remain = 0, rounds = 0;
    int free_order = 1;
    for (int v = 0; v < n; ++v) {
        remain[v] = init_remain ? (init_remain[v] != 0) : 1;
        n_remain += remain[v];
        member[v] = 0;
    }
    while (n_remain > 0) {
        if (max_rounds >= 0 && rounds >= max_rounds) { free(remain); free(seta); free(mis); return -2; }
        for (int v = 0; v < n; ++v) {
            seta[v] = 0;
            mis[v] = 0;
            if (!remain[v]) continue;
            int cnt = 0, has_nan = 0;
            double w_bar = 0.0;
            for (long long e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
                int u = col_idx[e];
                if (!remain[u]) continue;
                double wu = wts[u];
                if (wu != wu) has_nan = 1;
                if (cnt == 0 || wu > w_bar) w_bar = wu;
                ++cnt;
            }
            if (cnt == 0) seta[v] = 1;                                   /* heuristics.py:58-60 */
            else if (!has_nan && wts[v] >= w_bar / alpha) seta[v] = 1;   /* heuristics.py:61-63 */
        }
        for (int v = 0; v < n; ++v) {                                    /* ascending id, see above */
            if (!seta[v]) continue;
            int taken_nb = 0;
            for (long long e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
                int u = col_idx[e];
                if (seta[u]) free_order = 0;
                if (mis[u]) taken_nb = 1;
            }
            if (!taken_nb) mis[v] = 1;                                   /* heuristics.py:67-69 */
        }
        for (int v = 0; v < n; ++v) {
            if (!mis[v]) continue;
            member[v] = 1;
            if (remain[v]) { remain[v] = 0; --n_remain; }
            for (long long e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
                int u = col_idx[e];
                if (remain[u]) { remain[u] = 0; --n_remain; }            /* heuristics.py:69,71 */
            }
        }
        ++rounds;
    }
    if (order_free) *order_free = free_order;
    free(remain);
    free(seta);
    free(mis);
    return rounds;
}

#ifdef __cplusplus
}
#endif
