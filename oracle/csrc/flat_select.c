/* CPU ORACLE / CPU BASELINE (test infrastructure, NOT product code).
 *
 * k-selection of an exhaustive inner-product search the way faiss-cpu 1.7.4 does it for k >= 100
 * (reference call site: src/serving/retrieval.py:171 `self.index.search(q, k_search)` on an
 * `IndexFlatIP`, requirements.txt:13 faiss-cpu==1.7.4; faiss itself is not vendored and not installable
 * offline, so this restates its published algorithm): `exhaustive_inner_product_blas` runs an sgemm over a
 * (query block x database block) tile and hands the tile to a RESERVOIR result handler
 * (`ReservoirTopN`, capacity 2k per query): a score enters the reservoir only if it is strictly greater
 * than the query's threshold; a full reservoir is shrunk to its k best by a partition and the threshold
 * becomes the k-th best score; at the end the reservoir is shrunk once more and sorted, best first.
 *
 * The sgemm stays with MKL (torch-CPU); this file is the result handler.  Order is made total
 * (score descending, row id ascending) so results are deterministic; database blocks are fed in
 * increasing row order, so "strictly greater than the threshold" keeps the lower row id at the k boundary,
 * as faiss does.
 *
 * Built by oracle/build_c.py into oracle/_build/libflatselect.so (git-ignored); used by
 * oracle/flat_ip.search_reservoir only (tests + bench.py's cpu_baseline / --impl reference legs).
 */
#include <float.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define CHUNK 64

static inline int better(float sa, int64_t ia, float sb, int64_t ib) {
    return sa > sb || (sa == sb && ia < ib);
}

static inline void swap_pair(float* s, int64_t* id, int a, int b) {
    float ts = s[a]; s[a] = s[b]; s[b] = ts;
    int64_t ti = id[a]; id[a] = id[b]; id[b] = ti;
}

/* Quickselect: afterwards positions [0, k) hold the k best pairs (any order). */
static void select_k_best(float* s, int64_t* id, int n, int k) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        /* median of three as the pivot, parked at hi */
        if (better(s[mid], id[mid], s[lo], id[lo])) swap_pair(s, id, mid, lo);
        if (better(s[hi], id[hi], s[lo], id[lo])) swap_pair(s, id, hi, lo);
        if (better(s[mid], id[mid], s[hi], id[hi])) swap_pair(s, id, mid, hi);
        float ps = s[hi]; int64_t pi = id[hi];
        int store = lo;
        for (int j = lo; j < hi; ++j)
            if (better(s[j], id[j], ps, pi)) { swap_pair(s, id, j, store); ++store; }
        swap_pair(s, id, store, hi);
        if (store == k - 1 || store == k) return;   /* [0, k) are the k best either way */
        if (store < k - 1) lo = store + 1; else hi = store - 1;
    }
}

static float worst_of(const float* s, int n) {
    float m = s[0];
    for (int j = 1; j < n; ++j) m = s[j] < m ? s[j] : m;
    return m;
}

/* One (query block x database block) score tile: scores[q * ld + j] is query q against row row0 + j.
 * res_s / res_i: [nq, 2k] reservoirs, res_n: [nq] fill counts, thr: [nq] thresholds (start at -FLT_MAX).
 * `threads` is explicit because torchrun exports OMP_NUM_THREADS=1. */
void flat_select_add_block(const float* scores, int64_t ld, int64_t nq, int64_t w, int64_t row0, int32_t k,
                           float* res_s, int64_t* res_i, int32_t* res_n, float* thr, int32_t threads) {
    const int cap = 2 * k;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t q = 0; q < nq; ++q) {
        const float* s = scores + q * ld;
        float* rs = res_s + q * cap;
        int64_t* ri = res_i + q * cap;
        int n = res_n[q];
        float t = thr[q];
        for (int64_t j0 = 0; j0 < w; j0 += CHUNK) {
            const int64_t jn = (w - j0 < CHUNK) ? (w - j0) : CHUNK;
            float m = s[j0];
            for (int64_t j = 1; j < jn; ++j) m = s[j0 + j] > m ? s[j0 + j] : m;   /* vectorises */
            if (!(m > t)) continue;
            for (int64_t j = 0; j < jn; ++j) {
                const float v = s[j0 + j];
                if (v > t) {
                    rs[n] = v; ri[n] = row0 + j0 + j; ++n;
                    if (n == cap) {
                        select_k_best(rs, ri, n, k);
                        n = k;
                        t = worst_of(rs, k);
                    }
                }
            }
        }
        res_n[q] = n;
        thr[q] = t;
    }
}

typedef struct { float s; int64_t id; } pair_t;

static int cmp_pair(const void* a, const void* b) {
    const pair_t* x = (const pair_t*)a; const pair_t* y = (const pair_t*)b;
    if (better(x->s, x->id, y->s, y->id)) return -1;
    if (better(y->s, y->id, x->s, x->id)) return 1;
    return 0;
}

/* Final shrink + sort: D [nq, k] descending, I [nq, k]; unfilled slots carry -FLT_MAX / -1. */
void flat_select_finish(int64_t nq, int32_t k, float* res_s, int64_t* res_i, const int32_t* res_n,
                        float* D, int64_t* I, int32_t threads) {
    const int cap = 2 * k;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t q = 0; q < nq; ++q) {
        float* rs = res_s + q * cap;
        int64_t* ri = res_i + q * cap;
        int n = res_n[q];
        if (n > k) { select_k_best(rs, ri, n, k); n = k; }
        pair_t tmp[n > 0 ? n : 1];
        for (int j = 0; j < n; ++j) { tmp[j].s = rs[j]; tmp[j].id = ri[j]; }
        qsort(tmp, (size_t)n, sizeof(pair_t), cmp_pair);
        for (int j = 0; j < k; ++j) {
            D[q * k + j] = j < n ? tmp[j].s : -FLT_MAX;
            I[q * k + j] = j < n ? tmp[j].id : -1;
        }
    }
}
