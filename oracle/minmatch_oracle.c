/* minmatch_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle/Makefile).
 *
 * Plain-C sequential restatement of the reference's greedy tree builder for one distance matrix,
 * MinMatch::QuickBuild, in the two forms BuildTopology uses without --sample_ages / without a template tree:
 *   - QuickBuild(d, tree, sample_ages)            src/tree_builder.cpp:1060-1303 (the `else` branch, :1234-1296)
 *   - QuickBuild(d, tree, sample_ages, d_prior)   src/tree_builder.cpp:2357-2646 (the `else` branch, :2538-2640)
 * with Initialize (:58-146 / :1646-1735), Coalesce (:295-598 / :1843-2070), InitializeSym (:254-293) and
 * CoalesceSym (:967-1058).  Citations are relative to /root/reference/include/.
 *
 * What has to be exact for identical trees: float arithmetic as x86-64 SSE evaluates it (no contraction; this file is
 * compiled with -ffp-contract=off), the order in which candidate pairs are met (one draw of
 * std::uniform_real_distribution<double>(0,1) on a std::mt19937 re-seeded with 1 per tree, libstdc++'s
 * generate_canonical: two 32-bit outputs per draw), the early `break` of the row-minimum rescan, and the state that
 * survives from one tree to the next inside one MinMatch object (min_values_CF is never reset, :2400-2401).
 *
 * Parity pinned: tests/test_minmatch_cpu.py compares the merge lists with oracle/_ref/qblens (the reference's own
 * MinMatch compiled from /root/reference) on random, tie-heavy and Relate-shaped matrices, several trees per object.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- std::mt19937 (Matsumoto & Nishimura 1998; the parameters of the C++ standard's mt19937) ---- */
typedef struct {
    uint32_t s[624];
    int pos;
} mt_t;

static void mt_seed(mt_t *g, uint32_t seed)
{
    g->s[0] = seed;
    for (int i = 1; i < 624; i++) g->s[i] = 1812433253u * (g->s[i - 1] ^ (g->s[i - 1] >> 30)) + (uint32_t)i;
    g->pos = 624;
}

static uint32_t mt_next(mt_t *g)
{
    if (g->pos >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (g->s[i] & 0x80000000u) | (g->s[(i + 1) % 624] & 0x7fffffffu);
            g->s[i] = g->s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->pos = 0;
    }
    uint32_t y = g->s[g->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* libstdc++ generate_canonical<double,53>(mt19937): sum = g1 + g2*2^32 (rounded to double), / 2^64 */
static double mt_unif(mt_t *g)
{
    double lo = (double)mt_next(g);
    double hi = (double)mt_next(g);
    double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

/* ---- state of one MinMatch object (tree_builder.hpp:41-108) ---- */
typedef struct {
    int a, b;        /* lin1, lin2 */
    double dist, tie; /* dist, dist2 */
} cand_t;

typedef struct mm_oracle {
    int N;
    float thr, thr_cf;
    float *minv, *minv_cf, *minv_sym, *size, *sym;
    int *act, n_act, *conv, *upd;
    cand_t *cand, *cand_sym, best, best_sym;
    mt_t rng;
    long draws;
} mm_oracle;

static const double INF = (double)INFINITY;

mm_oracle *mmo_create(int N, double theta)
{
    mm_oracle *m = (mm_oracle *)calloc(1, sizeof(*m));
    m->N = N;
    m->thr = (float)(-0.2 * log(theta / (1.0 - theta)));    /* :43 */
    m->thr_cf = (float)(-0.001 * log(theta / (1.0 - theta))); /* :44 */
    m->minv = (float *)calloc(N, sizeof(float));
    m->minv_cf = (float *)calloc(N, sizeof(float)); /* zero-filled by resize(), and kept between trees */
    m->minv_sym = (float *)calloc(N, sizeof(float));
    m->size = (float *)calloc(N, sizeof(float));
    m->act = (int *)calloc(N, sizeof(int));
    m->conv = (int *)calloc(N, sizeof(int));
    m->upd = (int *)calloc(N, sizeof(int));
    m->cand = (cand_t *)calloc(N, sizeof(cand_t));
    m->cand_sym = (cand_t *)calloc(N, sizeof(cand_t));
    for (int k = 0; k < N; k++) {
        m->cand[k].a = m->cand[k].b = m->cand_sym[k].a = m->cand_sym[k].b = -1;
        m->cand[k].dist = m->cand[k].tie = m->cand_sym[k].dist = m->cand_sym[k].tie = INF;
    }
    return m;
}

void mmo_destroy(mm_oracle *m)
{
    if (!m) return;
    free(m->minv); free(m->minv_cf); free(m->minv_sym); free(m->size); free(m->sym);
    free(m->act); free(m->conv); free(m->upd); free(m->cand); free(m->cand_sym);
    free(m);
}

static int better(const cand_t *c, double dist, double tie) { return c->dist > dist || (c->dist == dist && c->tie > tie); }

/* the weight of a feasible pair: with a prior matrix, 0 if the pair is also mutually minimal there, else d+d^T */
static float pair_weight(const mm_oracle *m, const float *d, const float *cf, int x, int y)
{
    const int N = m->N;
    if (cf) {
        float w = (float)(1 - (cf[(size_t)x * N + y] <= m->minv_cf[x]) * (cf[(size_t)y * N + x] <= m->minv_cf[y]));
        if (!(w > 0)) return w;
    }
    return d[(size_t)x * N + y] + d[(size_t)y * N + x];
}

/* a feasible pair met in the scan: one draw, then both members keep the better of their candidate and this pair */
static void meet(mm_oracle *m, const float *d, const float *cf, int first, int second, int x, int y)
{
    float w = pair_weight(m, d, cf, x, y);
    double u = mt_unif(&m->rng);
    m->draws++;
    if (better(&m->cand[x], w, u)) { m->cand[x].a = first; m->cand[x].b = second; m->cand[x].dist = w; m->cand[x].tie = u; }
    if (better(&m->cand[y], w, u)) { m->cand[y].a = first; m->cand[y].b = second; m->cand[y].dist = w; m->cand[y].tie = u; }
}

/* Initialize: :58-146 (tmpl_tree == NULL) and :1646-1735 */
static void initialize(mm_oracle *m, const float *d, const float *cf)
{
    const int N = m->N;
    for (int p = 0; p < m->n_act; p++) {
        int k = m->act[p];
        m->cand[k].dist = m->cand[k].tie = INF;
        float v = m->minv[p];
        for (int q = 0; q < m->n_act; q++)
            if (v > d[(size_t)k * N + q] && m->act[q] != k) v = d[(size_t)k * N + q];
        m->minv[p] = v + m->thr;
    }
    if (cf)
        for (int p = 0; p < m->n_act; p++) {
            int k = m->act[p];
            float v = m->minv_cf[p]; /* starts from whatever the previous tree left */
            for (int q = 0; q < m->n_act; q++)
                if (v > cf[(size_t)k * N + q] && m->act[q] != k) v = cf[(size_t)k * N + q];
            m->minv_cf[p] = v + m->thr_cf;
        }
    for (int p = 0; p < m->n_act; p++) {
        int x = m->act[p];
        for (int q = p + 1; q < m->n_act; q++) {
            int y = m->act[q];
            if (m->minv[p] >= d[(size_t)x * N + y] && m->minv[q] >= d[(size_t)y * N + x]) {
                float w = pair_weight(m, d, cf, x, y);
                meet(m, d, cf, x, y, x, y);
                if (better(&m->best, m->cand[y].dist, m->cand[y].tie)) { /* :135-141: pair of the moment, tie of cand[y] */
                    m->best.a = x; m->best.b = y; m->best.dist = w; m->best.tie = m->cand[y].tie;
                }
            }
        }
    }
}

/* Coalesce(i, j): :295-598 (tmpl_tree == NULL) and :1843-2070 */
static void coalesce(mm_oracle *m, float *d, const float *cf, int i, int j)
{
    const int N = m->N;
    const float si = m->size[i], sj = m->size[j], sum = si + sj;
    float *di = d + (size_t)i * N, *dj = d + (size_t)j * N;
    float min_j = INFINITY;
    int n_upd = 0;
    m->best.dist = m->best.tie = INF;
    for (int p = 0; p < m->n_act; p++) {
        const int k = m->act[p];
        if (k == i || k == j) continue;
        float *dk = d + (size_t)k * N;
        const float dkj = dk[j], dki = dk[i], dik = di[k], djk = dj[k];
        float mk = m->minv[k];
        if (dik != djk) dj[k] = (si * dik + sj * djk) / sum;
        if (dki != dkj) dk[j] = (si * dki + sj * dkj) / sum;
        int min_changed = 0;
        if (dkj != dki && (fabsf(mk - m->thr - dkj) < 1e-4 || fabsf(mk - m->thr - dki) < 1e-4)) {
            const float old = mk - m->thr;
            mk = INFINITY;
            min_changed = 1;
            for (int q = 0; q < m->n_act; q++) {
                int l = m->act[q];
                if (l == i || l == k) continue;
                if (mk > dk[l]) {
                    mk = dk[l];
                    if (mk == old) break; /* :337-339 */
                }
            }
            mk += m->thr;
            m->minv[k] = mk;
        }
        cand_t *ck = &m->cand[k];
        const int touches = ck->a == j || ck->b == j || ck->a == i || ck->b == i;
        if (dkj != dki || djk != dik || touches) {
            if (min_changed || touches) {
                m->upd[n_upd++] = k;
                ck->dist = ck->tie = INF;
                for (int q = 0; q < p; q++) {
                    int l = m->act[q];
                    if (dk[l] <= mk && l != j && l != i && d[(size_t)l * N + k] <= m->minv[l]) meet(m, d, cf, k, l, k, l);
                }
            } else {
                for (int q = 0; q < n_upd; q++) {
                    int l = m->upd[q];
                    if (dk[l] <= mk && d[(size_t)l * N + k] <= m->minv[l]) meet(m, d, cf, k, l, l, k);
                }
            }
        } else {
            if (ck->a == i) ck->a = j;
            if (ck->b == i) ck->b = j;
            for (int q = 0; q < n_upd; q++) {
                int l = m->upd[q];
                if (dk[l] <= mk && d[(size_t)l * N + k] <= m->minv[l]) meet(m, d, cf, k, l, l, k);
            }
        }
        if (better(&m->best, ck->dist, ck->tie)) m->best = *ck;
        if (dj[k] < min_j) min_j = dj[k];
    }
    min_j += m->thr;
    m->minv[j] = min_j;
    m->cand[j].dist = m->cand[j].tie = INF;
    for (int p = 0; p < m->n_act; p++) {
        const int k = m->act[p];
        if (dj[k] <= min_j && d[(size_t)k * N + j] <= m->minv[k] && k != i && k != j) meet(m, d, cf, k, j, k, j);
    }
    if (better(&m->best, m->cand[j].dist, m->cand[j].tie)) m->best = m->cand[j];
}

/* InitializeSym: :254-293 */
static void initialize_sym(mm_oracle *m, const float *d)
{
    const int N = m->N;
    float *s = m->sym;
    for (int p = 0; p < m->n_act; p++)
        for (int q = p + 1; q < m->n_act; q++) {
            int x = m->act[p], y = m->act[q];
            s[(size_t)x * N + y] = d[(size_t)x * N + y] + d[(size_t)y * N + x];
            s[(size_t)y * N + x] = s[(size_t)x * N + y];
        }
    for (int p = 0; p < m->n_act; p++) {
        int x = m->act[p];
        m->cand_sym[x].dist = INF;
        for (int q = 0; q < m->n_act; q++) {
            int l = m->act[q];
            if (m->minv_sym[x] > s[(size_t)x * N + l] && l != x) {
                m->minv_sym[x] = s[(size_t)x * N + l];
                if (m->cand_sym[x].dist > m->minv_sym[x]) { m->cand_sym[x].a = x; m->cand_sym[x].b = l; m->cand_sym[x].dist = m->minv_sym[x]; }
                if (m->best_sym.dist > m->cand_sym[x].dist) { m->best_sym.a = x; m->best_sym.b = l; m->best_sym.dist = m->minv_sym[x]; }
            }
        }
    }
}

/* CoalesceSym(i, j): :967-1058 */
static void coalesce_sym(mm_oracle *m, int i, int j)
{
    const int N = m->N;
    float *s = m->sym;
    const float si = m->size[i], sj = m->size[j], sum = si + sj;
    float *row_i = s + (size_t)i * N, *row_j = s + (size_t)j * N;
    float min_j = INFINITY;
    m->best_sym.dist = INF;
    m->cand_sym[j].dist = INF;
    for (int p = 0; p < m->n_act; p++) {
        const int k = m->act[p];
        if (k == i || k == j) continue;
        float *row_k = s + (size_t)k * N;
        const float dkj = row_k[j], dki = row_k[i], dik = row_i[k], djk = row_j[k];
        float mk = m->minv_sym[k];
        if (dik != djk) row_j[k] = (si * dik + sj * djk) / sum;
        if (dki != dkj) row_k[j] = (si * dki + sj * dkj) / sum;
        if (dkj != dki) {
            if (fabsf(mk - dkj) < 1e-6 || fabsf(mk - dki) < 1e-6) {
                const float old = mk;
                mk = INFINITY;
                m->cand_sym[k].dist = INF;
                for (int q = 0; q < m->n_act; q++) {
                    int l = m->act[q];
                    if (l == i || l == k) continue;
                    if (mk > row_k[l]) {
                        mk = row_k[l];
                        if (m->cand_sym[k].dist > mk) { m->cand_sym[k].a = k; m->cand_sym[k].b = l; m->cand_sym[k].dist = mk; }
                        if (mk == old) break;
                    }
                }
                m->minv_sym[k] = mk;
            }
        } else {
            if (m->cand_sym[k].a == i) m->cand_sym[k].a = j;
            if (m->cand_sym[k].b == i) m->cand_sym[k].b = j;
        }
        if (m->best_sym.dist > m->cand_sym[k].dist) m->best_sym = m->cand_sym[k];
        if (row_j[k] < min_j) {
            min_j = row_j[k];
            if (m->cand_sym[j].dist > row_j[k]) { m->cand_sym[j].a = k; m->cand_sym[j].b = j; m->cand_sym[j].dist = row_j[k]; }
        }
    }
    m->minv_sym[j] = min_j;
    if (m->best_sym.dist > m->cand_sym[j].dist) m->best_sym = m->cand_sym[j];
}

/* QuickBuild.  d: N*N floats, modified in place as the reference does.  d_prior: N*N floats or NULL (the three-argument
 * form).  merges: 2*(N-1) ints, (left, right) = (convert_index[i], convert_index[j]) of node N+t.  info (may be NULL):
 * info[0] = number of draws, info[1] = first merge step that had no candidate (-1: none).  Returns 0. */
int mmo_quickbuild(mm_oracle *m, float *d, const float *d_prior, int *merges, long *info)
{
    const int N = m->N;
    float *cf = NULL;
    if (d_prior) {
        cf = (float *)malloc((size_t)N * N * sizeof(float));
        memcpy(cf, d_prior, (size_t)N * N * sizeof(float));
    }
    mt_seed(&m->rng, 1);
    m->draws = 0;
    m->n_act = N;
    for (int k = 0; k < N; k++) {
        m->act[k] = k;
        m->conv[k] = k;
        m->size[k] = 1.0f;
        m->minv[k] = INFINITY;
        m->minv_sym[k] = INFINITY;
    }
    m->best.dist = m->best.tie = INF;
    m->best_sym.dist = INF;
    initialize(m, d, cf);
    int use_sym = 0;
    long first_sym = -1;
    for (int node = N; node < 2 * N - 1; node++) {
        int i, j;
        if (m->best.dist == INF) { /* :1244-1256: no mutually minimal pair left */
            if (!use_sym) {
                if (!m->sym) m->sym = (float *)calloc((size_t)N * N, sizeof(float));
                initialize_sym(m, d);
                use_sym = 1;
                first_sym = node - N;
            }
            i = m->best_sym.a;
            j = m->best_sym.b;
        } else {
            i = m->best.a;
            j = m->best.b;
        }
        merges[2 * (node - N)] = m->conv[i];
        merges[2 * (node - N) + 1] = m->conv[j];
        if (cf) { /* :2583-2606 */
            const float si = m->size[i], sj = m->size[j], sum = si + sj;
            float *ci = cf + (size_t)i * N, *cj = cf + (size_t)j * N;
            m->minv_cf[j] = INFINITY;
            for (int p = 0; p < m->n_act; p++) {
                const int k = m->act[p];
                if (k == i || k == j) continue;
                float *ck = cf + (size_t)k * N;
                const float dkj = ck[j], dki = ck[i], dik = ci[k], djk = cj[k];
                if (dik != djk) cj[k] = (si * dik + sj * djk) / sum;
                if (dki != dkj) ck[j] = (si * dki + sj * dkj) / sum;
                if (m->minv_cf[j] > cj[k]) m->minv_cf[j] = cj[k];
            }
            m->minv_cf[j] += m->thr_cf;
        }
        coalesce(m, d, cf, i, j);
        if (use_sym) coalesce_sym(m, i, j);
        m->size[j] = m->size[i] + m->size[j];
        m->conv[j] = node;
        int p = 0;
        while (m->act[p] != i) p++;
        memmove(m->act + p, m->act + p + 1, (size_t)(m->n_act - p - 1) * sizeof(int));
        m->n_act--;
    }
    free(cf);
    if (info) { info[0] = m->draws; info[1] = first_sym; }
    return 0;
}
