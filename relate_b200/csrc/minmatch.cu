// minmatch.cu — the reference's greedy tree builder on the GPU.
//
// MinMatch::QuickBuild (src/tree_builder.cpp:1060-1303 and, with a prior matrix, :2357-2646; Initialize :58-146 /
// :1646-1735, Coalesce :295-598 / :1843-2070, InitializeSym :254-293, CoalesceSym :967-1058) merges the two clusters of
// the best mutually-minimal pair N-1 times.  The reference walks the clusters one after the other and draws one random
// number per feasible pair it meets, so the tree depends on the ORDER of the walk.  Here one CTA builds one tree and
// every merge step is a handful of data-parallel phases over the active clusters that reproduce that order exactly:
//
//   A  update row/column j of d (and of the prior matrix) from rows/columns i and j; row j's new minimum; which rows
//      have to look for a new minimum (their old one was d[k][i] or d[k][j]); which candidates name i or j
//   B  one warp per such row: the new minimum, with the reference's early `break` decided from three reductions
//      (smallest value, first position equal to the old minimum, first position below it)
//   C  U = rows whose minimum moved or whose candidate named i or j (the reference's `updated_cluster`), in order
//   D  count feasible pairs per row — rows of U against every earlier row, the other rows against the earlier members
//      of U, and row j against everybody, which is exactly the set and the order the reference meets them in —,
//      prefix-sum the counts: a pair's rank is the index of its random draw
//   E  write the pairs, give pair r the r-th next output of a device-side std::mt19937 (seeded with 1 per tree, two
//      32-bit outputs per double as libstdc++'s generate_canonical does), and let both members keep the
//      lexicographically smallest (weight, draw) among their old candidate and their new pairs: three rounds of atomicMin
//   F  the best candidate over all clusters = the next (i, j)
// plus the same for the symmetric fallback matrix once no mutually-minimal pair is left.  All float arithmetic uses
// explicit round-to-nearest intrinsics (no FMA contraction), so the values equal the reference's x86-64 SSE results and
// the merge lists are identical (tests/test_minmatch_gpu.py: against oracle/minmatch_oracle.c and against the
// reference's own MinMatch through oracle/_ref/qblens).
//
// State that survives from tree to tree inside one reference MinMatch object — min_values_CF (never reset) and the
// lin1/lin2 of candidates whose distance was reset — lives in the rp_minmatch handle for the same reason.
#include "../../include/relate_paint.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace rp {
int api_fail(int code, const std::string &msg); // paint_api.cu: sets rp_last_error()
}

namespace {

constexpr int MM_THREADS = 1024;
constexpr int MM_WARPS = MM_THREADS / 32;
constexpr unsigned FINF = 0x7f800000u;            // float +inf
constexpr unsigned long long DINF = 0x7ff0000000000000ull; // double +inf

struct MMState {
    int N;
    float thr, thr_cf;
    // matrices (row-major N x N)
    float *d, *cf, *sym;
    // persistent between trees
    float *minv_cf;
    int *cand_a, *cand_b;
    unsigned *cand_dist;           // float bits (non-negative: bit order == value order)
    unsigned long long *cand_tie;  // double bits
    int *csym_a, *csym_b;
    float *csym_dist;
    // per tree
    float *minv, *minv_sym, *size;
    int *conv, *act;
    int *flag;     // bit 0: candidate names i or j, bit 1: minimum recomputed, bit 2: member of U
    int *rescan;   // positions whose row needs a new minimum
    int *ulist;    // positions of U, ascending
    int *cnt;      // pairs per row (n_act + 1 entries), then exclusive offsets
    unsigned *key1;
    unsigned long long *key2;
    // pair buffer
    int cap;
    int *pa, *pb;
    float *pw;
    unsigned long long *ptie;
    // output
    int *merges;
    long long *info; // [0] draws, [1] first step without a candidate (-1), [2] steps on the fallback
};

struct MMShared {
    unsigned mt[624];
    float redf[MM_WARPS];
    unsigned redu[MM_WARPS];
    unsigned long long redull[MM_WARPS];
    int redi[MM_WARPS];
    int scan[MM_WARPS];
    int n_rescan, n_u, total, i, j, n_act, p_end;
    float bf;
    unsigned bu;
    unsigned long long bull;
    int bi;
};

__device__ __forceinline__ float mix(float si, float a, float sj, float b, float sum)
{
    return __fdiv_rn(__fadd_rn(__fmul_rn(si, a), __fmul_rn(sj, b)), sum);
}

__device__ __forceinline__ unsigned temper(unsigned y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// the next 624 words of std::mt19937 (three dependent thirds)
__device__ void mt_twist(MMShared &sh)
{
    const int t = threadIdx.x;
    const int lo[3] = {0, 227, 454}, hi[3] = {227, 454, 624};
    for (int ph = 0; ph < 3; ph++) {
        unsigned v = 0;
        const int idx = lo[ph] + t;
        if (idx < hi[ph]) {
            const unsigned y = (sh.mt[idx] & 0x80000000u) | (sh.mt[(idx + 1) % 624] & 0x7fffffffu);
            v = sh.mt[(idx + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncthreads();
        if (idx < hi[ph]) sh.mt[idx] = v;
        __syncthreads();
    }
}

__device__ __forceinline__ float block_min(float v, MMShared &sh)
{
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(~0u, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh.redf[threadIdx.x >> 5] = v;
    __syncthreads();
    v = sh.redf[threadIdx.x & 31];
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(~0u, v, o));
    return v;
}

// lexicographic minimum of (a, b, c) over the block; every thread gets the winner
__device__ __forceinline__ void block_lexmin(unsigned &a, unsigned long long &b, int &c, MMShared &sh)
{
    auto take = [&](unsigned oa, unsigned long long ob, int oc) {
        if (oa < a || (oa == a && (ob < b || (ob == b && oc < c)))) { a = oa; b = ob; c = oc; }
    };
    for (int o = 16; o; o >>= 1) take(__shfl_xor_sync(~0u, a, o), __shfl_xor_sync(~0u, b, o), __shfl_xor_sync(~0u, c, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sh.redu[threadIdx.x >> 5] = a; sh.redull[threadIdx.x >> 5] = b; sh.redi[threadIdx.x >> 5] = c; }
    __syncthreads();
    a = sh.redu[threadIdx.x & 31]; b = sh.redull[threadIdx.x & 31]; c = sh.redi[threadIdx.x & 31];
    for (int o = 16; o; o >>= 1) take(__shfl_xor_sync(~0u, a, o), __shfl_xor_sync(~0u, b, o), __shfl_xor_sync(~0u, c, o));
}

// exclusive prefix sum of arr[0..n) in place; returns the total to every thread
__device__ int block_exscan(int *arr, int n, MMShared &sh)
{
    const int t = threadIdx.x, per = (n + MM_THREADS - 1) / MM_THREADS;
    const int b = min(t * per, n), e = min(b + per, n);
    int s = 0;
    for (int q = b; q < e; q++) s += arr[q];
    int inc = s;
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(~0u, inc, o);
        if ((t & 31) >= o) inc += u;
    }
    __syncthreads();
    if ((t & 31) == 31) sh.scan[t >> 5] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
    for (int w = 0; w < MM_WARPS; w++) {
        int v = sh.scan[w];
        if (w < (t >> 5)) wbase += v;
        total += v;
    }
    int run = wbase + inc - s;
    for (int q = b; q < e; q++) { int v = arr[q]; arr[q] = run; run += v; }
    __syncthreads();
    return total;
}

// weight of a feasible pair (x, y): with a prior matrix 0 if the pair is mutually minimal there too, else d + d^T
__device__ __forceinline__ float pair_weight(const MMState &s, bool has_cf, int x, int y)
{
    const size_t N = s.N;
    if (has_cf && s.cf[x * N + y] <= s.minv_cf[x] && s.cf[y * N + x] <= s.minv_cf[y]) return 0.0f;
    return __fadd_rn(s.d[x * N + y], s.d[y * N + x]);
}

enum { ROW_SKIP = 0, ROW_BELOW = 1, ROW_LIST = 2, ROW_ABOVE = 3 };

// Feasible pairs of row p in the reference's order.  EMIT = false: count them; EMIT = true: write them at `base`.
// Warp-wide for ROW_BELOW / ROW_ABOVE / the j row (all lanes call it), returns the count to every lane.
template <bool EMIT>
__device__ int row_pairs_warp(const MMState &s, bool has_cf, int n_act, int p, int kind, int ci, int cj, int base)
{
    const size_t N = s.N;
    const int lane = threadIdx.x & 31;
    const bool jrow = p == n_act;
    const int x = jrow ? cj : s.act[p];
    const float mx = s.minv[x];
    const float *dx = s.d + x * N;
    const int q0 = kind == ROW_ABOVE ? p + 1 : 0, q1 = (kind == ROW_BELOW) ? p : n_act;
    int count = 0;
    for (int qb = q0; qb < q1; qb += 32) {
        const int q = qb + lane;
        bool ok = false;
        int y = -1;
        if (q < q1) {
            y = s.act[q];
            ok = y != ci && y != cj && dx[y] <= mx && s.d[y * N + x] <= s.minv[y];
        }
        const unsigned m = __ballot_sync(~0u, ok);
        if (EMIT && ok) {
            const int r = base + count + __popc(m & ((1u << lane) - 1));
            // candidate orientation: (row, partner), except for the new cluster's row: (partner, j)   (:565-573)
            s.pa[r] = jrow ? y : x;
            s.pb[r] = jrow ? x : y;
            s.pw[r] = pair_weight(s, has_cf, x, y);
        }
        count += __popc(m);
    }
    return count;
}

// a row outside U meets the earlier members of U only
template <bool EMIT>
__device__ int row_pairs_list(const MMState &s, bool has_cf, int p, int n_u, int base)
{
    const size_t N = s.N;
    const int x = s.act[p];
    const float mx = s.minv[x];
    const float *dx = s.d + x * N;
    int count = 0;
    for (int u = 0; u < n_u; u++) {
        const int pu = s.ulist[u];
        if (pu >= p) break;
        const int y = s.act[pu];
        if (dx[y] <= mx && s.d[y * N + x] <= s.minv[y]) {
            if (EMIT) {
                s.pa[base + count] = x;
                s.pb[base + count] = y;
                s.pw[base + count] = pair_weight(s, has_cf, x, y);
            }
            count++;
        }
    }
    return count;
}

// Phases D + E for the rows described by flag[] (init: every row looks above itself): count, rank, draw, apply.
__device__ void meet_pairs(const MMState &s, MMShared &sh, bool has_cf, bool init, int n_act, int ci, int cj, int n_u,
                           int &rng_pos, long long &draws)
{
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int rows = init ? n_act : n_act + 1; // + the new cluster's row
    // D: counts
    for (int p = warp; p < rows; p += MM_WARPS) {
        int kind;
        if (init) kind = ROW_ABOVE;
        else if (p == n_act) kind = ROW_LIST + 100; // j row: everybody
        else {
            const int k = s.act[p];
            kind = (k == ci || k == cj) ? ROW_SKIP : ((s.flag[k] & 4) ? ROW_BELOW : ROW_LIST);
        }
        if (kind == ROW_SKIP) { if (lane == 0) s.cnt[p] = 0; continue; }
        if (kind == ROW_LIST) continue; // thread-level below
        const int c = row_pairs_warp<false>(s, has_cf, n_act, p, kind, ci, cj, 0);
        if (lane == 0) s.cnt[p] = c;
    }
    if (!init)
        for (int p = t; p < n_act; p += MM_THREADS) {
            const int k = s.act[p];
            if (k == ci || k == cj || (s.flag[k] & 4)) continue;
            s.cnt[p] = n_u ? row_pairs_list<false>(s, has_cf, p, n_u, 0) : 0;
        }
    __syncthreads();
    const int total = block_exscan(s.cnt, rows, sh);
    if (total == 0) return;
    // E: segments of at most cap pairs (one segment unless a tie-heavy matrix meets a small buffer)
    int p0 = 0;
    while (p0 < rows) {
        // last row of the segment
        if (t == 0) sh.p_end = rows;
        __syncthreads();
        const int off0 = s.cnt[p0];
        for (int p = p0 + t; p < rows; p += MM_THREADS) {
            const int endp = (p + 1 < rows) ? s.cnt[p + 1] : total;
            if (endp - off0 > s.cap && s.cnt[p] - off0 <= s.cap) sh.p_end = max(p, p0 + 1);
        }
        __syncthreads();
        const int p1 = sh.p_end;
        const int seg_total = ((p1 < rows) ? s.cnt[p1] : total) - off0;
        __syncthreads();
        if (seg_total > 0) {
            for (int p = p0 + warp; p < p1; p += MM_WARPS) {
                int kind;
                if (init) kind = ROW_ABOVE;
                else if (p == n_act) kind = ROW_LIST + 100;
                else {
                    const int k = s.act[p];
                    kind = (k == ci || k == cj) ? ROW_SKIP : ((s.flag[k] & 4) ? ROW_BELOW : ROW_LIST);
                }
                if (kind == ROW_SKIP || kind == ROW_LIST) continue;
                row_pairs_warp<true>(s, has_cf, n_act, p, kind, ci, cj, s.cnt[p] - off0);
            }
            if (!init && n_u)
                for (int p = p0 + t; p < min(p1, n_act); p += MM_THREADS) {
                    const int k = s.act[p];
                    if (k == ci || k == cj || (s.flag[k] & 4)) continue;
                    row_pairs_list<true>(s, has_cf, p, n_u, s.cnt[p] - off0);
                }
            // draws, in rank order
            int cur = 0;
            while (cur < seg_total) {
                if (rng_pos == 312) { mt_twist(sh); rng_pos = 0; }
                const int n = min(312 - rng_pos, seg_total - cur);
                for (int r = t; r < n; r += MM_THREADS) {
                    const unsigned g1 = temper(sh.mt[2 * (rng_pos + r)]), g2 = temper(sh.mt[2 * (rng_pos + r) + 1]);
                    double u = __dmul_rn(__dadd_rn(__uint2double_rn(g1), __dmul_rn(__uint2double_rn(g2), 4294967296.0)),
                                         5.42101086242752217003726400434970855712890625e-20);
                    unsigned long long ub = (unsigned long long)__double_as_longlong(u);
                    if (u >= 1.0) ub = 0x3fefffffffffffffull;
                    s.ptie[cur + r] = ub;
                }
                rng_pos += n;
                cur += n;
                __syncthreads();
            }
            draws += seg_total;
            // apply: both members keep the smallest (weight, draw) among their candidate and the new pairs
            for (int p = t; p < n_act; p += MM_THREADS) { const int k = s.act[p]; __stcg(&s.key1[k], s.cand_dist[k]); }
            __syncthreads();
            for (int r = t; r < seg_total; r += MM_THREADS) {
                const unsigned wb = __float_as_uint(s.pw[r]);
                atomicMin(&s.key1[s.pa[r]], wb);
                atomicMin(&s.key1[s.pb[r]], wb);
            }
            __syncthreads();
            for (int p = t; p < n_act; p += MM_THREADS) {
                const int k = s.act[p];
                __stcg(&s.key2[k], s.cand_dist[k] == __ldcg(&s.key1[k]) ? s.cand_tie[k] : ~0ull);
            }
            __syncthreads();
            for (int r = t; r < seg_total; r += MM_THREADS) {
                const unsigned wb = __float_as_uint(s.pw[r]);
                const int a = s.pa[r], b = s.pb[r];
                if (wb == __ldcg(&s.key1[a])) atomicMin(&s.key2[a], s.ptie[r]);
                if (wb == __ldcg(&s.key1[b])) atomicMin(&s.key2[b], s.ptie[r]);
            }
            __syncthreads();
            for (int r = t; r < seg_total; r += MM_THREADS) {
                const unsigned wb = __float_as_uint(s.pw[r]);
                const unsigned long long tb = s.ptie[r];
                const int a = s.pa[r], b = s.pb[r];
                for (int e = 0; e < 2; e++) {
                    const int x = e ? b : a;
                    if (wb == __ldcg(&s.key1[x]) && tb == __ldcg(&s.key2[x]) &&
                        !(s.cand_dist[x] == wb && s.cand_tie[x] <= tb)) {
                        s.cand_a[x] = a; s.cand_b[x] = b; s.cand_dist[x] = wb; s.cand_tie[x] = tb;
                    }
                }
            }
            __syncthreads();
        }
        p0 = p1;
    }
}

// first position of `what` in act[0..n_act)
__device__ void remove_active(const MMState &s, MMShared &sh, int n_act, int what)
{
    const int t = threadIdx.x;
    if (t == 0) sh.bi = n_act;
    __syncthreads();
    for (int p = t; p < n_act; p += MM_THREADS)
        if (s.act[p] == what) sh.bi = p;
    __syncthreads();
    const int at = sh.bi;
    // shift left by one behind `at` (chunks of MM_THREADS, front to back)
    for (int b = at; b < n_act - 1; b += MM_THREADS) {
        const int p = b + t;
        int v = 0;
        if (p < n_act - 1) v = s.act[p + 1];
        __syncthreads();
        if (p < n_act - 1) s.act[p] = v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(MM_THREADS, 1) mm_quickbuild_kernel(MMState s, int has_cf_i)
{
    __shared__ MMShared sh;
    const bool has_cf = has_cf_i != 0;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const size_t N = s.N;
    const float thr = s.thr, thr_cf = s.thr_cf;
    const float finf = __uint_as_float(FINF);

    // rng.seed(1)
    if (t == 0) {
        unsigned v = 1u;
        sh.mt[0] = v;
        for (int q = 1; q < 624; q++) { v = 1812433253u * (v ^ (v >> 30)) + (unsigned)q; sh.mt[q] = v; }
    }
    int rng_pos = 312;
    long long draws = 0;
    int n_act = s.N;
    for (int k = t; k < s.N; k += MM_THREADS) {
        s.act[k] = k; s.conv[k] = k; s.size[k] = 1.0f; s.minv_sym[k] = finf; s.flag[k] = 0;
        s.cand_dist[k] = FINF; s.cand_tie[k] = DINF;
    }
    __syncthreads();
    // Initialize: row minima (+ the prior's, which start from what the previous tree left)
    for (int k = warp; k < s.N; k += MM_WARPS) {
        float m = finf, mc = finf;
        for (int l = lane; l < s.N; l += 32)
            if (l != k) {
                m = fminf(m, s.d[k * N + l]);
                if (has_cf) mc = fminf(mc, s.cf[k * N + l]);
            }
        for (int o = 16; o; o >>= 1) { m = fminf(m, __shfl_xor_sync(~0u, m, o)); mc = fminf(mc, __shfl_xor_sync(~0u, mc, o)); }
        if (lane == 0) {
            s.minv[k] = __fadd_rn(m, thr);
            if (has_cf) s.minv_cf[k] = __fadd_rn(fminf(s.minv_cf[k], mc), thr_cf);
        }
    }
    __syncthreads();
    meet_pairs(s, sh, has_cf, true, n_act, -1, -1, 0, rng_pos, draws);
    __syncthreads();

    bool use_sym = false;
    long long first_sym = -1, sym_steps = 0;
    unsigned sbest_d = FINF; // best_sym_candidate
    int sbest_a = -1, sbest_b = -1;

    for (int node = s.N; node < 2 * s.N - 1; node++) {
        // F (of the previous step): the best candidate over the active clusters
        unsigned ba = FINF;
        unsigned long long bb = DINF;
        int bp = 0x7fffffff;
        for (int p = t; p < n_act; p += MM_THREADS) {
            const int k = s.act[p];
            const unsigned da = s.cand_dist[k];
            const unsigned long long db = s.cand_tie[k];
            if (da < ba || (da == ba && (db < bb || (db == bb && p < bp)))) { ba = da; bb = db; bp = p; }
        }
        block_lexmin(ba, bb, bp, sh);
        int ci, cj;
        if (ba == FINF) { // no mutually minimal pair: symmetric fallback (:1244-1256)
            if (!use_sym) {
                use_sym = true;
                first_sym = node - s.N;
                // InitializeSym
                for (size_t e = t; e < (size_t)n_act * n_act; e += MM_THREADS) {
                    const int x = s.act[e / n_act], y = s.act[e % n_act];
                    if (x < y) {
                        const float v = __fadd_rn(s.d[x * N + y], s.d[y * N + x]);
                        s.sym[x * N + y] = v;
                        s.sym[y * N + x] = v;
                    }
                }
                __syncthreads();
                for (int p = warp; p < n_act; p += MM_WARPS) {
                    const int x = s.act[p];
                    float m = finf;
                    int mq = 0x7fffffff;
                    for (int q = lane; q < n_act; q += 32) {
                        const int l = s.act[q];
                        if (l == x) continue;
                        const float v = s.sym[x * N + l];
                        if (v < m) { m = v; mq = q; }
                    }
                    for (int o = 16; o; o >>= 1) {
                        const float om = __shfl_xor_sync(~0u, m, o);
                        const int oq = __shfl_xor_sync(~0u, mq, o);
                        if (om < m || (om == m && oq < mq)) { m = om; mq = oq; }
                    }
                    if (lane == 0) {
                        s.minv_sym[x] = m; // min_values_sym starts at +inf for every cluster
                        s.csym_dist[x] = m;
                        if (m < finf) { s.csym_a[x] = x; s.csym_b[x] = s.act[mq]; }
                    }
                }
                __syncthreads();
                unsigned a = FINF; unsigned long long b = 0; int c = 0x7fffffff;
                for (int p = t; p < n_act; p += MM_THREADS) {
                    const unsigned v = __float_as_uint(s.csym_dist[s.act[p]]);
                    if (v < a || (v == a && p < c)) { a = v; c = p; }
                }
                block_lexmin(a, b, c, sh);
                if (a < sbest_d) { sbest_d = a; sbest_a = s.csym_a[s.act[c]]; sbest_b = s.csym_b[s.act[c]]; }
            }
            ci = sbest_a;
            cj = sbest_b;
        } else {
            const int k = s.act[bp];
            ci = s.cand_a[k];
            cj = s.cand_b[k];
        }
        if (use_sym) sym_steps++;
        __syncthreads();
        const float si = s.size[ci], sj = s.size[cj], sum = __fadd_rn(si, sj);
        if (t == 0) {
            s.merges[2 * (node - s.N)] = s.conv[ci];
            s.merges[2 * (node - s.N) + 1] = s.conv[cj];
            sh.n_rescan = 0;
        }
        __syncthreads();

        // A: rows / columns j
        float *di = s.d + ci * N, *dj = s.d + cj * N;
        float min_j = finf, min_cf = finf;
        for (int p = t; p < n_act; p += MM_THREADS) {
            const int k = s.act[p];
            if (k == ci || k == cj) continue;
            if (has_cf) {
                float *ck = s.cf + k * N;
                const float ckj = ck[cj], cki = ck[ci], cik = s.cf[ci * N + k], cjk = s.cf[cj * N + k];
                float njk = cjk;
                if (cik != cjk) { njk = mix(si, cik, sj, cjk, sum); s.cf[cj * N + k] = njk; }
                if (cki != ckj) ck[cj] = mix(si, cki, sj, ckj, sum);
                min_cf = fminf(min_cf, njk);
            }
            float *dk = s.d + k * N;
            const float dkj = dk[cj], dki = dk[ci], dik = di[k], djk = dj[k];
            float njk = djk;
            if (dik != djk) { njk = mix(si, dik, sj, djk, sum); dj[k] = njk; }
            if (dki != dkj) dk[cj] = mix(si, dki, sj, dkj, sum);
            min_j = fminf(min_j, njk);
            const float mk = s.minv[k];
            int f = 0;
            const int ca = s.cand_a[k], cb = s.cand_b[k];
            if (ca == cj || cb == cj || ca == ci || cb == ci) f = 1;
            if (dkj != dki) {
                const float base = __fsub_rn(mk, thr);
                if ((double)fabsf(__fsub_rn(base, dkj)) < 1e-4 || (double)fabsf(__fsub_rn(base, dki)) < 1e-4)
                    s.rescan[atomicAdd(&sh.n_rescan, 1)] = p;
            }
            s.flag[k] = f;
        }
        min_j = block_min(min_j, sh);
        if (has_cf) min_cf = block_min(min_cf, sh);
        if (t == 0) {
            s.minv[cj] = __fadd_rn(min_j, thr);
            if (has_cf) s.minv_cf[cj] = __fadd_rn(min_cf, thr_cf);
            s.flag[cj] = 0;
            s.cand_dist[cj] = FINF; // mcandidates[j] starts over (:541-542)
            s.cand_tie[cj] = DINF;
        }
        __syncthreads();

        // B: new row minima
        const int n_rescan = sh.n_rescan;
        for (int it = warp; it < n_rescan; it += MM_WARPS) {
            const int k = s.act[s.rescan[it]];
            const float *dk = s.d + k * N;
            const float old = __fsub_rn(s.minv[k], thr);
            float m = finf;
            int pe = 0x7fffffff, pl = 0x7fffffff;
            for (int q = lane; q < n_act; q += 32) {
                const int l = s.act[q];
                if (l == ci || l == k) continue;
                const float v = dk[l];
                m = fminf(m, v);
                if (v == old) pe = min(pe, q);
                if (v < old) pl = min(pl, q);
            }
            for (int o = 16; o; o >>= 1) {
                m = fminf(m, __shfl_xor_sync(~0u, m, o));
                pe = min(pe, __shfl_xor_sync(~0u, pe, o));
                pl = min(pl, __shfl_xor_sync(~0u, pl, o));
            }
            if (lane == 0) {
                const float mk = (pe != 0x7fffffff && pe < pl) ? old : m; // the scan stops at the old minimum (:337-339)
                s.minv[k] = __fadd_rn(mk, thr);
                s.flag[k] |= 2;
            }
        }
        __syncthreads();

        // C: U in order; its members' candidates start over
        for (int p = t; p < n_act; p += MM_THREADS) {
            const int k = s.act[p];
            const bool in = k != ci && k != cj && s.flag[k] != 0;
            s.cnt[p] = in ? 1 : 0;
            if (in) { s.flag[k] |= 4; s.cand_dist[k] = FINF; s.cand_tie[k] = DINF; }
        }
        __syncthreads();
        const int n_u = block_exscan(s.cnt, n_act, sh);
        for (int p = t; p < n_act; p += MM_THREADS) {
            const int k = s.act[p];
            if (k != ci && k != cj && (s.flag[k] & 4)) s.ulist[s.cnt[p]] = p;
        }
        __syncthreads();

        // D, E
        meet_pairs(s, sh, has_cf, false, n_act, ci, cj, n_u, rng_pos, draws);
        __syncthreads();

        // the symmetric matrix follows once it is in use (CoalesceSym)
        if (use_sym) {
            if (t == 0) sh.n_rescan = 0;
            __syncthreads();
            float *yi = s.sym + ci * N, *yj = s.sym + cj * N;
            unsigned ja = FINF; unsigned long long jb = 0; int jp = 0x7fffffff;
            for (int p = t; p < n_act; p += MM_THREADS) {
                const int k = s.act[p];
                if (k == ci || k == cj) continue;
                float *yk = s.sym + k * N;
                const float dkj = yk[cj], dki = yk[ci], dik = yi[k], djk = yj[k];
                float njk = djk;
                if (dik != djk) { njk = mix(si, dik, sj, djk, sum); yj[k] = njk; }
                if (dki != dkj) yk[cj] = mix(si, dki, sj, dkj, sum);
                const unsigned nb = __float_as_uint(njk);
                if (nb < ja || (nb == ja && p < jp)) { ja = nb; jp = p; }
                if (dkj != dki) {
                    const float mk = s.minv_sym[k];
                    if ((double)fabsf(__fsub_rn(mk, dkj)) < 1e-6 || (double)fabsf(__fsub_rn(mk, dki)) < 1e-6)
                        s.rescan[atomicAdd(&sh.n_rescan, 1)] = p;
                } else {
                    if (s.csym_a[k] == ci) s.csym_a[k] = cj;
                    if (s.csym_b[k] == ci) s.csym_b[k] = cj;
                }
            }
            block_lexmin(ja, jb, jp, sh);
            __syncthreads();
            if (t == 0) {
                s.minv_sym[cj] = __uint_as_float(ja);
                s.csym_dist[cj] = __uint_as_float(ja);
                if (ja != FINF) { s.csym_a[cj] = s.act[jp]; s.csym_b[cj] = cj; }
            }
            const int n_rs = sh.n_rescan;
            for (int it = warp; it < n_rs; it += MM_WARPS) {
                const int k = s.act[s.rescan[it]];
                const float *yk = s.sym + k * N;
                const float old = s.minv_sym[k];
                float m = finf;
                int mq = 0x7fffffff, pe = 0x7fffffff, pl = 0x7fffffff;
                for (int q = lane; q < n_act; q += 32) {
                    const int l = s.act[q];
                    if (l == ci || l == k) continue;
                    const float v = yk[l];
                    if (v < m) { m = v; mq = q; }
                    if (v == old) pe = min(pe, q);
                    if (v < old) pl = min(pl, q);
                }
                for (int o = 16; o; o >>= 1) {
                    const float om = __shfl_xor_sync(~0u, m, o);
                    const int oq = __shfl_xor_sync(~0u, mq, o);
                    if (om < m || (om == m && oq < mq)) { m = om; mq = oq; }
                    pe = min(pe, __shfl_xor_sync(~0u, pe, o));
                    pl = min(pl, __shfl_xor_sync(~0u, pl, o));
                }
                if (lane == 0) {
                    if (pe != 0x7fffffff && pe < pl) { m = old; mq = pe; }
                    s.minv_sym[k] = m;
                    s.csym_dist[k] = m;
                    if (m < finf) { s.csym_a[k] = k; s.csym_b[k] = s.act[mq]; }
                }
            }
            __syncthreads();
            // best_sym_candidate: first cluster in order with the smallest candidate, the new cluster last
            unsigned a = FINF; unsigned long long b = 0; int c = 0x7fffffff;
            for (int p = t; p < n_act; p += MM_THREADS) {
                const int k = s.act[p];
                if (k == ci) continue;
                const unsigned v = __float_as_uint(s.csym_dist[k]);
                const int ord = (k == cj) ? n_act : p;
                if (v < a || (v == a && ord < c)) { a = v; c = ord; }
            }
            block_lexmin(a, b, c, sh);
            sbest_d = a;
            if (a != FINF) {
                const int k = (c == n_act) ? cj : s.act[c];
                sbest_a = s.csym_a[k];
                sbest_b = s.csym_b[k];
            }
            __syncthreads();
        }

        // bookkeeping
        if (t == 0) {
            s.size[cj] = sum;
            s.conv[cj] = node;
        }
        remove_active(s, sh, n_act, ci);
        n_act--;
        __syncthreads();
    }
    if (t == 0) {
        s.info[0] = draws;
        s.info[1] = first_sym;
        s.info[2] = sym_steps;
    }
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------------
struct rp_minmatch {
    int device = 0, N = 0;
    MMState s{};
    void *block = nullptr; // one allocation behind every array
    int *h_merges = nullptr; // pinned
    long long *h_info = nullptr;
    cudaStream_t stream = nullptr;
    float last_ms = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

#define MM_CUDA(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return rp::api_fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RP_ENODEVICE  \
                                                                                              : RP_ECUDA,     \
                                std::string(#call) + ": " + cudaGetErrorString(e_));                         \
    } while (0)

extern "C" int rp_minmatch_create(int device, int N, double theta, rp_minmatch **out)
{
    if (!(theta > 0 && theta < 1)) return rp::api_fail(RP_EINVAL, "rp_minmatch_create: bad argument");
    return rp_minmatch_create_thresholds(device, N, (float)(-0.2 * std::log(theta / (1.0 - theta))),    // tree_builder.cpp:43
                                         (float)(-0.001 * std::log(theta / (1.0 - theta))), out);        // :44
}

extern "C" int rp_minmatch_create_thresholds(int device, int N, float threshold, float threshold_cf, rp_minmatch **out)
{
    if (!out || N < 2) return rp::api_fail(RP_EINVAL, "rp_minmatch_create: bad argument");
    int ndev = 0;
    MM_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return rp::api_fail(RP_ENODEVICE, "rp_minmatch_create: no such device");
    MM_CUDA(cudaSetDevice(device));
    auto *h = new rp_minmatch;
    h->device = device;
    h->N = N;
    MMState &s = h->s;
    s.N = N;
    s.thr = threshold;
    s.thr_cf = threshold_cf;
    const size_t nn = (size_t)N * N;
    s.cap = (int)std::max<size_t>((size_t)N + 1, std::min<size_t>(nn / 2 + 1, (size_t)1 << 22));
    if (const char *e = getenv("RP_MINMATCH_CAP")) s.cap = std::max(N + 1, atoi(e));
    // layout of the single block (8-byte items first)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_tie = take(8 * (size_t)N), o_key2 = take(8 * (size_t)N), o_ptie = take(8 * (size_t)s.cap), o_info = take(64);
    const size_t o_d = take(4 * nn), o_cf = take(4 * nn), o_sym = take(4 * nn);
    const size_t o_f[5] = {take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N)};
    const size_t o_i[12] = {take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N),
                            take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N),
                            take(4 * (size_t)N), take(4 * (size_t)(N + 2)), take(4 * (size_t)N), take(4 * (size_t)N)};
    const size_t o_pa = take(4 * (size_t)s.cap), o_pb = take(4 * (size_t)s.cap), o_pw = take(4 * (size_t)s.cap);
    const size_t o_merges = take(8 * (size_t)N);
    cudaError_t e = cudaMalloc(&h->block, off);
    if (e != cudaSuccess) {
        delete h;
        return rp::api_fail(RP_ENOMEM, std::string("rp_minmatch_create: cudaMalloc: ") + cudaGetErrorString(e));
    }
    char *b = (char *)h->block;
    s.cand_tie = (unsigned long long *)(b + o_tie);
    s.key2 = (unsigned long long *)(b + o_key2);
    s.ptie = (unsigned long long *)(b + o_ptie);
    s.info = (long long *)(b + o_info);
    s.d = (float *)(b + o_d); s.cf = (float *)(b + o_cf); s.sym = (float *)(b + o_sym);
    s.minv_cf = (float *)(b + o_f[0]); s.csym_dist = (float *)(b + o_f[1]); s.minv = (float *)(b + o_f[2]);
    s.minv_sym = (float *)(b + o_f[3]); s.size = (float *)(b + o_f[4]);
    s.cand_a = (int *)(b + o_i[0]); s.cand_b = (int *)(b + o_i[1]); s.cand_dist = (unsigned *)(b + o_i[2]);
    s.csym_a = (int *)(b + o_i[3]); s.csym_b = (int *)(b + o_i[4]); s.conv = (int *)(b + o_i[5]); s.act = (int *)(b + o_i[6]);
    s.flag = (int *)(b + o_i[7]); s.rescan = (int *)(b + o_i[8]); s.cnt = (int *)(b + o_i[9]); s.ulist = (int *)(b + o_i[10]);
    s.key1 = (unsigned *)(b + o_i[11]);
    s.pa = (int *)(b + o_pa); s.pb = (int *)(b + o_pb); s.pw = (float *)(b + o_pw);
    s.merges = (int *)(b + o_merges);
    // a fresh MinMatch object: min_values_CF = 0 (vector::resize), candidates name nobody (lin1 = lin2 = -1)
    MM_CUDA(cudaMemset(s.minv_cf, 0, 4 * (size_t)N));
    MM_CUDA(cudaMemset(s.cand_a, 0xff, 4 * (size_t)N));
    MM_CUDA(cudaMemset(s.cand_b, 0xff, 4 * (size_t)N));
    MM_CUDA(cudaMemset(s.csym_a, 0xff, 4 * (size_t)N));
    MM_CUDA(cudaMemset(s.csym_b, 0xff, 4 * (size_t)N));
    MM_CUDA(cudaMemset(s.csym_dist, 0x7f, 4 * (size_t)N));
    MM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    MM_CUDA(cudaEventCreate(&h->e0));
    MM_CUDA(cudaEventCreate(&h->e1));
    MM_CUDA(cudaMallocHost(&h->h_merges, 8 * (size_t)N));
    MM_CUDA(cudaMallocHost(&h->h_info, 64));
    *out = h;
    return RP_OK;
}

extern "C" void rp_minmatch_destroy(rp_minmatch *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    if (h->h_merges) cudaFreeHost(h->h_merges);
    if (h->h_info) cudaFreeHost(h->h_info);
    if (h->block) cudaFree(h->block);
    delete h;
}

static int mm_run(rp_minmatch *h, bool has_prior, int *merges, rp_minmatch_stats *st)
{
    MM_CUDA(cudaEventRecord(h->e0, h->stream));
    mm_quickbuild_kernel<<<1, MM_THREADS, 0, h->stream>>>(h->s, has_prior ? 1 : 0);
    MM_CUDA(cudaGetLastError());
    MM_CUDA(cudaEventRecord(h->e1, h->stream));
    MM_CUDA(cudaMemcpyAsync(h->h_merges, h->s.merges, 8 * (size_t)(h->N - 1), cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaMemcpyAsync(h->h_info, h->s.info, 24, cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(merges, h->h_merges, 8 * (size_t)(h->N - 1));
    MM_CUDA(cudaEventElapsedTime(&h->last_ms, h->e0, h->e1));
    if (st) {
        st->ms_kernel = h->last_ms;
        st->draws = h->h_info[0];
        st->first_fallback_step = (int)h->h_info[1];
        st->fallback_steps = (int)h->h_info[2];
        st->launches = 1;
    }
    return RP_OK;
}

extern "C" int rp_minmatch_quickbuild(rp_minmatch *h, const float *d, const float *d_prior, int *merges, rp_minmatch_stats *st)
{
    if (!h || !d || !merges) return rp::api_fail(RP_EINVAL, "rp_minmatch_quickbuild: null argument");
    MM_CUDA(cudaSetDevice(h->device));
    const size_t bytes = 4 * (size_t)h->N * h->N;
    MM_CUDA(cudaMemcpyAsync(h->s.d, d, bytes, cudaMemcpyHostToDevice, h->stream));
    if (d_prior) MM_CUDA(cudaMemcpyAsync(h->s.cf, d_prior, bytes, cudaMemcpyHostToDevice, h->stream));
    return mm_run(h, d_prior != nullptr, merges, st);
}

extern "C" int rp_minmatch_quickbuild_device(rp_minmatch *h, const float *dev_d, const float *dev_prior, int *merges,
                                             rp_minmatch_stats *st)
{
    if (!h || !dev_d || !merges) return rp::api_fail(RP_EINVAL, "rp_minmatch_quickbuild_device: null argument");
    MM_CUDA(cudaSetDevice(h->device));
    const size_t bytes = 4 * (size_t)h->N * h->N;
    MM_CUDA(cudaMemcpyAsync(h->s.d, dev_d, bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (dev_prior) MM_CUDA(cudaMemcpyAsync(h->s.cf, dev_prior, bytes, cudaMemcpyDeviceToDevice, h->stream));
    return mm_run(h, dev_prior != nullptr, merges, st);
}
