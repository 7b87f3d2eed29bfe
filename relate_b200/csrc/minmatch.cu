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
//   B  the new minimum of each such row (up to eight rows in one block-wide pass, else a warp per row), with the reference's
//      early `break` decided from three reductions (smallest value, first position equal to the old minimum, first position
//      below it)
//   C  U = rows whose minimum moved or whose candidate named i or j (the reference's `updated_cluster`), in order
//   D  find the feasible pairs — rows of U against every earlier row, the other rows against the earlier members of U, and row j
//      against everybody, which is exactly the set and the order the reference meets them in — and rank them in that order:
//      a pair's rank is the index of its random draw.  Three sizes: a handful of pairs are ranked by counting in shared memory;
//      hundreds (blocks of identical haplotypes) keep one bit mask per position and rank through ballots and per-warp counts;
//      anything (Initialize, more than 63 members of U) is counted row by row, prefix-summed and cut into segments
//   E  give pair r the r-th next output of a device-side std::mt19937 (seeded with 1 per tree, two 32-bit outputs per double as
//      libstdc++'s generate_canonical does), and let both members keep the lexicographically smallest (weight, draw) among their
//      old candidate and their new pairs (in shared memory, or three rounds of atomicMin over the pair buffer)
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

constexpr int MM_MAX_WARPS = 32;
constexpr unsigned KINF = 0xff800000u;                       // fkey(+inf)
constexpr unsigned long long DINF = 0x7ff0000000000000ull;   // double +inf
constexpr int RS_MAX = 8;    // rows that look for a new minimum: up to this many block-wide in one pass
constexpr int U_MAX = 8;     // members of U: up to this many through the block-wide pair test
constexpr int PAIR_MAX = 128; // pairs of a step: up to this many ranked / applied in shared memory
constexpr int IMAX = 0x7fffffff;

struct MMState {
    int N;
    float thr, thr_cf;
    int force_general; // tests: skip the small-case paths
    int use_smem;      // the per-cluster arrays fit into shared memory (40 bytes per cluster)
    int pre_init;      // Initialize's O(N^2) passes were done by the grid-wide kernels in front (minima in minv / minv_cf, pairs in the buffer)
    int *init_cnt;     // their per-row counts / offsets (N + 1) and, behind them, the total
    // matrices (row-major N x N)
    float *d, *cf, *sym;
    float *dT, *cfT; // transposes, kept in step: a column of d is a row of dT (a warp reading a column touches 32 lines)
    // persistent between trees
    float *minv_cf;
    int *cand_a, *cand_b;
    unsigned *cand_dist;           // fkey(dist): order-preserving key of the float
    unsigned long long *cand_tie;  // double bits (ties are in [0,1) or +inf: bit order == value order)
    int *csym_a, *csym_b;
    float *csym_dist;
    // per tree
    float *minv, *minv_sym, *size;
    int *conv, *act;
    int *flag;     // bit 0: candidate names i or j, bit 1: minimum recomputed, bit 2: member of U
    int *rescan;   // positions whose row needs a new minimum
    int *ulist;    // positions of U, ascending
    int *cnt;      // pairs per row (n_act + 1 entries), then exclusive offsets
    int *wcnt;     // medium path: pairs per U row and warp of positions, 64 x (N / 32 + 1)
    unsigned *key1;
    unsigned long long *key2;
    // pair buffer of the general path
    int cap;
    int *pa, *pb;
    unsigned *pw;
    unsigned long long *ptie;
    // output
    int *merges;
    long long *info; // [0] draws, [1] first step without a candidate (-1), [2] steps on the fallback, [3] steps on the general path,
                     // [12] steps on the medium path ([4..11]: per-phase cycles with -DMM_PROF)
};

struct MMShared {
    unsigned mt[624];
    unsigned redu[MM_MAX_WARPS], redu2[MM_MAX_WARPS];
    unsigned long long redull[MM_MAX_WARPS];
    int redi[MM_MAX_WARPS];
    int scan[MM_MAX_WARPS];
    int n_rescan, total, p_end, pos_i;
    // rows that look for a new minimum
    int rs_pos[RS_MAX], rs_cl[RS_MAX], rs_pe[RS_MAX], rs_pl[RS_MAX];
    unsigned rs_m[RS_MAX];
    float rs_old[RS_MAX];
    // U
    int u_pos[U_MAX], u_cl[U_MAX];
    float u_minv[U_MAX];
    // pairs of a small step
    unsigned long long pk[PAIR_MAX], pt[PAIR_MAX];
    int pa[PAIR_MAX], pb[PAIR_MAX], prank[PAIR_MAX];
    unsigned pw[PAIR_MAX];
};

// order-preserving key of a float (-0 counts as +0): a < b  <=>  fkey(a) < fkey(b)
__device__ __forceinline__ unsigned fkey(float v)
{
    const unsigned b = __float_as_uint(v + 0.0f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float funkey(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

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

// libstdc++ generate_canonical<double, 53>(mt19937): (g1 + g2 * 2^32) / 2^64, as the bits of the double
__device__ __forceinline__ unsigned long long canonical(unsigned g1, unsigned g2)
{
    const double u = __dmul_rn(__dadd_rn(__uint2double_rn(g1), __dmul_rn(__uint2double_rn(g2), 4294967296.0)),
                               5.42101086242752217003726400434970855712890625e-20);
    return u >= 1.0 ? 0x3fefffffffffffffull : (unsigned long long)__double_as_longlong(u);
}

// the next 624 words of std::mt19937 (three dependent thirds)
template <int TH>
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

// two minima at once; every thread gets both (REDUX on order-preserving keys)
template <int TH>
__device__ __forceinline__ void block_min2(float &a, float &b, MMShared &sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned ka = __reduce_min_sync(~0u, fkey(a)), kb = __reduce_min_sync(~0u, fkey(b));
    __syncthreads();
    if (lane == 0) { sh.redu[warp] = ka; sh.redu2[warp] = kb; }
    __syncthreads();
    ka = __reduce_min_sync(~0u, lane < TH / 32 ? sh.redu[lane] : ~0u);
    kb = __reduce_min_sync(~0u, lane < TH / 32 ? sh.redu2[lane] : ~0u);
    a = funkey(ka);
    b = funkey(kb);
}

// lexicographic minimum of (a, b, c) over a warp, four REDUX
__device__ __forceinline__ void warp_lexmin(unsigned &a, unsigned long long &b, int &c)
{
    const unsigned m = __reduce_min_sync(~0u, a);
    bool e = a == m;
    const unsigned hi = __reduce_min_sync(~0u, e ? (unsigned)(b >> 32) : ~0u);
    e = e && (unsigned)(b >> 32) == hi;
    const unsigned lo = __reduce_min_sync(~0u, e ? (unsigned)b : ~0u);
    e = e && (unsigned)b == lo;
    c = __reduce_min_sync(~0u, e ? c : IMAX);
    a = m;
    b = ((unsigned long long)hi << 32) | lo;
}

// lexicographic minimum of (a, b, c) over the block; every thread gets the winner
template <int TH>
__device__ __forceinline__ void block_lexmin(unsigned &a, unsigned long long &b, int &c, MMShared &sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_lexmin(a, b, c);
    __syncthreads();
    if (lane == 0) { sh.redu[warp] = a; sh.redull[warp] = b; sh.redi[warp] = c; }
    __syncthreads();
    const bool in = lane < TH / 32;
    a = in ? sh.redu[lane] : ~0u;
    b = in ? sh.redull[lane] : ~0ull;
    c = in ? sh.redi[lane] : IMAX;
    warp_lexmin(a, b, c);
}

// exclusive prefix sum of arr[0..n) in place; returns the total to every thread
template <int TH>
__device__ int block_exscan(int *arr, int n, MMShared &sh)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, per = (n + TH - 1) / TH;
    const int b = min(t * per, n), e = min(b + per, n);
    int s = 0;
    for (int q = b; q < e; q++) s += arr[q];
    int inc = s;
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(~0u, inc, o);
        if (lane >= o) inc += u;
    }
    __syncthreads();
    if (lane == 31) sh.scan[warp] = inc;
    __syncthreads();
    const int wc = lane < TH / 32 ? sh.scan[lane] : 0;
    const int wbase = __reduce_add_sync(~0u, lane < warp ? wc : 0), total = __reduce_add_sync(~0u, wc);
    int run = wbase + inc - s;
    for (int q = b; q < e; q++) { int v = arr[q]; arr[q] = run; run += v; }
    __syncthreads();
    return total;
}

// weight of a feasible pair (x, y): with a prior matrix 0 if the pair is mutually minimal there too, else d + d^T.  Symmetric in
// (x, y); everything is read from row x of a matrix or of its transpose, so callers pass the cluster their warp shares as x.
__device__ __forceinline__ float pair_weight(const MMState &s, bool has_cf, int x, int y)
{
    const size_t o = (size_t)x * s.N + y;
    const float dxy = s.d[o], dyx = s.dT[o]; // (all four loads before the first use: one round trip, not three)
    if (has_cf) {
        const float cxy = s.cf[o], cyx = s.cfT[o];
        if (cxy <= s.minv_cf[x] && cyx <= s.minv_cf[y]) return 0.0f;
    }
    return __fadd_rn(dxy, dyx);
}

enum { ROW_SKIP = 0, ROW_BELOW = 1, ROW_LIST = 2, ROW_ABOVE = 3, ROW_J = 4 };

// General path.  Feasible pairs of row p in the reference's order.  EMIT = false: count them; EMIT = true: write them at
// `base`.  Warp-wide (all lanes call it), four 32-element pieces per round so that their loads overlap.
template <bool EMIT>
__device__ int row_pairs_warp(const MMState &s, bool has_cf, int n_act, int p, int kind, int ci, int cj, int base)
{
    const size_t N = s.N;
    const int lane = threadIdx.x & 31;
    const bool jrow = kind == ROW_J;
    const int x = jrow ? cj : s.act[p];
    const float mx = s.minv[x];
    const float *dx = s.d + x * N;
    const int q0 = kind == ROW_ABOVE ? p + 1 : 0, q1 = (kind == ROW_BELOW) ? p : n_act;
    int count = 0;
    for (int qb = q0; qb < q1; qb += 128) {
        int y[4];
        float v[4];
        bool ok[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int q = qb + 32 * c + lane;
            y[c] = q < q1 ? s.act[q] : -1;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            ok[c] = y[c] >= 0 && y[c] != ci && y[c] != cj;
            v[c] = ok[c] ? dx[y[c]] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) ok[c] = ok[c] && v[c] <= mx && s.dT[x * N + y[c]] <= s.minv[y[c]];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const unsigned m = __ballot_sync(~0u, ok[c]);
            if (EMIT && ok[c]) {
                const int r = base + count + __popc(m & ((1u << lane) - 1));
                // candidate orientation: (row, partner), except for the new cluster's row: (partner, j)   (:565-573)
                s.pa[r] = jrow ? y[c] : x;
                s.pb[r] = jrow ? x : y[c];
                s.pw[r] = fkey(pair_weight(s, has_cf, x, y[c]));
            }
            count += __popc(m);
        }
    }
    return count;
}

// a row outside U meets the earlier members of U only.  One thread per row: d[x][y] is read as dT[y][x] and d[y][x] as such, so
// that the threads of a warp (neighbouring x, the same y) share cache lines
template <bool EMIT>
__device__ int row_pairs_list(const MMState &s, bool has_cf, int p, int n_u, int base)
{
    const size_t N = s.N;
    const int x = s.act[p];
    const float mx = s.minv[x];
    int count = 0;
    for (int u0 = 0; u0 < n_u; u0 += 4) {
        int y[4];
        float a[4], b[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int pu = u0 + e < n_u ? s.ulist[u0 + e] : 0x7fffffff;
            y[e] = pu < p ? s.act[pu] : -1;
        }
        if (y[0] < 0) break;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            a[e] = y[e] >= 0 ? s.dT[y[e] * N + x] : 0.f; // d[x][y]
            b[e] = y[e] >= 0 ? s.d[y[e] * N + x] : 0.f;  // d[y][x]
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (y[e] >= 0 && a[e] <= mx && b[e] <= s.minv[y[e]]) {
                if (EMIT) {
                    s.pa[base + count] = x;
                    s.pb[base + count] = y[e];
                    s.pw[base + count] = fkey(pair_weight(s, has_cf, y[e], x));
                }
                count++;
            }
    }
    return count;
}

__device__ __forceinline__ int row_kind(const MMState &s, bool init, int n_act, int p, int ci, int cj)
{
    if (init) return ROW_ABOVE;
    if (p == n_act) return ROW_J;
    const int k = s.act[p];
    return (k == ci || k == cj) ? ROW_SKIP : ((s.flag[k] & 4) ? ROW_BELOW : ROW_LIST);
}

// the next `total` draws of the tree's mt19937 as ptie[0..total)
template <int TH>
__device__ void draw_ties(const MMState &s, MMShared &sh, int total, int &rng_pos, long long &draws)
{
    const int t = threadIdx.x;
    int cur = 0;
    while (cur < total) {
        if (rng_pos == 312) { mt_twist<TH>(sh); rng_pos = 0; }
        const int n = min(312 - rng_pos, total - cur);
        for (int r = t; r < n; r += TH)
            s.ptie[cur + r] = canonical(temper(sh.mt[2 * (rng_pos + r)]), temper(sh.mt[2 * (rng_pos + r) + 1]));
        rng_pos += n;
        cur += n;
        __syncthreads();
    }
    draws += total;
}

// Pairs 0..total-1 of the pair buffer, in the order the reference meets them: pair r takes the r-th next draw; both members keep
// the lexicographically smallest (weight, draw) among their candidate and the new pairs (three rounds of atomicMin on keys
// that are initialised from the members' current candidates).
template <int TH>
__device__ void draw_and_apply(const MMState &s, MMShared &sh, int total, int &rng_pos, long long &draws)
{
    const int t = threadIdx.x;
    draw_ties<TH>(s, sh, total, rng_pos, draws);
    for (int r = t; r < total; r += TH) {
        const int a = s.pa[r], b = s.pb[r];
        __stcg(&s.key1[a], s.cand_dist[a]);
        __stcg(&s.key1[b], s.cand_dist[b]);
    }
    __syncthreads();
    // (pairs of one row are neighbours in the buffer and share a member: one atomic per warp and member, not one per pair)
    for (int r0 = 0; r0 < total; r0 += TH) {
        const int r = r0 + t;
        const bool in = r < total;
        const unsigned wb = in ? s.pw[r] : ~0u;
        for (int e = 0; e < 2; e++) {
            const int x = in ? (e ? s.pb[r] : s.pa[r]) : -1;
            const unsigned grp = __match_any_sync(~0u, x);
            const unsigned m = __reduce_min_sync(grp, wb);
            if (in && (threadIdx.x & 31) == __ffs(grp) - 1) atomicMin(&s.key1[x], m);
        }
    }
    __syncthreads();
    for (int r = t; r < total; r += TH) {
        const int a = s.pa[r], b = s.pb[r];
        __stcg(&s.key2[a], s.cand_dist[a] == __ldcg(&s.key1[a]) ? s.cand_tie[a] : ~0ull);
        __stcg(&s.key2[b], s.cand_dist[b] == __ldcg(&s.key1[b]) ? s.cand_tie[b] : ~0ull);
    }
    __syncthreads();
    for (int r0 = 0; r0 < total; r0 += TH) {
        const int r = r0 + t;
        const bool in = r < total;
        const unsigned wb = in ? s.pw[r] : ~0u;
        const unsigned long long tb = in ? s.ptie[r] : ~0ull;
        for (int e = 0; e < 2; e++) {
            const int x = in ? (e ? s.pb[r] : s.pa[r]) : -1;
            const bool cand = in && wb == __ldcg(&s.key1[x]);
            const unsigned grp = __match_any_sync(~0u, x);
            const unsigned hi = __reduce_min_sync(grp, cand ? (unsigned)(tb >> 32) : ~0u);
            const unsigned lo = __reduce_min_sync(grp, (cand && (unsigned)(tb >> 32) == hi) ? (unsigned)tb : ~0u);
            if (in && (threadIdx.x & 31) == __ffs(grp) - 1 && !(hi == ~0u && lo == ~0u)) atomicMin(&s.key2[x], ((unsigned long long)hi << 32) | lo);
        }
    }
    __syncthreads();
    for (int r = t; r < total; r += TH) {
        const unsigned wb = s.pw[r];
        const unsigned long long tb = s.ptie[r];
        const int a = s.pa[r], b = s.pb[r];
        for (int e = 0; e < 2; e++) {
            const int x = e ? b : a;
            if (wb == __ldcg(&s.key1[x]) && tb == __ldcg(&s.key2[x]) && !(s.cand_dist[x] == wb && s.cand_tie[x] <= tb)) {
                s.cand_a[x] = a; s.cand_b[x] = b; s.cand_dist[x] = wb; s.cand_tie[x] = tb;
            }
        }
    }
    __syncthreads();
}

// General path of phases D + E (init: every row looks above itself): count, rank, draw, apply.  Any number of pairs.
template <int TH>
__device__ void meet_pairs(const MMState &s, MMShared &sh, bool has_cf, bool init, int n_act, int ci, int cj, int n_u,
                           int &rng_pos, long long &draws)
{
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int rows = init ? n_act : n_act + 1; // + the new cluster's row
    for (int p = warp; p < rows; p += (TH / 32)) {
        const int kind = row_kind(s, init, n_act, p, ci, cj);
        if (kind == ROW_SKIP) { if (lane == 0) s.cnt[p] = 0; continue; }
        if (kind == ROW_LIST) continue; // thread-level below
        const int c = row_pairs_warp<false>(s, has_cf, n_act, p, kind, ci, cj, 0);
        if (lane == 0) s.cnt[p] = c;
    }
    if (!init)
        for (int p = t; p < n_act; p += TH) {
            const int k = s.act[p];
            if (k == ci || k == cj || (s.flag[k] & 4)) continue;
            s.cnt[p] = n_u ? row_pairs_list<false>(s, has_cf, p, n_u, 0) : 0;
        }
    __syncthreads();
    const int total = block_exscan<TH>(s.cnt, rows, sh);
    if (total == 0) return;
    // segments of at most cap pairs (one segment unless a tie-heavy matrix meets a small buffer)
    int p0 = 0;
    while (p0 < rows) {
        if (t == 0) sh.p_end = rows;
        __syncthreads();
        const int off0 = s.cnt[p0];
        for (int p = p0 + t; p < rows; p += TH) {
            const int endp = (p + 1 < rows) ? s.cnt[p + 1] : total;
            if (endp - off0 > s.cap && s.cnt[p] - off0 <= s.cap) sh.p_end = max(p, p0 + 1);
        }
        __syncthreads();
        const int p1 = sh.p_end;
        const int seg_total = ((p1 < rows) ? s.cnt[p1] : total) - off0;
        __syncthreads();
        if (seg_total > 0) {
            for (int p = p0 + warp; p < p1; p += (TH / 32)) {
                const int kind = row_kind(s, init, n_act, p, ci, cj);
                if (kind == ROW_SKIP || kind == ROW_LIST) continue;
                row_pairs_warp<true>(s, has_cf, n_act, p, kind, ci, cj, s.cnt[p] - off0);
            }
            if (!init && n_u)
                for (int p = p0 + t; p < min(p1, n_act); p += TH) {
                    const int k = s.act[p];
                    if (k == ci || k == cj || (s.flag[k] & 4)) continue;
                    row_pairs_list<true>(s, has_cf, p, n_u, s.cnt[p] - off0);
                }
            draw_and_apply<TH>(s, sh, seg_total, rng_pos, draws);
        }
        p0 = p1;
    }
}

// Small-case path of phases D + E: at most U_MAX members of U and at most PAIR_MAX pairs.  Every pair with a member of U is
// tested by the thread of the OTHER cluster's position (both in U: the later one), row j by every thread; the pairs found
// are ranked by (row position, partner position) — the order the reference meets them in — in shared memory.
// Returns false (nothing changed) if there are more than PAIR_MAX pairs.
template <int TH, int NU>
__device__ bool meet_pairs_small(const MMState &s, MMShared &sh, bool has_cf, int n_act, int ci, int cj, int n_u, int &rng_pos,
                                 long long &draws)
{
    const size_t N = s.N;
    const int t = threadIdx.x;
    const float *dj = s.d + cj * N, *dTj = s.dT + cj * N;
    const float mj = s.minv[cj];
    // E positions per thread and round: their loads are issued together (a CTA of a few warps cannot hide an L2 round trip
    // per position otherwise)
    constexpr int E = TH > 512 ? 1 : (TH > 256 ? (NU <= 2 ? 2 : 1) : (NU <= 2 ? 4 : (NU <= 4 ? 2 : 1))), NA = NU ? NU : 1;
    for (int q0 = t; q0 < n_act; q0 += TH * E) {
        int k[E], onm[E];
        float mk[E], a[E][NA], b[E][NA], ja[E], jb[E];
#pragma unroll
        for (int e = 0; e < E; e++) {
            const int q = q0 + e * TH;
            k[e] = q < n_act ? s.act[q] : -1;
            if (k[e] == ci || k[e] == cj) k[e] = -1;
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
            onm[e] = 0;
            if (k[e] < 0) continue;
            const int q = q0 + e * TH;
            const bool in_u = (s.flag[k[e]] & 4) != 0;
            mk[e] = s.minv[k[e]];
#pragma unroll
            for (int ui = 0; ui < NU; ui++) {
                const bool on = ui < n_u && q != sh.u_pos[ui] && !(in_u && q < sh.u_pos[ui]);
                onm[e] |= on ? (1 << ui) : 0;
                a[e][ui] = on ? s.d[sh.u_cl[ui] * N + k[e]] : 0.f;
                b[e][ui] = on ? s.dT[sh.u_cl[ui] * N + k[e]] : 0.f;
            }
            ja[e] = dj[k[e]];
            jb[e] = dTj[k[e]];
        }
#pragma unroll
        for (int e = 0; e < E; e++) {
            if (k[e] < 0) continue;
            const int q = q0 + e * TH;
#pragma unroll
            for (int ui = 0; ui < NU; ui++)
                if (((onm[e] >> ui) & 1) && b[e][ui] <= mk[e] && a[e][ui] <= sh.u_minv[ui]) {
                    const int idx = atomicAdd(&sh.total, 1);
                    if (idx < PAIR_MAX) {
                        const int pu = sh.u_pos[ui], u = sh.u_cl[ui];
                        const bool k_is_row = q > pu;
                        sh.pk[idx] = ((unsigned long long)max(q, pu) << 32) | (unsigned)min(q, pu);
                        sh.pa[idx] = k_is_row ? k[e] : u;
                        sh.pb[idx] = k_is_row ? u : k[e];
                        sh.pw[idx] = fkey(pair_weight(s, has_cf, u, k[e]));
                    }
                }
            if (ja[e] <= mj && jb[e] <= mk[e]) {
                const int idx = atomicAdd(&sh.total, 1);
                if (idx < PAIR_MAX) {
                    sh.pk[idx] = ((unsigned long long)n_act << 32) | (unsigned)q;
                    sh.pa[idx] = k[e]; // (:565-573)
                    sh.pb[idx] = cj;
                    sh.pw[idx] = fkey(pair_weight(s, has_cf, cj, k[e]));
                }
            }
        }
    }
    __syncthreads();
    const int total = sh.total;
    if (total == 0) return true;
    if (total > PAIR_MAX) return false;
    if (t < total) {
        int rank = 0;
        const unsigned long long mine = sh.pk[t];
        for (int r = 0; r < total; r++) rank += sh.pk[r] < mine;
        sh.prank[t] = rank;
    }
    int base = 0;
    while (base < total) {
        if (rng_pos == 312) { mt_twist<TH>(sh); rng_pos = 0; }
        const int n = min(312 - rng_pos, total - base);
        if (t < total) {
            const int r = sh.prank[t] - base;
            if (r >= 0 && r < n) sh.pt[t] = canonical(temper(sh.mt[2 * (rng_pos + r)]), temper(sh.mt[2 * (rng_pos + r) + 1]));
        }
        rng_pos += n;
        base += n;
        __syncthreads();
    }
    draws += total;
    bool win[2] = {false, false};
    if (t < total) {
        for (int e = 0; e < 2; e++) {
            const int x = e ? sh.pb[t] : sh.pa[t];
            unsigned bw = s.cand_dist[x];
            unsigned long long bt = s.cand_tie[x];
            int w = -1;
            for (int r = 0; r < total; r++)
                if ((sh.pa[r] == x || sh.pb[r] == x) && (sh.pw[r] < bw || (sh.pw[r] == bw && sh.pt[r] < bt))) { bw = sh.pw[r]; bt = sh.pt[r]; w = r; }
            win[e] = w == t;
        }
    }
    __syncthreads();
    if (t < total)
        for (int e = 0; e < 2; e++)
            if (win[e]) {
                const int x = e ? sh.pb[t] : sh.pa[t];
                s.cand_a[x] = sh.pa[t]; s.cand_b[x] = sh.pb[t]; s.cand_dist[x] = sh.pw[t]; s.cand_tie[x] = sh.pt[t];
            }
    return true;
}

// Medium path of phases D + E: up to 63 members of U, any number of pairs that fits the pair buffer, up to MED_E positions per
// thread.  Thread q tests position q against every member of U it is responsible for (the pair {q, u} belongs to q if q comes
// before u, or if q is outside U) and against row j, once, and keeps the outcomes as a bit mask; ballots and per-warp counts
// turn the masks into the pairs' ranks in the reference's order (row position, then partner position).  Tie-rich data (many
// identical haplotypes) has hundreds of feasible pairs in most steps — they pass through here.
template <int TH>
__device__ bool meet_pairs_medium(const MMState &s, MMShared &sh, bool has_cf, int n_act, int ci, int cj, int n_u, int &rng_pos,
                                  long long &draws)
{
    constexpr int MED_E = TH <= 256 ? 4 : (TH <= 512 ? 20 : 10); // positions per thread: 256 threads serve N < 512, the others N <= 10 240
    if (n_u > 63 || n_act > TH * MED_E) return false;
    const size_t N = s.N;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int npw = (n_act + 31) >> 5; // warps of positions
    int *wcnt = s.wcnt;                // [n_u + 1][npw]: pairs of U row ui (row j: index n_u) found in that warp of positions
    const float mj = s.minv[cj];
    const float *dj = s.d + cj * N, *dTj = s.dT + cj * N;
    unsigned long long bits[MED_E];
#pragma unroll
    for (int e = 0; e < MED_E; e++) {
        bits[e] = 0;
        if (e * TH >= n_act) break;
        const int q = e * TH + t;
        const int k = q < n_act ? s.act[q] : -1;
        const bool live = k >= 0 && k != ci && k != cj;
        unsigned long long b = 0;
        int nb = 0; // members of U before position q
        if (live) {
            const bool in_u = (s.flag[k] & 4) != 0;
            const float mk = s.minv[k];
            for (int u0 = 0; u0 < n_u; u0 += 4) {
                float a[4], c[4], mu[4];
                bool on[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int ui = u0 + j;
                    const int pu = ui < n_u ? s.ulist[ui] : -1;
                    on[j] = pu >= 0 && pu != q && (q < pu || !in_u);
                    nb += pu >= 0 && pu < q;
                    const int u = on[j] ? s.act[pu] : 0;
                    a[j] = on[j] ? s.d[u * N + k] : 0.f;  // d[u][k]
                    c[j] = on[j] ? s.dT[u * N + k] : 0.f; // d[k][u]
                    mu[j] = on[j] ? s.minv[u] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (on[j] && c[j] <= mk && a[j] <= mu[j]) b |= 1ull << (u0 + j);
            }
            if (dj[k] <= mj && dTj[k] <= mk) b |= 1ull << 63;
            // the pairs of this position's own row (outside U: the members of U before it)
            s.cnt[q] = in_u ? 0 : __popcll(b & ((1ull << nb) - 1));
        } else if (q < n_act)
            s.cnt[q] = 0;
        bits[e] = b;
        const int pw = q >> 5;
        for (int ui = 0; ui < n_u; ui++) {
            const unsigned m = __ballot_sync(~0u, ((b >> ui) & 1) && q < s.ulist[ui]);
            if (lane == 0 && pw < npw) wcnt[ui * npw + pw] = __popc(m);
        }
        const unsigned m = __ballot_sync(~0u, (b >> 63) != 0);
        if (lane == 0 && pw < npw) wcnt[n_u * npw + pw] = __popc(m);
    }
    __syncthreads();
    // per U row (and row j): exclusive prefix over the warps of positions, total into the row's count
    for (int ui = warp; ui <= n_u; ui += TH / 32) {
        int run = 0;
        for (int w0 = 0; w0 < npw; w0 += 32) {
            const int w = w0 + lane;
            const int v = w < npw ? wcnt[ui * npw + w] : 0;
            int inc = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(~0u, inc, o);
                if (lane >= o) inc += up;
            }
            if (w < npw) wcnt[ui * npw + w] = run + inc - v;
            run += __shfl_sync(~0u, inc, 31);
        }
        if (lane == 0) s.cnt[ui < n_u ? s.ulist[ui] : n_act] = run;
    }
    __syncthreads();
    const int total = block_exscan<TH>(s.cnt, n_act + 1, sh);
    if (total == 0) return true;
    if (total > s.cap) return false;
    if (n_u > U_MAX) { // many members of U: one block-wide minimum each would cost more than the pair buffer with its three rounds of atomicMin
        // the pairs, at their ranks
#pragma unroll
        for (int e = 0; e < MED_E; e++) {
            if (e * TH >= n_act) break;
            const int q = e * TH + t;
            const unsigned long long b = bits[e];
            const int k = q < n_act ? s.act[q] : -1;
            const int pw = q >> 5;
            int own = 0;
            for (int ui = 0; ui < n_u; ui++) {
                const int pu = s.ulist[ui];
                const bool hit = (b >> ui) & 1;
                const unsigned m = __ballot_sync(~0u, hit && q < pu);
                if (!hit) continue;
                const int u = s.act[pu];
                int r, row, partner;
                if (q < pu) { r = s.cnt[pu] + wcnt[ui * npw + pw] + __popc(m & ((1u << lane) - 1)); row = u; partner = k; }
                else { r = s.cnt[q] + own++; row = k; partner = u; }
                s.pa[r] = row;
                s.pb[r] = partner;
                s.pw[r] = fkey(pair_weight(s, has_cf, u, k));
            }
            const bool hj = (b >> 63) != 0;
            const unsigned m = __ballot_sync(~0u, hj);
            if (hj) {
                const int r = s.cnt[n_act] + wcnt[n_u * npw + pw] + __popc(m & ((1u << lane) - 1));
                s.pa[r] = k; // (:565-573)
                s.pb[r] = cj;
                s.pw[r] = fkey(pair_weight(s, has_cf, cj, k));
            }
        }
        __syncthreads();
        draw_and_apply<TH>(s, sh, total, rng_pos, draws);
        return true;
    }
    draw_ties<TH>(s, sh, total, rng_pos, draws);
    // Every pair has a member in U or is (k, j).  That member's candidate is a block-wide minimum over the positions; the other
    // member's candidate belongs to the thread of its position, which updates it on the spot.  No pair buffer, no atomics.
    for (int ui = 0; ui <= n_u; ui++) {
        const bool jrow = ui == n_u;
        const int pu = jrow ? n_act : s.ulist[ui]; // (row j comes after every position)
        const int u = jrow ? cj : s.act[pu];
        unsigned bw = ~0u;
        unsigned long long bt = ~0ull;
        int bq = IMAX; // this thread's best pair with u
#pragma unroll
        for (int e = 0; e < MED_E; e++) {
            if (e * TH >= n_act) break;
            const int q = e * TH + t;
            const unsigned long long b = bits[e];
            const bool hit = jrow ? (b >> 63) != 0 : ((b >> ui) & 1) != 0;
            const bool urow = q < pu; // the pair is met in u's row (else in q's own row, after the earlier members of U)
            const unsigned m = __ballot_sync(~0u, hit && urow);
            if (!hit) continue;
            const int k = s.act[q];
            const int r = urow ? s.cnt[pu] + wcnt[ui * npw + (q >> 5)] + __popc(m & ((1u << lane) - 1))
                               : s.cnt[q] + __popcll(b & ((1ull << ui) - 1));
            const unsigned wk = fkey(pair_weight(s, has_cf, u, k));
            const unsigned long long tie = s.ptie[r];
            if (wk < s.cand_dist[k] || (wk == s.cand_dist[k] && tie < s.cand_tie[k])) {
                s.cand_dist[k] = wk;
                s.cand_tie[k] = tie;
                s.cand_a[k] = (urow && !jrow) ? u : k; // (row, partner); the new cluster's row: (partner, j)   (:565-573)
                s.cand_b[k] = (urow && !jrow) ? k : u;
            }
            if (wk < bw || (wk == bw && tie < bt)) { bw = wk; bt = tie; bq = q; }
        }
        block_lexmin<TH>(bw, bt, bq, sh);
        if (t == 0 && bq != IMAX && (bw < s.cand_dist[u] || (bw == s.cand_dist[u] && bt < s.cand_tie[u]))) {
            const int k = s.act[bq];
            s.cand_dist[u] = bw;
            s.cand_tie[u] = bt;
            s.cand_a[u] = (bq < pu && !jrow) ? u : k;
            s.cand_b[u] = (bq < pu && !jrow) ? k : u;
        }
        __syncthreads();
    }
    return true;
}

// Phase B, small case: the rows that look for a new minimum, all in one block-wide pass (thread q holds column act[q] of each).
template <int TH, int NR>
__device__ __forceinline__ void rescan_rows(const MMState &s, MMShared &sh, int n_act, int ci, int n_rescan)
{
    const size_t N = s.N;
    const int t = threadIdx.x, lane = t & 31;
    constexpr int E = TH > 512 ? 1 : (TH > 256 ? (NR <= 2 ? 2 : 1) : (NR <= 2 ? 4 : (NR <= 4 ? 2 : 1)));
    for (int qb = 0; qb < n_act; qb += TH * E) {
        int l[E];
        float v[E][NR];
#pragma unroll
        for (int e = 0; e < E; e++) {
            const int q = qb + e * TH + t;
            l[e] = q < n_act ? s.act[q] : -1;
            if (l[e] == ci) l[e] = -1;
        }
#pragma unroll
        for (int e = 0; e < E; e++)
#pragma unroll
            for (int r = 0; r < NR; r++) v[e][r] = (r < n_rescan && l[e] >= 0 && l[e] != sh.rs_cl[r]) ? s.d[sh.rs_cl[r] * N + l[e]] : 0.f;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            if (r >= n_rescan) break;
            const float old = sh.rs_old[r];
            unsigned m = ~0u;
            int pe = IMAX, pl = IMAX;
#pragma unroll
            for (int e = E - 1; e >= 0; e--) { // (descending, so that the smallest position wins)
                const int q = qb + e * TH + t;
                if (l[e] >= 0 && l[e] != sh.rs_cl[r]) {
                    m = min(m, fkey(v[e][r]));
                    if (v[e][r] == old) pe = q;
                    if (v[e][r] < old) pl = q;
                }
            }
            m = __reduce_min_sync(~0u, m);
            pe = __reduce_min_sync(~0u, pe);
            pl = __reduce_min_sync(~0u, pl);
            if (lane == 0) {
                if (m != ~0u) atomicMin(&sh.rs_m[r], m);
                if (pe != IMAX) atomicMin(&sh.rs_pe[r], pe);
                if (pl != IMAX) atomicMin(&sh.rs_pl[r], pl);
            }
        }
    }
}

#ifdef MM_PROF
#define MM_MARK(i) do { if (threadIdx.x == 0) { long long now_ = clock64(); prof[i] += now_ - last_; last_ = now_; } } while (0)
#else
#define MM_MARK(i) do { } while (0)
#endif

template <int TH, bool SMEM>
__global__ void __launch_bounds__(TH, 1) mm_quickbuild_kernel(MMState s, int has_cf_i)
{
    __shared__ MMShared sh;
#ifdef MM_PROF
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last_ = clock64();
#endif
    const bool has_cf = has_cf_i != 0;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const size_t N = s.N;
    const float thr = s.thr, thr_cf = s.thr_cf;
    const float finf = __int_as_float(0x7f800000);

    // The per-cluster arrays are rewritten in every step, and a global store drops the line from L1: kept in global memory
    // every phase starts with a chain of L2 round trips (measured: 12 us per merge step, independent of N).  They live in
    // shared memory whenever they fit (N <= ~5000); what outlives the tree is loaded here and written back at the end.
    float *const g_minv_cf = s.minv_cf, *const g_minv = s.minv;
    int *const g_cand_a = s.cand_a, *const g_cand_b = s.cand_b;
    if (SMEM) { // (a template parameter, not s.use_smem: the compiler then knows these pointers are shared memory and emits LDS / STS)
        extern __shared__ __align__(16) unsigned char dyn[];
        const size_t n4 = ((size_t)s.N + 3) & ~(size_t)3;
        s.cand_tie = (unsigned long long *)dyn;
        unsigned *w = (unsigned *)(dyn + 8 * n4);
        s.cand_dist = w;             w += n4;
        s.cand_a = (int *)w;         w += n4;
        s.cand_b = (int *)w;         w += n4;
        s.act = (int *)w;            w += n4;
        s.flag = (int *)w;           w += n4;
        s.minv = (float *)w;         w += n4;
        s.minv_cf = (float *)w;      w += n4;
        s.cnt = (int *)w;            // n4 + 4 entries
        for (int k = t; k < s.N; k += TH) { s.minv_cf[k] = g_minv_cf[k]; s.cand_a[k] = g_cand_a[k]; s.cand_b[k] = g_cand_b[k]; }
    }

    // rng.seed(1)
    if (t == 0) {
        unsigned v = 1u;
        sh.mt[0] = v;
        for (int q = 1; q < 624; q++) { v = 1812433253u * (v ^ (v >> 30)) + (unsigned)q; sh.mt[q] = v; }
    }
    int rng_pos = 312;
    long long draws = 0, general_steps = 0, medium_steps = 0;
    bool last_was_big = false;
    int n_act = s.N;
    for (int k = t; k < s.N; k += TH) {
        s.act[k] = k; s.conv[k] = k; s.size[k] = 1.0f; s.minv_sym[k] = finf; s.flag[k] = 0;
        s.cand_dist[k] = KINF; s.cand_tie[k] = DINF;
    }
    __syncthreads();
    if (s.pre_init) {
        // Initialize's O(N^2) passes were done across the whole GPU (mm_init_* kernels): row minima in g_minv / g_minv_cf (the latter
        // copied above), the feasible pairs in the pair buffer in the reference's order
        if (SMEM)
            for (int k = t; k < s.N; k += TH) s.minv[k] = g_minv[k];
    } else {
        // Initialize: row minima (+ the prior's, which start from what the previous tree left)
        for (int k = warp; k < s.N; k += (TH / 32)) {
            float m = finf, mc = finf;
            for (int lb = 0; lb < s.N; lb += 128) {
                float v[4], c[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int l = lb + 32 * e + lane;
                    const bool ok = l < s.N && l != k;
                    v[e] = ok ? s.d[k * N + l] : finf;
                    c[e] = (ok && has_cf) ? s.cf[k * N + l] : finf;
                }
#pragma unroll
                for (int e = 0; e < 4; e++) { m = fminf(m, v[e]); mc = fminf(mc, c[e]); }
            }
            for (int o = 16; o; o >>= 1) { m = fminf(m, __shfl_xor_sync(~0u, m, o)); mc = fminf(mc, __shfl_xor_sync(~0u, mc, o)); }
            if (lane == 0) {
                s.minv[k] = __fadd_rn(m, thr);
                if (has_cf) s.minv_cf[k] = __fadd_rn(fminf(s.minv_cf[k], mc), thr_cf);
            }
        }
    }
    __syncthreads();
    if (s.pre_init && s.init_cnt[s.N + 1] <= s.cap) draw_and_apply<TH>(s, sh, s.init_cnt[s.N + 1], rng_pos, draws);
    else meet_pairs<TH>(s, sh, has_cf, true, n_act, -1, -1, 0, rng_pos, draws); // (more pairs than the buffer holds: in segments)
    __syncthreads();

    MM_MARK(0);
    bool use_sym = false;
    long long first_sym = -1, sym_steps = 0;
    unsigned sbest_d = KINF; // best_sym_candidate
    int sbest_a = -1, sbest_b = -1;

    for (int node = s.N; node < 2 * s.N - 1; node++) {
        // F (of the previous step): the best candidate over the active clusters
        unsigned ba = KINF;
        unsigned long long bb = DINF;
        int bp = IMAX;
        for (int p = t; p < n_act; p += TH) {
            const int k = s.act[p];
            const unsigned da = s.cand_dist[k];
            const unsigned long long db = s.cand_tie[k];
            if (da < ba || (da == ba && (db < bb || (db == bb && p < bp)))) { ba = da; bb = db; bp = p; }
        }
        block_lexmin<TH>(ba, bb, bp, sh);
        int ci, cj;
        if (ba == KINF) { // no mutually minimal pair: symmetric fallback (:1244-1256)
            if (!use_sym) {
                use_sym = true;
                first_sym = node - s.N;
                // InitializeSym
                for (size_t e = t; e < (size_t)n_act * n_act; e += TH) {
                    const int x = s.act[e / n_act], y = s.act[e % n_act];
                    if (x < y) {
                        const float v = __fadd_rn(s.d[x * N + y], s.dT[x * N + y]);
                        s.sym[x * N + y] = v;
                        s.sym[y * N + x] = v;
                    }
                }
                __syncthreads();
                for (int p = warp; p < n_act; p += (TH / 32)) {
                    const int x = s.act[p];
                    unsigned m = ~0u;
                    int mq = IMAX;
                    for (int q = lane; q < n_act; q += 32) {
                        const int l = s.act[q];
                        if (l == x) continue;
                        const unsigned v = fkey(s.sym[x * N + l]);
                        if (v < m) { m = v; mq = q; }
                    }
                    for (int o = 16; o; o >>= 1) {
                        const unsigned om = __shfl_xor_sync(~0u, m, o);
                        const int oq = __shfl_xor_sync(~0u, mq, o);
                        if (om < m || (om == m && oq < mq)) { m = om; mq = oq; }
                    }
                    if (lane == 0) {
                        const float mv = (mq == IMAX || m >= KINF) ? finf : funkey(m); // min_values_sym starts at +inf for every cluster
                        s.minv_sym[x] = mv;
                        s.csym_dist[x] = mv;
                        if (mv < finf) { s.csym_a[x] = x; s.csym_b[x] = s.act[mq]; }
                    }
                }
                __syncthreads();
                unsigned a = KINF; unsigned long long b = 0; int c = IMAX;
                for (int p = t; p < n_act; p += TH) {
                    const unsigned v = fkey(s.csym_dist[s.act[p]]);
                    if (v < a || (v == a && p < c)) { a = v; c = p; }
                }
                block_lexmin<TH>(a, b, c, sh);
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
        MM_MARK(1);
        const float si = s.size[ci], sj = s.size[cj], sum = __fadd_rn(si, sj);
        if (t == 0) {
            s.merges[2 * (node - s.N)] = s.conv[ci];
            s.merges[2 * (node - s.N) + 1] = s.conv[cj];
            sh.n_rescan = 0;
            sh.total = 0;
        }
        __syncthreads();

        // A: rows / columns j
        float *di = s.d + ci * N, *dj = s.d + cj * N, *dTi = s.dT + ci * N, *dTj = s.dT + cj * N;
        float min_j = finf, min_cf = finf;
        constexpr int EA = TH > 512 ? 1 : (TH > 256 ? 2 : 4); // positions per thread and round, loads issued together
        for (int p0 = t; p0 < n_act; p0 += TH * EA) {
            int k[EA];
            float dkj[EA], dki[EA], dik[EA], djk[EA], ckj[EA], cki[EA], cik[EA], cjk[EA];
#pragma unroll
            for (int e = 0; e < EA; e++) {
                const int p = p0 + e * TH;
                k[e] = p < n_act ? s.act[p] : -1;
                if (k[e] == ci) sh.pos_i = p;
                if (k[e] == ci || k[e] == cj) k[e] = -1;
            }
#pragma unroll
            for (int e = 0; e < EA; e++) {
                if (k[e] < 0) continue;
                dkj[e] = dTj[k[e]]; dki[e] = dTi[k[e]]; dik[e] = di[k[e]]; djk[e] = dj[k[e]];
                if (has_cf) { ckj[e] = s.cfT[cj * N + k[e]]; cki[e] = s.cfT[ci * N + k[e]]; cik[e] = s.cf[ci * N + k[e]]; cjk[e] = s.cf[cj * N + k[e]]; }
            }
#pragma unroll
            for (int e = 0; e < EA; e++) {
                if (k[e] < 0) continue;
                const int kk = k[e], p = p0 + e * TH;
                if (has_cf) {
                    float njk = cjk[e];
                    if (cik[e] != cjk[e]) { njk = mix(si, cik[e], sj, cjk[e], sum); s.cf[cj * N + kk] = njk; s.cfT[kk * N + cj] = njk; }
                    if (cki[e] != ckj[e]) { const float nkj = mix(si, cki[e], sj, ckj[e], sum); s.cf[kk * N + cj] = nkj; s.cfT[cj * N + kk] = nkj; }
                    min_cf = fminf(min_cf, njk);
                }
                float njk = djk[e];
                if (dik[e] != djk[e]) { njk = mix(si, dik[e], sj, djk[e], sum); dj[kk] = njk; s.dT[kk * N + cj] = njk; }
                if (dki[e] != dkj[e]) { const float nkj = mix(si, dki[e], sj, dkj[e], sum); s.d[kk * N + cj] = nkj; dTj[kk] = nkj; }
                min_j = fminf(min_j, njk);
                const float mk = s.minv[kk];
                int f = 0;
                const int ca = s.cand_a[kk], cb = s.cand_b[kk];
                if (ca == cj || cb == cj || ca == ci || cb == ci) f = 1;
                if (dkj[e] != dki[e]) {
                    const float base = __fsub_rn(mk, thr);
                    if ((double)fabsf(__fsub_rn(base, dkj[e])) < 1e-4 || (double)fabsf(__fsub_rn(base, dki[e])) < 1e-4) {
                        const int idx = atomicAdd(&sh.n_rescan, 1);
                        s.rescan[idx] = p;
                        if (idx < RS_MAX) { sh.rs_pos[idx] = p; sh.rs_cl[idx] = kk; sh.rs_old[idx] = base; sh.rs_m[idx] = ~0u; sh.rs_pe[idx] = IMAX; sh.rs_pl[idx] = IMAX; }
                    }
                }
                s.flag[kk] = f;
            }
        }
        block_min2<TH>(min_j, min_cf, sh);
        if (t == 0) {
            s.minv[cj] = __fadd_rn(min_j, thr);
            if (has_cf) s.minv_cf[cj] = __fadd_rn(min_cf, thr_cf);
            s.flag[cj] = 0;
            s.cand_dist[cj] = KINF; // mcandidates[j] starts over (:541-542)
            s.cand_tie[cj] = DINF;
        }
        __syncthreads();

        MM_MARK(2);
        // B: new row minima
        const int n_rescan = sh.n_rescan;
        if (n_rescan > 0 && n_rescan <= RS_MAX && !s.force_general) {
            // all rows in one block-wide pass: thread q holds column act[q] of each of them
            if (n_rescan == 1) rescan_rows<TH, 1>(s, sh, n_act, ci, n_rescan);
            else if (n_rescan == 2) rescan_rows<TH, 2>(s, sh, n_act, ci, n_rescan);
            else if (n_rescan <= 4) rescan_rows<TH, 4>(s, sh, n_act, ci, n_rescan);
            else rescan_rows<TH, RS_MAX>(s, sh, n_act, ci, n_rescan);
            __syncthreads();
            if (t < n_rescan) {
                const int k = sh.rs_cl[t];
                const float m = sh.rs_m[t] == ~0u ? finf : funkey(sh.rs_m[t]);
                const float mk = (sh.rs_pe[t] != IMAX && sh.rs_pe[t] < sh.rs_pl[t]) ? sh.rs_old[t] : m; // the scan stops at the old minimum (:337-339)
                s.minv[k] = __fadd_rn(mk, thr);
                s.flag[k] |= 2;
            }
        } else {
            for (int it = warp; it < n_rescan; it += (TH / 32)) {
                const int k = s.act[s.rescan[it]];
                const float *dk = s.d + k * N;
                const float old = __fsub_rn(s.minv[k], thr);
                unsigned m = ~0u;
                int pe = IMAX, pl = IMAX;
                for (int qb = 0; qb < n_act; qb += 128) {
                    float v[4];
                    bool on[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int q = qb + 32 * e + lane;
                        const int l = q < n_act ? s.act[q] : -1;
                        on[e] = l >= 0 && l != ci && l != k;
                        v[e] = on[e] ? dk[l] : 0.f;
                    }
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (on[e]) {
                            const int q = qb + 32 * e + lane;
                            m = min(m, fkey(v[e]));
                            if (v[e] == old) pe = min(pe, q);
                            if (v[e] < old) pl = min(pl, q);
                        }
                }
                m = __reduce_min_sync(~0u, m);
                pe = __reduce_min_sync(~0u, pe);
                pl = __reduce_min_sync(~0u, pl);
                if (lane == 0) {
                    const float mv = m == ~0u ? finf : funkey(m);
                    const float mk = (pe != IMAX && pe < pl) ? old : mv;
                    s.minv[k] = __fadd_rn(mk, thr);
                    s.flag[k] |= 2;
                }
            }
        }
        __syncthreads();

        MM_MARK(3);
        // C: U in order; its members' candidates start over
        int n_u = 0;
        for (int pb = 0; pb < n_act; pb += TH) {
            const int p = pb + t;
            const int k = p < n_act ? s.act[p] : -1;
            const bool in = k >= 0 && k != ci && k != cj && s.flag[k] != 0;
            const unsigned m = __ballot_sync(~0u, in);
            if (lane == 0) sh.scan[warp] = __popc(m);
            __syncthreads();
            const int wc = lane < TH / 32 ? sh.scan[lane] : 0;
            const int off = n_u + __reduce_add_sync(~0u, lane < warp ? wc : 0), tot = __reduce_add_sync(~0u, wc);
            if (in) {
                const int idx = off + __popc(m & ((1u << lane) - 1));
                s.ulist[idx] = p;
                if (idx < U_MAX) { sh.u_pos[idx] = p; sh.u_cl[idx] = k; sh.u_minv[idx] = s.minv[k]; }
                s.flag[k] |= 4;
                s.cand_dist[k] = KINF;
                s.cand_tie[k] = DINF;
            }
            n_u += tot;
            __syncthreads();
        }

        MM_MARK(4);
        // D, E
        bool done = false;
        // (a step that follows one with hundreds of pairs usually has hundreds too: blocks of identical haplotypes; skip the attempt)
        if (n_u <= U_MAX && !s.force_general && !last_was_big) {
            if (n_u == 0) done = meet_pairs_small<TH, 0>(s, sh, has_cf, n_act, ci, cj, n_u, rng_pos, draws);
            else if (n_u == 1) done = meet_pairs_small<TH, 1>(s, sh, has_cf, n_act, ci, cj, n_u, rng_pos, draws);
            else if (n_u == 2) done = meet_pairs_small<TH, 2>(s, sh, has_cf, n_act, ci, cj, n_u, rng_pos, draws);
            else if (n_u <= 4) done = meet_pairs_small<TH, 4>(s, sh, has_cf, n_act, ci, cj, n_u, rng_pos, draws);
            else done = meet_pairs_small<TH, U_MAX>(s, sh, has_cf, n_act, ci, cj, n_u, rng_pos, draws);
        }
        if (!done && !(s.force_general & 1)) {
            __syncthreads();
            const long long before = draws;
            done = meet_pairs_medium<TH>(s, sh, has_cf, n_act, ci, cj, n_u, rng_pos, draws);
            medium_steps += done;
            last_was_big = done && draws - before > PAIR_MAX;
        }
        if (!done) {
            general_steps++;
            __syncthreads();
            meet_pairs<TH>(s, sh, has_cf, false, n_act, ci, cj, n_u, rng_pos, draws);
        }
        __syncthreads();

        MM_MARK(5);
        // the symmetric matrix follows once it is in use (CoalesceSym)
        if (use_sym) {
            if (t == 0) sh.n_rescan = 0;
            __syncthreads();
            float *yi = s.sym + ci * N, *yj = s.sym + cj * N;
            unsigned ja = KINF; unsigned long long jb = 0; int jp = IMAX;
            for (int p = t; p < n_act; p += TH) {
                const int k = s.act[p];
                if (k == ci || k == cj) continue;
                float *yk = s.sym + k * N;
                const float dkj = yk[cj], dki = yk[ci], dik = yi[k], djk = yj[k];
                float njk = djk;
                if (dik != djk) { njk = mix(si, dik, sj, djk, sum); yj[k] = njk; }
                if (dki != dkj) yk[cj] = mix(si, dki, sj, dkj, sum);
                const unsigned nb = fkey(njk);
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
            block_lexmin<TH>(ja, jb, jp, sh);
            __syncthreads();
            if (t == 0) {
                const float mv = (jp == IMAX || ja >= KINF) ? finf : funkey(ja);
                s.minv_sym[cj] = mv;
                s.csym_dist[cj] = mv;
                if (mv < finf) { s.csym_a[cj] = s.act[jp]; s.csym_b[cj] = cj; }
            }
            const int n_rs = sh.n_rescan;
            for (int it = warp; it < n_rs; it += (TH / 32)) {
                const int k = s.act[s.rescan[it]];
                const float *yk = s.sym + k * N;
                const float old = s.minv_sym[k];
                unsigned m = ~0u;
                int mq = IMAX, pe = IMAX, pl = IMAX;
                for (int q = lane; q < n_act; q += 32) {
                    const int l = s.act[q];
                    if (l == ci || l == k) continue;
                    const float v = yk[l];
                    const unsigned vk = fkey(v);
                    if (vk < m) { m = vk; mq = q; }
                    if (v == old) pe = min(pe, q);
                    if (v < old) pl = min(pl, q);
                }
                for (int o = 16; o; o >>= 1) {
                    const unsigned om = __shfl_xor_sync(~0u, m, o);
                    const int oq = __shfl_xor_sync(~0u, mq, o);
                    if (om < m || (om == m && oq < mq)) { m = om; mq = oq; }
                }
                pe = __reduce_min_sync(~0u, pe);
                pl = __reduce_min_sync(~0u, pl);
                if (lane == 0) {
                    float mv = (mq == IMAX || m >= KINF) ? finf : funkey(m);
                    if (pe != IMAX && pe < pl) { mv = old; mq = pe; }
                    s.minv_sym[k] = mv;
                    s.csym_dist[k] = mv;
                    if (mv < finf) { s.csym_a[k] = k; s.csym_b[k] = s.act[mq]; }
                }
            }
            __syncthreads();
            // best_sym_candidate: first cluster in order with the smallest candidate, the new cluster last
            unsigned a = KINF; unsigned long long b = 0; int c = IMAX;
            for (int p = t; p < n_act; p += TH) {
                const int k = s.act[p];
                if (k == ci) continue;
                const unsigned v = fkey(s.csym_dist[k]);
                const int ord = (k == cj) ? n_act : p;
                if (v < a || (v == a && ord < c)) { a = v; c = ord; }
            }
            block_lexmin<TH>(a, b, c, sh);
            sbest_d = a;
            if (a != KINF) {
                const int k = (c == n_act) ? cj : s.act[c];
                sbest_a = s.csym_a[k];
                sbest_b = s.csym_b[k];
            }
            __syncthreads();
        }

        // bookkeeping: the new cluster's size and name, cluster i leaves the list
        if (t == 0) {
            s.size[cj] = sum;
            s.conv[cj] = node;
        }
        const int at = sh.pos_i;
        for (int b = at; b < n_act - 1; b += TH) {
            const int p = b + t;
            int v = 0;
            if (p < n_act - 1) v = s.act[p + 1];
            __syncthreads();
            if (p < n_act - 1) s.act[p] = v;
        }
        n_act--;
        __syncthreads();
        MM_MARK(6);
    }
    if (SMEM)
        for (int k = t; k < s.N; k += TH) { g_minv_cf[k] = s.minv_cf[k]; g_cand_a[k] = s.cand_a[k]; g_cand_b[k] = s.cand_b[k]; }
    if (t == 0) {
#ifdef MM_PROF
        for (int q = 0; q < 7; q++) s.info[4 + q] = prof[q];
#endif
        s.info[0] = draws;
        s.info[1] = first_sym;
        s.info[2] = sym_steps;
        s.info[3] = general_steps;
        s.info[12] = medium_steps;
    }
}

// ---- Initialize across the whole GPU --------------------------------------------------------------------------------------
// The tree's CTA spent 5-7 % of a tree in the three O(N^2) passes of Initialize (row minima, count, emit).  They have no order
// dependence beyond a prefix sum, so a warp per row on all SMs does them in front of the tree kernel.
__global__ void mm_init_rowmin_kernel(MMState s, int has_cf)
{
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= s.N) return;
    const size_t N = s.N;
    const float finf = __int_as_float(0x7f800000);
    float m = finf, mc = finf;
    for (int lb = 0; lb < s.N; lb += 128) {
        float v[4], c[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int l = lb + 32 * e + lane;
            const bool ok = l < s.N && l != k;
            v[e] = ok ? s.d[k * N + l] : finf;
            c[e] = (ok && has_cf) ? s.cf[k * N + l] : finf;
        }
#pragma unroll
        for (int e = 0; e < 4; e++) { m = fminf(m, v[e]); mc = fminf(mc, c[e]); }
    }
    for (int o = 16; o; o >>= 1) { m = fminf(m, __shfl_xor_sync(~0u, m, o)); mc = fminf(mc, __shfl_xor_sync(~0u, mc, o)); }
    if (lane == 0) {
        s.minv[k] = __fadd_rn(m, s.thr);
        if (has_cf) s.minv_cf[k] = __fadd_rn(fminf(s.minv_cf[k], mc), s.thr_cf); // (starts from what the previous tree left)
    }
}

// row p against the rows above it (Initialize's order: :87-143); EMIT = false counts, EMIT = true writes at the row's offset
template <bool EMIT>
__global__ void mm_init_pairs_kernel(MMState s, int has_cf)
{
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= s.N) return;
    if (EMIT && s.init_cnt[s.N + 1] > s.cap) return; // (the tree kernel does it itself, in segments)
    const size_t N = s.N;
    const int x = p;
    const float mx = s.minv[x];
    int count = 0;
    const int base = EMIT ? s.init_cnt[p] : 0;
    for (int qb = p + 1; qb < s.N; qb += 128) {
        float v[4], vt[4];
        bool ok[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int y = qb + 32 * c + lane;
            ok[c] = y < s.N;
            v[c] = ok[c] ? s.d[x * N + y] : 0.f;
            vt[c] = ok[c] ? s.dT[x * N + y] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int y = qb + 32 * c + lane;
            ok[c] = ok[c] && v[c] <= mx && vt[c] <= s.minv[y];
            const unsigned m = __ballot_sync(~0u, ok[c]);
            if (EMIT && ok[c]) {
                const int r = base + count + __popc(m & ((1u << lane) - 1));
                s.pa[r] = x;
                s.pb[r] = y;
                s.pw[r] = fkey(pair_weight(s, has_cf != 0, x, y));
            }
            count += __popc(m);
        }
    }
    if (!EMIT && lane == 0) s.init_cnt[p] = count;
}

// exclusive prefix sum of init_cnt[0..N) in place, the total into init_cnt[N] and init_cnt[N + 1]
__global__ void mm_init_scan_kernel(MMState s)
{
    __shared__ int part[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, n = s.N, per = (n + 1023) / 1024;
    const int b = min(t * per, n), e = min(b + per, n);
    int sum = 0;
    for (int q = b; q < e; q++) sum += s.init_cnt[q];
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(~0u, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) part[warp] = inc;
    __syncthreads();
    const int wc = part[lane];
    const int wbase = __reduce_add_sync(~0u, lane < warp ? wc : 0), total = __reduce_add_sync(~0u, wc);
    int run = wbase + inc - sum;
    for (int q = b; q < e; q++) { const int v = s.init_cnt[q]; s.init_cnt[q] = run; run += v; }
    if (t == 0) { s.init_cnt[n] = total; s.init_cnt[n + 1] = total; }
}

// dT = d^T (z = 0) and cfT = cf^T (z = 1), 32 x 32 tiles through shared memory
__global__ void mm_transpose_kernel(const float *d, float *dT, const float *cf, float *cfT, int N)
{
    __shared__ float tile[32][33];
    const float *src = blockIdx.z ? cf : d;
    float *dst = blockIdx.z ? cfT : dT;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int x = x0 + threadIdx.x, y = y0 + r;
        if (x < N && y < N) tile[r][threadIdx.x] = src[(size_t)y * N + x];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int x = y0 + threadIdx.x, y = x0 + r;
        if (x < N && y < N) dst[(size_t)y * N + x] = tile[threadIdx.x][r];
    }
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------------
struct rp_minmatch {
    int device = 0, N = 0;
    MMState s{};
    void *block = nullptr; // one allocation behind every array
    size_t dyn_smem = 0;
    int threads = 256;
    int *h_merges = nullptr; // pinned
    long long *h_info = nullptr;
    cudaStream_t stream = nullptr;
    float last_ms = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr; // e2: end of the call's work; waited for without spinning
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

static int mm_create(int device, int N, float threshold, float threshold_cf, rp_minmatch *h);

#define RP_TRY_MM(expr)               \
    do {                              \
        int rc_ = (expr);             \
        if (rc_ != RP_OK) return rc_; \
    } while (0)

// the state of a freshly constructed MinMatch object: min_values_CF = 0 (vector::resize), candidates name nobody (lin1 = lin2 = -1)
static int mm_reset_state(rp_minmatch *h)
{
    const MMState &s = h->s;
    const size_t n = 4 * (size_t)h->N;
    MM_CUDA(cudaMemset(s.minv_cf, 0, n));
    MM_CUDA(cudaMemset(s.cand_a, 0xff, n));
    MM_CUDA(cudaMemset(s.cand_b, 0xff, n));
    MM_CUDA(cudaMemset(s.csym_a, 0xff, n));
    MM_CUDA(cudaMemset(s.csym_b, 0xff, n));
    MM_CUDA(cudaMemset(s.csym_dist, 0x7f, n));
    return RP_OK;
}

extern "C" int rp_minmatch_reset(rp_minmatch *h)
{
    if (!h) return rp::api_fail(RP_EINVAL, "rp_minmatch_reset: null argument");
    MM_CUDA(cudaSetDevice(h->device));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    return mm_reset_state(h);
}

extern "C" int rp_minmatch_create_thresholds(int device, int N, float threshold, float threshold_cf, rp_minmatch **out)
{
    if (!out || N < 2) return rp::api_fail(RP_EINVAL, "rp_minmatch_create: bad argument");
    *out = nullptr;
    int ndev = 0;
    MM_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return rp::api_fail(RP_ENODEVICE, "rp_minmatch_create: no such device");
    MM_CUDA(cudaSetDevice(device));
    auto *h = new rp_minmatch;
    h->device = device;
    h->N = N;
    const int rc = mm_create(device, N, threshold, threshold_cf, h);
    if (rc != RP_OK) {
        rp_minmatch_destroy(h); // (keeps the error message: destroy does not touch it)
        return rc;
    }
    *out = h;
    return RP_OK;
}

static int mm_create(int device, int N, float threshold, float threshold_cf, rp_minmatch *h)
{
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
    const size_t o_tie = take(8 * (size_t)N), o_key2 = take(8 * (size_t)N), o_ptie = take(8 * (size_t)s.cap), o_info = take(128);
    const size_t o_d = take(4 * nn), o_cf = take(4 * nn), o_sym = take(4 * nn), o_dT = take(4 * nn), o_cfT = take(4 * nn);
    const size_t o_f[5] = {take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N)};
    const size_t o_i[12] = {take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N),
                            take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N), take(4 * (size_t)N),
                            take(4 * (size_t)N), take(4 * (size_t)(N + 2)), take(4 * (size_t)N), take(4 * (size_t)N)};
    const size_t o_pa = take(4 * (size_t)s.cap), o_pb = take(4 * (size_t)s.cap), o_pw = take(4 * (size_t)s.cap);
    const size_t o_icnt = take(4 * ((size_t)N + 4));
    const size_t o_merges = take(8 * (size_t)N), o_wcnt = take(4 * 64 * ((size_t)N / 32 + 2));
    cudaError_t e = cudaMalloc(&h->block, off);
    if (e != cudaSuccess) {
        h->block = nullptr;
        return rp::api_fail(RP_ENOMEM, std::string("rp_minmatch_create: cudaMalloc: ") + cudaGetErrorString(e));
    }
    char *b = (char *)h->block;
    s.cand_tie = (unsigned long long *)(b + o_tie);
    s.key2 = (unsigned long long *)(b + o_key2);
    s.ptie = (unsigned long long *)(b + o_ptie);
    s.info = (long long *)(b + o_info);
    s.d = (float *)(b + o_d); s.cf = (float *)(b + o_cf); s.sym = (float *)(b + o_sym); s.dT = (float *)(b + o_dT); s.cfT = (float *)(b + o_cfT);
    s.minv_cf = (float *)(b + o_f[0]); s.csym_dist = (float *)(b + o_f[1]); s.minv = (float *)(b + o_f[2]);
    s.minv_sym = (float *)(b + o_f[3]); s.size = (float *)(b + o_f[4]);
    s.cand_a = (int *)(b + o_i[0]); s.cand_b = (int *)(b + o_i[1]); s.cand_dist = (unsigned *)(b + o_i[2]);
    s.csym_a = (int *)(b + o_i[3]); s.csym_b = (int *)(b + o_i[4]); s.conv = (int *)(b + o_i[5]); s.act = (int *)(b + o_i[6]);
    s.flag = (int *)(b + o_i[7]); s.rescan = (int *)(b + o_i[8]); s.cnt = (int *)(b + o_i[9]); s.ulist = (int *)(b + o_i[10]);
    s.key1 = (unsigned *)(b + o_i[11]);
    s.pa = (int *)(b + o_pa); s.pb = (int *)(b + o_pb); s.pw = (unsigned *)(b + o_pw);
    s.merges = (int *)(b + o_merges);
    s.wcnt = (int *)(b + o_wcnt);
    s.init_cnt = (int *)(b + o_icnt);
    s.pre_init = getenv("RP_MINMATCH_PRE_INIT") ? atoi(getenv("RP_MINMATCH_PRE_INIT")) : 1;
    RP_TRY_MM(mm_reset_state(h));
    s.force_general = getenv("RP_MINMATCH_GENERAL") ? atoi(getenv("RP_MINMATCH_GENERAL")) : 0;
    {
        const size_t n4 = ((size_t)N + 3) & ~(size_t)3, want = 40 * n4 + 16;
        int max_optin = 0;
        MM_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        // one CTA per tree; the merge loop is bound by instruction issue on its SM, and everything every warp does per phase
        // (reductions, barriers) counts once per warp: few warps for small N (measured: N=200 1.24 / 1.30 / 1.67 ms per tree with
        // 256 / 512 / 1024 threads, N=1000 10.4 / 9.7 / 11.0 ms, N=5000 134 / 101 / 100 ms on synthetic input and 160 / 146 ms
        // with 512 / 1024 threads on real, tie-rich matrices)
        h->threads = getenv("RP_MINMATCH_THREADS") ? atoi(getenv("RP_MINMATCH_THREADS")) : (N < 512 ? 256 : (N <= 2560 ? 512 : 1024));
        if (h->threads != 256 && h->threads != 512) h->threads = 1024;
        cudaFuncAttributes fa;
        const void *fn = h->threads == 256 ? (const void *)mm_quickbuild_kernel<256, true>
                         : h->threads == 512 ? (const void *)mm_quickbuild_kernel<512, true> : (const void *)mm_quickbuild_kernel<1024, true>;
        MM_CUDA(cudaFuncGetAttributes(&fa, fn));
        if (!getenv("RP_MINMATCH_NO_SMEM") && want + fa.sharedSizeBytes <= (size_t)max_optin) {
            // (a per-function setting shared by every handle of the process, whatever its N: the device's maximum, once and for all)
            MM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin - (int)fa.sharedSizeBytes));
            h->dyn_smem = want;
            s.use_smem = 1;
        }
    }
    MM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    MM_CUDA(cudaEventCreate(&h->e0));
    MM_CUDA(cudaEventCreate(&h->e1));
    MM_CUDA(cudaEventCreateWithFlags(&h->e2, cudaEventBlockingSync | cudaEventDisableTiming));
    MM_CUDA(cudaMallocHost(&h->h_merges, 8 * (size_t)N));
    MM_CUDA(cudaMallocHost(&h->h_info, 128));
    return RP_OK;
}

extern "C" void rp_minmatch_destroy(rp_minmatch *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->e0) cudaEventDestroy(h->e0);
    if (h->e1) cudaEventDestroy(h->e1);
    if (h->e2) cudaEventDestroy(h->e2);
    if (h->h_merges) cudaFreeHost(h->h_merges);
    if (h->h_info) cudaFreeHost(h->h_info);
    if (h->block) cudaFree(h->block);
    delete h;
}

static int mm_run(rp_minmatch *h, bool has_prior, int *merges, rp_minmatch_stats *st)
{
    MM_CUDA(cudaEventRecord(h->e0, h->stream));
    {
        const dim3 grid((h->N + 31) / 32, (h->N + 31) / 32, has_prior ? 2 : 1);
        mm_transpose_kernel<<<grid, dim3(32, 8), 0, h->stream>>>(h->s.d, h->s.dT, h->s.cf, h->s.cfT, h->N);
    }
    if (h->s.pre_init) {
        const int rows = (h->N + 7) / 8, hc0 = has_prior ? 1 : 0;
        mm_init_rowmin_kernel<<<rows, 256, 0, h->stream>>>(h->s, hc0);
        mm_init_pairs_kernel<false><<<rows, 256, 0, h->stream>>>(h->s, hc0);
        mm_init_scan_kernel<<<1, 1024, 0, h->stream>>>(h->s);
        mm_init_pairs_kernel<true><<<rows, 256, 0, h->stream>>>(h->s, hc0);
    }
    const int hc = has_prior ? 1 : 0;
#define MM_LAUNCH(TH_)                                                                                             \
    do {                                                                                                           \
        if (h->s.use_smem) mm_quickbuild_kernel<TH_, true><<<1, TH_, h->dyn_smem, h->stream>>>(h->s, hc);          \
        else mm_quickbuild_kernel<TH_, false><<<1, TH_, 0, h->stream>>>(h->s, hc);                                 \
    } while (0)
    if (h->threads == 256) MM_LAUNCH(256);
    else if (h->threads == 512) MM_LAUNCH(512);
    else MM_LAUNCH(1024);
#undef MM_LAUNCH
    MM_CUDA(cudaGetLastError());
    MM_CUDA(cudaEventRecord(h->e1, h->stream));
    MM_CUDA(cudaMemcpyAsync(h->h_merges, h->s.merges, 8 * (size_t)(h->N - 1), cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaMemcpyAsync(h->h_info, h->s.info, 128, cudaMemcpyDeviceToHost, h->stream));
    // a tree takes milliseconds and a consumer may drive many handles from as many host threads (one tree per SM): block, do not spin
    MM_CUDA(cudaEventRecord(h->e2, h->stream));
    MM_CUDA(cudaEventSynchronize(h->e2));
    memcpy(merges, h->h_merges, 8 * (size_t)(h->N - 1));
    MM_CUDA(cudaEventElapsedTime(&h->last_ms, h->e0, h->e1));
    if (st) {
        st->ms_kernel = h->last_ms;
        st->draws = h->h_info[0];
        st->first_fallback_step = (int)h->h_info[1];
        st->fallback_steps = (int)h->h_info[2];
        st->general_steps = (int)h->h_info[3];
        st->medium_steps = (int)h->h_info[12];
        st->launches = h->s.pre_init ? 6 : 2;
    }
#ifdef MM_PROF
    fprintf(stderr, "mm_prof cycles: init %lld | F %lld A %lld B %lld C %lld DE %lld sym %lld book %lld\n", h->h_info[4], h->h_info[5], h->h_info[6],
            h->h_info[7], h->h_info[8], h->h_info[9], h->h_info[10], h->h_info[11] - 0);
#endif
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
