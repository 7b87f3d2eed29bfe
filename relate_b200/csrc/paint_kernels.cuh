// paint_kernels.cuh — sm_100a kernels for Relate's chromosome-painting hot path.
//
// What they replace: FastPainting::PaintSteppingStones
// (/root/reference/include/src/fast_painting.cpp:18-618).  Design notes are in DESIGN.md;
// in short:
//   * genotypes live in HBM as bits: G[s][n/32] (SNP-major, one coalesced row per visited
//     site) and GT[k][s/32] (haplotype-major, used once to enumerate a target's sites);
//   * a job is (target k, direction).  Forward and backward are independent recursions,
//     and the backward one is carried in the emission-weighted variable g = b * m_s so that
//     both directions are the same 3-FP32-op update  x <- (x + R) * (mis ? tau : 1);  S += x;
//   * one team (a warp, or one CTA of several warps for large N) owns the job's N-vector in
//     registers: 32 haplotypes (one genotype word) per register block, rotated so that the
//     target itself always sits in slot 0 of its word;
//   * FP32 adds are issued as packed add.f32x2 (FADD2, sm_100+ only), the mismatch multiply
//     is a predicated FMUL, predicates come 7 at a time from R2P;
//   * per step one warp butterfly (+ one bar.sync and a shared-memory exchange for
//     multi-warp teams) gives every thread the normalising sum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include <cooperative_groups.h>

namespace rp {

// ---------------------------------------------------------------------------------------
// fast_log, bit-exact with /root/reference/include/src/fast_log.hpp:6-22 as an x86-64 SSE
// build evaluates it (one float rounding per operation, no FMA contraction).
__device__ __forceinline__ float fast_log_dev(float val)
{
    int x = __float_as_int(val);
    const int log_2 = ((x >> 23) & 255) - 128;
    x &= ~(255 << 23);
    x += 127 << 23;
    val = __int_as_float(x);
    float t = __fmul_rn(-1.0f / 3, val);
    t = __fadd_rn(t, 2.0f);
    t = __fmul_rn(t, val);
    t = __fsub_rn(t, 2.0f / 3);
    t = __fadd_rn(t, (float)log_2);
    return __fmul_rn(t, 0.69314718f);
}

__global__ void fast_log_kernel(const float *in, float *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fast_log_dev(in[i]);
}

// FP32 issue-rate microbenchmark with the paint step's instruction mix and no memory traffic:
// per pair  x = x + R (packed), x.lo *= m, x.hi *= m (scalar), S = S + x (packed)  -> 6 lane-ops, 4 (PACKED) or
// 6 issue slots.  Used only to measure the roofline denominator.
template <bool PACKED>
__global__ void __launch_bounds__(256) peak_fp32_kernel(float *out, int iters, float m, float r)
{
    float2 a[16];
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = make_float2(1.0f + j + threadIdx.x, 2.0f + j);
    float2 S0 = make_float2(0.f, 0.f), S1 = make_float2(0.f, 0.f);
    const float2 R2 = make_float2(r, r);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            float2 v;
            if (PACKED) v = __fadd2_rn(a[j], R2);
            else v = make_float2(__fadd_rn(a[j].x, r), __fadd_rn(a[j].y, r));
            v.x = __fmul_rn(v.x, m);
            v.y = __fmul_rn(v.y, m);
            a[j] = v;
            if (PACKED) { if (j & 1) S1 = __fadd2_rn(S1, v); else S0 = __fadd2_rn(S0, v); }
            else { if (j & 1) { S1.x = __fadd_rn(S1.x, v.x); S1.y = __fadd_rn(S1.y, v.y); } else { S0.x = __fadd_rn(S0.x, v.x); S0.y = __fadd_rn(S0.y, v.y); } }
        }
    }
    float acc = S0.x + S0.y + S1.x + S1.y;
#pragma unroll
    for (int j = 0; j < 16; j++) acc += a[j].x + a[j].y;
    if (acc == 123.456f) out[threadIdx.x] = acc; // keep the work alive
}

// ---------------------------------------------------------------------------------------
// Bit packing.  hap: L*N chars; G: L rows of `wps` words, bit (n&31) of word n>>5 = hap[s][n]=='1'.
// One thread per output word; a warp reads 1 KiB of consecutive chars.
// padbit: value of the slots past N in a partial last word (the painter's phantom haplotypes, see paint_jobs)
__global__ void pack_snp_major_kernel(const unsigned char *__restrict__ hap, int N, int L, uint32_t *__restrict__ G,
                                      int wps, int padbit)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)L * wps;
    if (idx >= total) return;
    const int s = (int)(idx / wps), w = (int)(idx % wps);
    const int n0 = w * 32;
    uint32_t bits = 0;
    if (n0 < N) {
        const unsigned char *p = hap + (size_t)s * N + n0;
        const int cnt = min(32, N - n0);
        if (cnt == 32 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint4 v = __ldg(q + h);
                uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // chars '0'=0x30, '1'=0x31: gather bit 0 of each byte
                    uint32_t b = ws[j] & 0x01010101u;
                    b = (b | (b >> 7) | (b >> 14) | (b >> 21)) & 0xFu;
                    bits |= b << (h * 16 + j * 4);
                }
            }
        } else {
            for (int j = 0; j < cnt; j++) bits |= (uint32_t)(p[j] == '1') << j;
        }
        if (padbit && cnt < 32) bits |= ~0u << cnt;
    }
    G[idx] = bits;
}

// GT[k][s>>5] bit (s&31) = hap[s][k].  Tile: 32 SNP-groups (1024 SNPs) x 32 haplotypes (one G word column).
// blockDim = (32, 32): threadIdx.y = SNP group within the tile, threadIdx.x = SNP within the group.
__global__ void transpose_bits_kernel(const uint32_t *__restrict__ G, int wps, int N, int L,
                                      uint32_t *__restrict__ GT, int lw)
{
    __shared__ uint32_t tile[32][33];
    const int wcol = blockIdx.x;                // G word column: haplotypes 32*wcol .. +31
    const int sg0 = blockIdx.y * 32;            // first SNP group of the tile
    const int sg = sg0 + threadIdx.y;
    const int s = sg * 32 + threadIdx.x;
    uint32_t word = (s < L) ? G[(size_t)s * wps + wcol] : 0u;
#pragma unroll
    for (int b = 0; b < 32; b++) {
        uint32_t bal = __ballot_sync(0xffffffffu, (word >> b) & 1u);
        if (threadIdx.x == b) tile[b][threadIdx.y] = bal; // hap b of the column, SNP group threadIdx.y
    }
    __syncthreads();
    // write: row = haplotype (threadIdx.y), consecutive threads -> consecutive SNP groups
    const int k = wcol * 32 + threadIdx.y;
    const int sgw = sg0 + threadIdx.x;
    if (k < N && sgw < lw) GT[(size_t)k * lw + sgw] = tile[threadIdx.y][threadIdx.x];
}

// ---------------------------------------------------------------------------------------
// Site enumeration.  der(k) = {0} U {s in [1,L-2] : H[s][k]=1} U {L-1}  (fast_painting.cpp:52-131)
__device__ __forceinline__ uint32_t interior_mask(uint32_t word, int sw, int L)
{
    // keep only SNPs in [1, L-2]
    const int s0 = sw * 32;
    if (s0 == 0) word &= ~1u;
    const int hi = L - 1 - s0; // first excluded bit index within this word
    if (hi <= 0) return 0u;
    if (hi < 32) word &= (1u << hi) - 1u;
    return word;
}

// one warp per target: counts[kk] = D_k
__global__ void count_sites_kernel(const uint32_t *__restrict__ GT, int lw, int L, int k0, int nt,
                                   int *__restrict__ counts)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nt) return;
    const uint32_t *row = GT + (size_t)(k0 + warp) * lw;
    int c = 0;
    for (int w = lane; w < lw; w += 32) c += __popc(interior_mask(row[w], w, L));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[warp] = c + 2;
}

// single block exclusive scan of nt counts into 64-bit offsets (off[nt] = total)
__global__ void scan_counts_kernel(const int *__restrict__ counts, int nt, long long *__restrict__ off)
{
    __shared__ long long part[1024];
    const int t = threadIdx.x, T = blockDim.x;
    const int per = (nt + T - 1) / T;
    const int b = t * per, e = min(nt, b + per);
    long long s = 0;
    for (int i = b; i < e; i++) s += counts[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long acc = 0;
        for (int i = 0; i < T; i++) { long long v = part[i]; part[i] = acc; acc += v; }
        off[nt] = acc;
    }
    __syncthreads();
    long long acc = part[t];
    for (int i = b; i < e; i++) { off[i] = acc; acc += counts[i]; }
}

template <typename ENT>
__global__ void fill_sites_kernel(const uint32_t *__restrict__ GT, int lw, int L, int k0, int nt,
                                  const long long *__restrict__ off, ENT *__restrict__ ent)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nt) return;
    const uint32_t *row = GT + (size_t)(k0 + warp) * lw;
    ENT *e = ent + off[warp];
    if (lane == 0) e[0].site = 0;
    int base = 1;
    for (int w0 = 0; w0 < lw; w0 += 32) {
        const int w = w0 + lane;
        uint32_t word = (w < lw) ? interior_mask(row[w], w, L) : 0u;
        const int c = __popc(word);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int pos = base + incl - c;
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            e[pos++].site = w * 32 + b;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) e[base].site = L - 1;
}

// The same with one CTA of NW warps per target: warp s counts, then fills, the contiguous share s of the row's words.
template <typename ENT, int NW>
__global__ void __launch_bounds__(32 * NW) fill_sites_cta_kernel(const uint32_t *__restrict__ GT, int lw, int L, int k0, int nt,
                                                                 const long long *__restrict__ off, ENT *__restrict__ ent)
{
    __shared__ int s_cnt[NW];
    const int kk = blockIdx.x, lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
    if (kk >= nt) return;
    const uint32_t *row = GT + (size_t)(k0 + kk) * lw;
    ENT *e = ent + off[kk];
    const int groups = (lw + 31) >> 5; // groups of 32 words
    const int w_lo = (int)(((long long)groups * seg / NW) << 5), w_hi = min(lw, (int)(((long long)groups * (seg + 1) / NW) << 5));
    int c = 0;
    for (int w = w_lo + lane; w < w_hi; w += 32) c += __popc(interior_mask(row[w], w, L));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) s_cnt[seg] = c;
    __syncthreads();
    int base = 1;
    for (int sgm = 0; sgm < seg; sgm++) base += s_cnt[sgm];
    if (threadIdx.x == 0) e[0].site = 0;
    for (int w0 = w_lo; w0 < w_hi; w0 += 32) {
        const int w = w0 + lane;
        uint32_t word = (w < w_hi) ? interior_mask(row[w], w, L) : 0u;
        const int cw = __popc(word);
        int incl = cw;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int pos = base + incl - cw;
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            e[pos++].site = w * 32 + b;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (seg == NW - 1 && lane == 0) e[base].site = L - 1;
}

// ---------------------------------------------------------------------------------------
// Window boundary sites (fast_painting.cpp:60-69,98-107,150):
//   begin[0] = 0;  begin[w] = last visited site < wb[w];  end[w] = first visited site >= wb[w+1];  end[W-1] = L-1.
// One thread per (target, window).  ia/ib are indices into the target's site list.
template <typename ENT>
__global__ void boundaries_kernel(const ENT *__restrict__ ent, const long long *__restrict__ off, int nt, int W,
                                  const int *__restrict__ wb, int *__restrict__ ia, int *__restrict__ ib,
                                  int *__restrict__ site_begin, int *__restrict__ site_end)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nt * W) return;
    const int kk = idx / W, w = idx % W;
    const ENT *e = ent + off[kk];
    const int D = (int)(off[kk + 1] - off[kk]);
    // first index with site >= x
    auto lower = [&](int x) {
        int lo = 0, hi = D;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (e[mid].site < x) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    int a = (w == 0) ? 0 : lower(wb[w]) - 1;
    int b = (w == W - 1) ? D - 1 : lower(wb[w + 1]);
    if (a < 0) a = 0;
    if (b > D - 1) b = D - 1;
    ia[idx] = a;
    ib[idx] = b;
    site_begin[idx] = e[a].site;
    site_end[idx] = e[b].site;
}

// ---------------------------------------------------------------------------------------
// Recombination tables (fast_painting.cpp:72-81,112-121,132-141 and the c_i used at :260,351,455,555):
//   x_i = sum r[der_i .. der_{i+1}-1]  (x_m = r[L-1]);  rho_i = 1-exp(-x_i), capped at 0.99;
//   nor_i = -x_i + log(ntheta)  (cap: log(0.01)+log(ntheta));  c_i = rho_i / ((1-rho_i)(N-1)).
// Short gaps are summed in the reference's order; long ones use a double-double prefix of r.
// One warp per target; the running sum of nor gives the log-scale bases at the boundary sites:
//   lsA[w] = sum_{j<ia[w]} nor_j                      (forward, :279-280)
//   lsB[w] = log(N-1) - D*log(ntheta) + sum_{j>ib[w]} nor_j   (backward, :399,471-472)
struct TableConsts {
    double log_ntheta, log_small, Nm1;
};

// SEQ_GAP: gaps of at most this many SNPs are summed in the reference's order (fp64 verification mode: 16);
// the fp32 path takes every gap from the double-double prefix (SEQ_GAP = 0), which is exact to ~1e-30.
template <typename ENT, int SEQ_GAP>
__global__ void tables_kernel(ENT *__restrict__ ent, const long long *__restrict__ off, int nt, int L, int W,
                              const double *__restrict__ r, const double *__restrict__ Phi,
                              const double *__restrict__ Plo, TableConsts tc, const int *__restrict__ ia,
                              const int *__restrict__ ib, double *__restrict__ lsA, double *__restrict__ lsB,
                              double *__restrict__ nor_out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nt) return;
    ENT *e = ent + off[warp];
    double *no = nor_out ? nor_out + off[warp] : nullptr; // per-entry nor_i, kept for the window repaint
    const int D = (int)(off[warp + 1] - off[warp]);
    const int *ja = ia + (size_t)warp * W, *jb = ib + (size_t)warp * W;
    double *oa = lsA + (size_t)warp * W, *ob = lsB + (size_t)warp * W;
    int qa = 0, qb = 0;
    // windows whose forward boundary is index 0 have an empty sum
    while (qa < W && ja[qa] == 0) { if (lane == 0) oa[qa] = 0.0; qa++; }
    int nexta = qa < W ? ja[qa] - 1 : 0x7fffffff, nextb = jb[0];
    double carry = 0.0;
    constexpr int TILES = 4; // independent 32-entry tiles per iteration (latency hiding: one warp per target)
    for (int i0 = 0; i0 < D; i0 += 32 * TILES) {
        double nor[TILES];
#pragma unroll
        for (int u = 0; u < TILES; u++) {
            const int i = i0 + 32 * u + lane;
            nor[u] = 0.0;
            if (i < D) {
                const int a = e[i].site;
                const int b = (i + 1 < D) ? e[i + 1].site : L;
                double x;
                if (SEQ_GAP > 0 && b - a <= SEQ_GAP) {
                    x = r[a];
                    for (int s = a + 1; s < b; s++) x += r[s];
                } else {
                    x = (Phi[b] - Phi[a]) + (Plo[b] - Plo[a]);
                }
                nor[u] = -x + tc.log_ntheta;
                double rho = 1.0 - exp(-x);
                if (rho > 0.99) {
                    rho = 0.99;
                    nor[u] = tc.log_small + tc.log_ntheta;
                }
                e[i].set_c(rho / ((1.0 - rho) * tc.Nm1));
                if (no) no[i] = nor[u];
            }
        }
#pragma unroll
        for (int u = 0; u < TILES; u++) {
            const int t0 = i0 + 32 * u;
            double incl = nor[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            incl += carry; // cum_i = sum_{j<=i} nor_j
            // forward bases need cum_{ia-1}; backward ones cum_{ib} (finalised below)
            while (nexta < t0 + 32) {
                double v = __shfl_sync(0xffffffffu, incl, nexta - t0);
                if (lane == 0) oa[qa] = v;
                qa++;
                nexta = qa < W ? ja[qa] - 1 : 0x7fffffff;
            }
            while (nextb < t0 + 32) {
                double v = __shfl_sync(0xffffffffu, incl, nextb - t0);
                if (lane == 0) ob[qb] = v; // provisional: cum_{ib}
                qb++;
                nextb = qb < W ? jb[qb] : 0x7fffffff;
            }
            carry = __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    // carry == cum_{m}; lsB = norm + (cum_m - cum_ib)
    const double norm = log(tc.Nm1) - (double)D * tc.log_ntheta;
    __syncwarp();
    for (int w = lane; w < W; w += 32) ob[w] = norm + (carry - ob[w]);
}

// The same tables with one CTA (NW warps) per target: warp s owns the contiguous segment [D*s/NW, D*(s+1)/NW) rounded to
// 32-entry tiles, scans it locally, and the segments' totals are combined through shared memory afterwards.  Eight
// times the parallelism of the warp-per-target kernel (one warp per target left the kernel latency-bound: 0.22 ms at
// config 2), at the price of a different association of the fp64 sums (relative 1e-16): used by the fp32 painter; the
// fp64 verification mode keeps the sequential order of tables_kernel.
template <typename ENT, int NW>
__global__ void __launch_bounds__(32 * NW) tables_cta_kernel(ENT *__restrict__ ent, const long long *__restrict__ off, int nt, int L,
                                                             int W, const double *__restrict__ Phi, const double *__restrict__ Plo,
                                                             TableConsts tc, const int *__restrict__ ia, const int *__restrict__ ib,
                                                             double *__restrict__ lsA, double *__restrict__ lsB,
                                                             double *__restrict__ nor_out)
{
    __shared__ double s_tot[NW];
    const int kk = blockIdx.x, lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
    if (kk >= nt) return;
    ENT *e = ent + off[kk];
    double *no = nor_out ? nor_out + off[kk] : nullptr;
    const int D = (int)(off[kk + 1] - off[kk]);
    const int *ja = ia + (size_t)kk * W, *jb = ib + (size_t)kk * W;
    double *oa = lsA + (size_t)kk * W, *ob = lsB + (size_t)kk * W;
    const int tiles = (D + 31) >> 5;
    auto seg_begin = [&](int sgm) { return (int)(((long long)tiles * sgm / NW) << 5); }; // first entry of segment sgm
    const int i_lo = seg_begin(seg), i_hi = min(D, seg_begin(seg + 1));
    // boundaries inside this segment, in window order: forward ones need cum_{ia-1} (index ia-1), backward ones cum_{ib}
    int qa = 0, qb = 0;
    while (qa < W && ja[qa] - 1 < i_lo) qa++; // includes ia == 0 (empty sum, written below)
    while (qb < W && jb[qb] < i_lo) qb++;
    int nexta = qa < W ? ja[qa] - 1 : 0x7fffffff, nextb = qb < W ? jb[qb] : 0x7fffffff;
    double carry = 0.0;
    constexpr int TILES = 2;
    for (int i0 = i_lo; i0 < i_hi; i0 += 32 * TILES) {
        double nor[TILES];
#pragma unroll
        for (int u = 0; u < TILES; u++) {
            const int i = i0 + 32 * u + lane;
            nor[u] = 0.0;
            if (i < i_hi) {
                const int a = e[i].site;
                const int b = (i + 1 < D) ? e[i + 1].site : L;
                const double x = (Phi[b] - Phi[a]) + (Plo[b] - Plo[a]);
                nor[u] = -x + tc.log_ntheta;
                double rho = 1.0 - exp(-x);
                if (rho > 0.99) {
                    rho = 0.99;
                    nor[u] = tc.log_small + tc.log_ntheta;
                }
                e[i].set_c(rho / ((1.0 - rho) * tc.Nm1));
                if (no) no[i] = nor[u];
            }
        }
#pragma unroll
        for (int u = 0; u < TILES; u++) {
            const int t0 = i0 + 32 * u;
            double incl = nor[u];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                double v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            incl += carry; // sum of nor over [i_lo, i]
            while (nexta < min(t0 + 32, i_hi)) {
                double v = __shfl_sync(0xffffffffu, incl, nexta - t0);
                if (lane == 0) oa[qa] = v; // provisional: without the earlier segments
                qa++;
                nexta = qa < W ? ja[qa] - 1 : 0x7fffffff;
            }
            while (nextb < min(t0 + 32, i_hi)) {
                double v = __shfl_sync(0xffffffffu, incl, nextb - t0);
                if (lane == 0) ob[qb] = v;
                qb++;
                nextb = qb < W ? jb[qb] : 0x7fffffff;
            }
            carry = __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    if (lane == 0) s_tot[seg] = carry;
    __syncthreads();
    // segment of an entry index, and the sum of the segments before it
    auto prefix_before = [&](int idx) {
        double p = 0.0;
        for (int sgm = 0; sgm < NW; sgm++)
            if (seg_begin(sgm + 1) <= idx) p += s_tot[sgm];
        return p;
    };
    double total = 0.0;
    for (int sgm = 0; sgm < NW; sgm++) total += s_tot[sgm];
    const double norm = log(tc.Nm1) - (double)D * tc.log_ntheta;
    for (int w = threadIdx.x; w < W; w += blockDim.x) {
        oa[w] = (ja[w] == 0) ? 0.0 : oa[w] + prefix_before(ja[w] - 1);
        ob[w] = norm + (total - (ob[w] + prefix_before(jb[w])));
    }
}

// ---------------------------------------------------------------------------------------
// Per-step table entries.
struct __align__(8) EntF { // fp32 mode
    int site;
    float c;
    __device__ __forceinline__ void set_c(double v) { c = (float)v; }
};
struct __align__(16) EntD { // fp64 verification mode
    int site;
    int pad;
    double c;
    __device__ __forceinline__ void set_c(double v) { c = v; }
};

template <typename T> struct Real;
template <> struct Real<float> {
    using Ent = EntF;
    using V2 = float2;
    static __device__ __forceinline__ V2 add2(V2 a, V2 b) { return __fadd2_rn(a, b); }
    static __device__ __forceinline__ V2 mk(float x, float y) { return make_float2(x, y); }
};
template <> struct Real<double> {
    using Ent = EntD;
    using V2 = double2;
    static __device__ __forceinline__ V2 add2(V2 a, V2 b) { return make_double2(a.x + b.x, a.y + b.y); }
    static __device__ __forceinline__ V2 mk(double x, double y) { return make_double2(x, y); }
};

template <typename T> struct PaintConsts { // fast_painting.hpp:26-39, already in the kernel's arithmetic type
    T tau;        // theta_ratio + 1.0: the mismatch multiplier
    T prior_n;    // ntheta/(N-1)
    T ntheta;
    T inv_ntheta;
    T lower, upper; // rescaling band (1e-10, 1e10)
};

#ifndef RP_SPIN_NS
#define RP_SPIN_NS 64 // poll interval of a team waiting for a parked chain
#endif

struct PaintParams {
    const uint32_t *G;     // SNP-major bits
    int wps;               // words per SNP row
    int N, L, W;
    int nfw;               // 32-haplotype words per row, a partial last word included: ceil(N / 32)
    int tailn;             // N % 32
    int padbit;            // genotype bit of the phantom slots of a partial last word: 0 if tau <= 1, 1 if tau > 1
    int k0, nt;            // targets [k0, k0+nt); each is one forward and one backward job
    const void *ent;       // EntF[] or EntD[], padded by 4 valid entries at both ends
    const long long *off;  // [nt+1]
    const int *ia, *ib;    // [nt][W] boundary indices into the site list
    const double *lsA, *lsB; // [nt][W] log-scale bases
    float *alpha, *beta;   // [nt][W][N]
    float *ls_alpha, *ls_beta; // [nt][W]
    int *queue;            // job counters: [0] forward, [1] backward pops; [2], [3] pushes onto the ready queues
    // Chain segments (load balance): a job is (segment, chain), segment-major; a chain's state is parked in HBM between
    // its segments, so whichever team is free continues it.  nseg == 1: a job is a whole chain, nothing is parked.
    int nseg;
    int *segready;         // [2][nt*(nseg-1)] ready queue of parked chains (job+1 once pushed), zeroed before the launch
    char *segstate;        // [2][nt] parked states of segstride bytes: team vector, tail elements, scalars
    size_t segstride;
    double *scratch;       // fp64 mode only: [gridDim.x][N] staging rows for backward stepping stones
    int hshift;            // h: headroom bits of the single-REDUX sum for multipliers > 1 (theta > 1/2), else 0
    int k1c, k2c;          // (283-h)<<23 and (h-29)<<23: exponent arithmetic of the fixed-point unit 2^(E-29+h)
    int xlo[2], xhi[2];    // per direction: float bits of band_lower/chk resp. band_upper/chk (a hair inside)
    PaintConsts<float> cf;
    PaintConsts<double> cd;
};

template <typename T> __device__ __forceinline__ const PaintConsts<T> &paint_consts(const PaintParams &P);
template <> __device__ __forceinline__ const PaintConsts<float> &paint_consts<float>(const PaintParams &P) { return P.cf; }
template <> __device__ __forceinline__ const PaintConsts<double> &paint_consts<double>(const PaintParams &P) { return P.cd; }

// Launch geometry.  The team of T = blockDim.x threads owns a job's N-vector: thread t owns the WPT adjacent
// genotype words t*WPT + j; a word's 32 haplotypes sit in 32 registers, rotated by rot = k & 31 so that the target
// is slot 0 of its word.  Register budget: 32*WPT state registers (x2 for fp64) + ~30.
// DENSE selects a tighter register cap for multi-warp teams of <= 160 threads with two words per thread
// (96 registers -> 4 CTAs of 160 threads, or 7 of 96, per SM instead of 3 / 5).
template <typename T, int WPT, bool MULTI, bool DENSE = false>
struct PaintCfg {
    static constexpr int kStateRegs = 32 * WPT * (int)(sizeof(T) / 4);
    // register caps: multi-warp teams 80 (fp32, one word per thread: <=384 threads, 2 CTAs/SM), 128 (fp32, two
    // words: <=512 threads) or 168 (fp64: <=384 threads); single-warp teams 128 or 255
    static constexpr bool kF64 = sizeof(T) == 8;
    static constexpr int kMaxThreads = MULTI ? (DENSE ? 160 : ((kStateRegs <= 32 || kF64) ? 384 : 512)) : 32;
    static constexpr int kMinBlocks = MULTI ? (DENSE ? 4 : (kStateRegs <= 32 ? 2 : 1)) : (kStateRegs <= 32 ? 16 : 8);
};

template <int WPT> __device__ __forceinline__ void load_words(uint32_t (&w)[WPT], const char *p)
{
    if (WPT == 1) {
        w[0] = *reinterpret_cast<const uint32_t *>(p);
    } else if (WPT == 2) {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        w[0] = v.x;
        w[WPT - 1] = v.y;
    } else {
#pragma unroll
        for (int j = 0; j < WPT; j++) w[j] = reinterpret_cast<const uint32_t *>(p)[j];
    }
}

#ifndef RP_SUM_FLOOR_LOG2
#define RP_SUM_FLOOR_LOG2 23 // least fixed-point units the single-REDUX team sum must carry (else: butterfly on the rare path).
                            // Measured at config 2 (scripts/acc_check.py, sweep of 20/23/25/27): kernel 2.20 / 2.21 / 2.25 / 2.31 ms,
                            // max relative error of the stepping stones 4.1e-6 / 8.6e-7 / 8.4e-7 / 7.6e-7
#endif
#ifndef RP_TMA_ROWS
#define RP_TMA_ROWS 0 // experiment (multi-warp fp32 teams): genotype rows through a cp.async.bulk + mbarrier ring in shared memory,
                      // issued by one thread four steps ahead, instead of one coalesced LDG per thread one step ahead.
                      // Measured on B200 (profiles/r02): see DESIGN.md section 4; off by default.
#endif
#ifndef RP_PF
#define RP_PF 1 // L1 prefetch hints for the genotype rows / site tables of later steps
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// ---- bulk-async (TMA, 1-D) row ring helpers (RP_TMA_ROWS) ----
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void opaque(float &x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void opaque(double &x) { asm volatile("" : "+d"(x)); }

// CLUSTER: the team is a thread-block cluster (N beyond one CTA's reach: up to 16 x 512 threads x 2 words).  CTA r
// owns the words of team threads r*T .. r*T+T-1; the per-step sum adds the CTAs' sums through distributed shared
// memory (one cluster barrier per step, partial sums double-buffered), the job index is fetched by CTA 0.
template <typename T, int WPT, bool MULTI, int DIR, bool CLUSTER = false>
__device__ __forceinline__ void paint_jobs(const PaintParams &P, int *s_job, T (*s_part)[32], float *s_cpart = nullptr)
{
    static_assert(!CLUSTER || (MULTI && sizeof(T) == 4), "cluster teams are multi-warp fp32 teams");
    namespace cg = cooperative_groups;
    const int crank = CLUSTER ? (int)cg::this_cluster().block_rank() : 0;
    const int csize = CLUSTER ? (int)cg::this_cluster().num_blocks() : 1;
    using RT = Real<T>;
    using V2 = typename RT::V2;
    using Ent = typename RT::Ent;
    constexpr int ES = DIR ? -1 : 1; // direction of travel through the site list

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int TT = blockDim.x, TW = TT >> 5;
    const int gt = crank * TT + t; // index of this thread within the team
    const PaintConsts<T> &K = paint_consts<T>(P);
    // loop constants live in registers (opaque to the compiler, which would otherwise re-read the constant bank
    // every step)
    // multi-warp / fp64 teams: the rare path is entered on S outside [band_lo, band_hi], the rescaling band divided by
    // chk and widened by a hair; the handler repeats the reference's exact test on B = chk*S
    T tau = K.tau, band_lo = K.lower / (DIR ? K.ntheta : (T)1) * (T)1.000002, band_hi = K.upper / (DIR ? K.ntheta : (T)1) * (T)0.999998;
    opaque(tau);
    float Nf = (float)P.N; // every element gains R per step: S_new <= S + N*R bounds the coming sum
    opaque(Nf);
    opaque(band_lo);
    opaque(band_hi);
    const Ent *ents = reinterpret_cast<const Ent *>(P.ent);
    T *scratch = reinterpret_cast<T *>(P.scratch) + (size_t)blockIdx.x * P.N; // dereferenced in fp64 mode only
    int *queue = P.queue + DIR;
    unsigned rowbytes = (unsigned)P.wps * 4u;
    asm volatile("" : "+r"(rowbytes));

    // Long-lived per-thread facts are kept as multipliers / opaque registers rather than predicates: R2P
    // rewrites P0-P6 four times per word, so a predicate cannot survive a step, and ptxas would otherwise
    // rematerialise pointers and masks from the constant bank every iteration.
    bool valid[WPT];
    T vmul[WPT];           // 1 for words this thread owns, 0 otherwise: R*vmul keeps an unused slot at exactly 0
#pragma unroll
    for (int j = 0; j < WPT; j++) {
        valid[j] = (gt * WPT + j) < P.nfw;
        vmul[j] = valid[j] ? (T)1 : (T)0;
        opaque(vmul[j]);
    }
    // this thread's WPT adjacent words within a row: one (vector) load per visited site, 4*WPT bytes per lane,
    // coalesced across the warp.  Rows are padded to 16 bytes and a word past the last full one is either the tail
    // word or padding, so threads beyond the row's end read word 0 and ignore it.
    const char *gthr = reinterpret_cast<const char *>(P.G + ((gt * WPT + WPT - 1) < P.wps ? gt * WPT : 0));
    asm volatile("" : "+l"(gthr));
    // The last word of a row may be partial (P.tailn = N % 32 haplotypes).  Its owner treats it as a full word: the
    // missing slots are phantom haplotypes whose genotype bit the packer chose so that they always take the SMALLER of
    // the two multipliers: bit 0 (mismatching wherever the target is derived) when tau <= 1, bit 1 (never mismatching)
    // when tau > 1, i.e. theta > 1/2.  All phantoms carry the same value, which every thread tracks in one scalar (xph,
    // the same two roundings per step as the registers, so it is bit-identical to them) and the owner subtracts
    // nph * xph from its sum.  A phantom is then never larger than any real element (same R, never a larger
    // multiplier), so xph <= S/N and the subtraction costs no precision.  No extra load, predicate or branch per step.
    T nph = (T)0;
#pragma unroll
    for (int j = 0; j < WPT; j++)
        if (P.tailn > 0 && gt * WPT + j == P.nfw - 1) nph = (T)(32 - P.tailn);
    opaque(nph);

    const T chk = DIR ? K.ntheta : (T)1;        // the band is tested on chk*S  (B = ntheta*G backward)
    const T resc_R = DIR ? K.inv_ntheta : (T)1; // R after a rescale, before *c_i

    constexpr bool kTma = RP_TMA_ROWS && MULTI && !CLUSTER && sizeof(T) == 4;
    constexpr int kRing = 4;
    __shared__ __align__(16) uint32_t s_rows[kTma ? kRing : 1][kTma ? 1024 : 4]; // a team of <= 512 threads x 2 words
    __shared__ __align__(8) uint64_t s_mbar[kRing];
    unsigned ring_phase = 0; // bit s: parity the next wait on slot s expects
    if (kTma) {
        if (t == 0)
            for (int sl = 0; sl < kRing; sl++) mbar_init(&s_mbar[sl], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    // (thread 0) request the genotype row of `site` into ring slot sl
    auto ring_issue = [&](int site, int sl) {
        mbar_expect_tx(&s_mbar[sl], rowbytes);
        bulk_g2s(&s_rows[sl][0], P.G + (size_t)site * P.wps, rowbytes, &s_mbar[sl]);
    };

    const int nseg = CLUSTER ? 1 : P.nseg;
    for (;;) {
        int kk = 0;
        if (CLUSTER) {
            cg::cluster_group cl = cg::this_cluster();
            cl.sync(); // every CTA is done with the previous job (and with s_job)
            if (gt == 0) {
                const int v = atomicAdd(queue, 1);
                for (int r = 0; r < csize; r++) *cl.map_shared_rank(s_job, r) = v;
            }
            cl.sync();
            kk = *s_job;
        } else {
            // Ready queue.  Pop position `pos` of this direction: the first nt positions are the chains' first segments
            // (ready from the start); position nt+i is the i-th chain segment that gets parked and pushed (`ready[i]`
            // = job+1, 0 while nobody has pushed it yet).  Every lane of the polling warp polls, so the warp stays
            // converged (a lone polling lane leaves the warp split for the whole job: 2x slower).
            const int njobs = P.nt * nseg;
            if (!MULTI || warp == 0) {
                kk = 0;
                if (lane == 0) kk = atomicAdd(queue, 1);
                kk = __shfl_sync(0xffffffffu, kk, 0);
                if (nseg > 1 && kk >= P.nt && kk < njobs) {
                    const volatile int *slot = P.segready + (size_t)DIR * (njobs - P.nt) + (kk - P.nt);
                    int v; // the loop condition is a warp vote: all lanes leave together
                    do {
                        v = *slot;
                        if (v == 0) __nanosleep(RP_SPIN_NS);
                    } while (!__all_sync(0xffffffffu, v != 0));
                    __threadfence(); // the parked state was written before the push
                    kk = __shfl_sync(0xffffffffu, v - 1, 0); // provably warp-uniform
                }
            }
            if (MULTI) {
                __syncthreads(); // everyone is done with the previous job (and has read s_job)
                if (t == 0) *s_job = kk;
                __syncthreads();
                kk = *s_job;
            }
            if (kk >= njobs) break;
        }
        // job = segment * nt + chain.  A chain is cut into nseg segments; between segments its state is parked in HBM
        // and the chain is pushed onto the ready queue, so whichever team is free continues it: teams on lightly
        // loaded SM sub-partitions get through more segments, and all chains finish together.
        int seg = 0;
        if (CLUSTER) {
            if (kk >= P.nt) break;
        } else if (nseg > 1) {
            seg = kk / P.nt;
            kk -= seg * P.nt;
        }
        const int k = P.k0 + kk;
        const long long base = P.off[kk];
        const int m = (int)(P.off[kk + 1] - base) - 1;
        // steps [pbeg, pend) of 0..m belong to this segment (cuts at even steps: the pipeline alternates two register sets)
        int pbeg = 0, pend = m + 1;
        if (!CLUSTER && nseg > 1) {
            if (seg > 0) pbeg = (int)(((long long)(m + 1) * seg / nseg) & ~1LL);
            if (seg < nseg - 1) pend = (int)(((long long)(m + 1) * (seg + 1) / nseg) & ~1LL);
        }
        const Ent *pe = ents + base + (DIR ? m : 0); // entry of step p is pe[p*ES]
        const int rot = k & 31, wk = k >> 5;
        bool own[WPT];
#pragma unroll
        for (int j = 0; j < WPT; j++) own[j] = (gt * WPT + j == wk);
        T ownmul[WPT]; // 0 on the thread/word holding the target (slot 0 after rotation), else 1
#pragma unroll
        for (int j = 0; j < WPT; j++) { ownmul[j] = own[j] ? (T)0 : (T)1; opaque(ownmul[j]); }

        // boundary bookkeeping, in the order the job meets the windows: forward w = q, backward w = W-1-q
        const int *bidx = (DIR ? P.ib : P.ia) + (size_t)kk * P.W;
        const double *lsb = (DIR ? P.lsB : P.lsA) + (size_t)kk * P.W;
        float *outv = (DIR ? P.beta : P.alpha) + (size_t)kk * P.W * P.N;
        float *outl = (DIR ? P.ls_beta : P.ls_alpha) + (size_t)kk * P.W;
        auto bpos = [&](int qq) -> int { return DIR ? (m - bidx[P.W - 1 - qq]) : bidx[qq]; };
        int q = 0;

        // state
        V2 a[WPT][16];
#pragma unroll
        for (int j = 0; j < WPT; j++)
#pragma unroll
            for (int e = 0; e < 16; e++) a[j][e] = RT::mk((T)0, (T)0);
        T xph = (T)0, mph = (T)1; // phantom value and its multiplier (tau where the target is derived)
        double lsr = 0.0; // log-scale added by rescaling
        // fixed-point scale of the REDUX sum (fp32 single-warp teams), from the exponent of the previous sum
        // tlo / thi: the rare path is taken when the integer (high) part of the fixed-point sum falls outside
        // [tlo, thi].  That is a superset of {B outside the rescaling band} u {fewer than 64 units left}: the sum is
        // hi + d with |d| <= 16, so B < lower implies hi < X + 17 <= 1.5 X + 64 (X = lower/chk in fixed-point units) and
        // B > upper implies hi > Y - 17.  The handler repeats the exact test, so results do not change; what changes is
        // that the branch needs only REDUX -> I2F -> two compares instead of the whole float reconstruction of B.
        float k1 = 0.f, k2 = 0.f, tlo = 0.f, thi = 0.f;
        // The fixed-point unit is 2^(E-29), E the exponent of an upper bound of the coming sum: S_new <= S + N*R (every
        // element gains R and is then multiplied by tau or 1), so the 32 lane sums, rounded to that unit, add up to less
        // than 2^30 in ONE integer REDUX; at least kSumFloor units (2^23: the lanes' roundings are <= 16 units, ~2 typical)
        // are required of the result, else the step goes through the rare path, which redoes the sum with the shuffle
        // butterfly (the sum then lost more than 2^6 of its bound in one step).
        constexpr float kSumFloor = (float)(1u << RP_SUM_FLOOR_LOG2);
        float blo_c = __int_as_float(P.xlo[DIR]), bhi_c = __int_as_float(P.xhi[DIR]); // band edges / chk, a hair inside
        opaque(blo_c);
        opaque(bhi_c);
        auto set_scale_e = [&](int ebs) {
            k1 = __int_as_float(P.k1c - ebs);   // 2^(29 - E)
            k2 = __int_as_float(ebs + P.k2c);   // 2^(E - 29)
            tlo = fmaf(k1, blo_c, kSumFloor + 17.0f); // band edges in fixed-point units: (lower / chk) * k1, (upper / chk) * k1
            thi = fmaf(k1, bhi_c, -17.0f);
        };
        // bound = S + N*R for the step whose additive term is R
        auto set_scale = [&](float Scur, float Rnext) {
            const float bound = fmaxf(fmaf(Rnext, Nf, Scur), 1e-30f);
            set_scale_e(__float_as_int(bound) & 0x7f800000);
        };
        if (!MULTI && sizeof(T) == 4) set_scale(0.0f, DIR ? 1.0f : (float)K.prior_n);

        // x <- (x + R) * (mis ? tau : 1);  returns the team-wide sum.  mis = target derived && reference
        // ancestral; tdm is all-ones when the target is derived at the site (always, except SNP 0 / L-1).
        auto step_local = [&](const uint32_t (&w)[WPT], uint32_t tdm, T R) -> T {
            // partial sums: four independent chains when a thread owns one word, two when it owns two
            // (register budget of the 96-register multi-warp variant)
            constexpr int NACC = (WPT == 1) ? 4 : 2;
            V2 S0, S1, S2, S3;
#pragma unroll
            for (int j = 0; j < WPT; j++) {
                const T Rj = R * vmul[j];
                const V2 R2 = RT::mk(Rj, Rj);
                const uint32_t nb = ~w[j] & tdm;
                const uint32_t mw = __funnelshift_r(nb, nb, rot);
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    V2 v = RT::add2(a[j][e], R2);
                    if (mw & (1u << (2 * e))) v.x *= tau;
                    if (mw & (2u << (2 * e))) v.y *= tau;
                    if (e == 0) v.x *= ownmul[j];
                    a[j][e] = v;
                    const int ch = e % NACC;
                    if (j == 0 && e < NACC) { // the first pairs seed the accumulators
                        if (ch == 0) S0 = v; else if (ch == 1) S1 = v; else if (ch == 2) S2 = v; else S3 = v;
                    } else if (ch == 0) S0 = RT::add2(S0, v);
                    else if (ch == 1) S1 = RT::add2(S1, v);
                    else if (ch == 2) S2 = RT::add2(S2, v);
                    else S3 = RT::add2(S3, v);
                }
            }
            S0 = RT::add2(S0, S1);
            if (NACC == 4) S0 = RT::add2(S0, RT::add2(S2, S3));
            T S = S0.x + S0.y;
            xph = (xph + R) * mph;
            S -= nph * xph;
            return S;
        };
        // team-wide sum: warp butterfly, then (multi-warp teams) one bar.sync and a shared-memory exchange
        auto reduce = [&](T S, int parity, float tlo_e, int &sa_out, bool &rare) -> T {
            if (!MULTI && sizeof(T) == 4) {
                // Single-warp fp32 teams: the 32 lane sums are rounded to the fixed-point unit chosen by set_scale and added
                // by one integer REDUX (order-free; ~30 cycles of dependent latency instead of the ~150 of a 5-level
                // shuffle butterfly).  v4/v5 carried 46 bits in two REDUX with a fixed 2^hshift of headroom over the
                // previous sum; the exact bound S + N*R makes one 32-bit word enough.
                const int sa = __reduce_add_sync(0xffffffffu, __float2int_rn((float)S * k1));
                const float fsa = (float)sa;
                sa_out = sa;       // sa < kSumFloor: handled on the rare path (the step is redone with the butterfly)
                rare = (fsa < tlo_e) || (fsa > thi);
                return (T)(fsa * k2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
            if (MULTI) {
                T *part = s_part[parity];
                if (lane == 0) part[warp] = S;
                __syncthreads();
                if (sizeof(T) == 4) { // rows are zero beyond TW: whole float4s, fixed order in every thread
                    const float4 *p4 = reinterpret_cast<const float4 *>(part);
                    float4 v = p4[0];
                    float acc = (v.x + v.y) + (v.z + v.w);
#pragma unroll
                    for (int i = 1; i < 4; i++) {
                        if (4 * i < TW) {
                            v = p4[i];
                            acc += (v.x + v.y) + (v.z + v.w);
                        }
                    }
                    S = (T)acc;
                } else {
                    S = part[0];
                    for (int ww = 1; ww < TW; ww++) S += part[ww];
                }
                if (CLUSTER) { // add the CTAs' sums, in rank order, through distributed shared memory
                    cg::cluster_group cl = cg::this_cluster();
                    if (t == 0) s_cpart[parity] = (float)S;
                    cl.sync();
                    float acc = *cl.map_shared_rank(s_cpart + parity, 0);
                    for (int r = 1; r < csize; r++) acc += *cl.map_shared_rank(s_cpart + parity, r);
                    S = (T)acc;
                }
            }
            return S;
        };

        // writes the team's vector (x + addR, target forced to 0; or all ones) to o[0..N)
        auto store_vec = [&](auto *o, T addR, bool ones) {
            using O = typename std::remove_pointer<decltype(o)>::type;
#pragma unroll
            for (int j = 0; j < WPT; j++) {
                if (valid[j]) {
                    const int n0 = (gt * WPT + j) * 32;
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        T vx = a[j][e].x + addR, vy = a[j][e].y + addR;
                        if (ones) { vx = (T)1; vy = (T)1; }
                        else if (e == 0 && own[j]) vx = (T)0;
                        const int ix = n0 + ((2 * e + rot) & 31), iy = n0 + ((2 * e + 1 + rot) & 31);
                        if (ix < P.N) o[ix] = (O)vx; // (a partial last word: the phantom slots are not stored)
                        if (iy < P.N) o[iy] = (O)vy;
                    }
                }
            }
        };

        // target's own allele at the first and the last step (SNP 0 / SNP L-1 are visited regardless, :52-59,150)
        const int site_first = pe[0].site, site_last = pe[m * ES].site;
        const uint32_t td_first = ((P.G[(size_t)site_first * P.wps + wk] >> rot) & 1u) ? 0xffffffffu : 0u;
        const uint32_t td_last = ((P.G[(size_t)site_last * P.wps + wk] >> rot) & 1u) ? 0xffffffffu : 0u;

        // ---- pipeline state ---------------------------------------------------------------------------
        // Two register sets (A/B) alternate between consecutive steps, so nothing is moved between steps:
        // while step p computes from set X, the genotype words of step p+1 are loaded into set Y (their site
        // index was loaded during step p-1) together with entry p+2 (site index and c, one vector load).
        // Entries up to 2 past either end of the target's list are read and never used (the table is padded).
        uint32_t wA[WPT], wB[WPT];
        int sA, sB;   // site index whose words go INTO set A / B next
        int fA = 0, fB = 0; // site index whose row is prefetched into L1 at the top of the next step computing from set A / B
        T cA, cB;     // c of the step that computes from set A / B
        const int site_pbeg = pe[pbeg * ES].site; // == site_first unless this job continues a parked chain
        load_words(wA, gthr + (size_t)(unsigned)site_pbeg * rowbytes);
        cA = (T)pe[pbeg * ES].c;
        sB = pe[(pbeg + 1) * ES].site;  // step pbeg+1's words go into set B during step pbeg
        sA = 0;
        cB = (T)pe[(pbeg + 1) * ES].c;
        const Ent *pnx = pe + (pbeg + 2) * ES; // entry p+2 while step p runs
        int s_ahead = 0; // (thread 0) site of step p+kRing+1, loaded a step before its row is requested
        if (kTma) {
            __syncthreads(); // nobody still reads the ring of the previous job
            if (t == 0) {
                for (int d = 1; d < kRing; d++)
                    if (pbeg + d < pend) ring_issue(pe[(pbeg + d) * ES].site, (pbeg + d) & (kRing - 1));
                s_ahead = pe[min(pbeg + kRing, m) * ES].site;
            }
        }

        T R = DIR ? (T)1 : K.prior_n; // step 0 is x = (0 + R0) * m
        uint32_t tdm = td_first;
        mph = (td_first && !P.padbit) ? tau : (T)1;
        int q1 = q;
        bool post = false;

        if (DIR && seg == 0) { // beta at SNP L-1 is all ones, the target included (:413-418, :433-448)
            int qe = q;
            while (qe < P.W && bpos(qe) == 0) qe++;
            for (int qq = q; qq < qe; qq++) {
                const int w = P.W - 1 - qq;
                store_vec(outv + (size_t)w * P.N, (T)0, true);
                if (gt == 0) outl[w] = (float)lsb[w];
            }
            q = qe;
        }
        // pev = the step before which the rare-path handler must run: a stepping-stone store (forward: alpha
        // after step bpos, i.e. before step bpos+1; backward: beta at step bpos), the last step (target allele
        // mask), or the step after a backward store (its finalisation)
        auto next_event = [&]() -> int {
            const int nb = q < P.W ? bpos(q) + (DIR ? 0 : 1) : 0x7fffffff;
            return nb < m ? nb : m;
        };
        int pev = 1; // the handler always runs after step 0 (it switches the allele mask to all-ones)

        // rare path, run between step p and step p+1
        auto handler = [&](int p, T S, T cthis, T Sl, bool lowprec) {
            if (!MULTI && sizeof(T) == 4 && lowprec) { // the fixed-point sum ran out of bits: redo it in floating point
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) Sl += __shfl_xor_sync(0xffffffffu, Sl, o);
                S = Sl;
                R = S * cthis;
                set_scale((float)S, (float)R);
            }
            const T B = chk * S;
            bool rescaled = false;
            if (p == 0) { tdm = 0xffffffffu; mph = P.padbit ? (T)1 : tau; } // steps 1..m-1 visit sites where the target is derived
            if (p != 0 && (B < K.lower || B > K.upper)) { // :334-347, :538-551; no test at the first site
                rescaled = true;
                if (sizeof(T) == 4) { // fp32 state: one reciprocal, then multiplies (<= 1 ulp from the division)
                    const T inv = (T)1 / B;
#pragma unroll
                    for (int j = 0; j < WPT; j++)
#pragma unroll
                        for (int e = 0; e < 16; e++) { a[j][e].x *= inv; a[j][e].y *= inv; }
                    xph *= inv;
                } else { // fp64 verification mode divides, as the reference does
#pragma unroll
                    for (int j = 0; j < WPT; j++)
#pragma unroll
                        for (int e = 0; e < 16; e++) { a[j][e].x /= B; a[j][e].y /= B; }
                    xph /= B;
                }
                lsr += DIR ? (double)fast_log_dev((float)B) : log((double)B);
                R = resc_R * cthis;
                if (!MULTI && sizeof(T) == 4) set_scale((float)resc_R, (float)R); // the state now sums to 1 (forward) or 1/ntheta
            }
            if (DIR && post) { // finalise the backward stepping stone(s) of step p: divide by B if it rescaled
                for (int qq = q; qq < q1; qq++) {
                    const int w = P.W - 1 - qq;
                    float *o = outv + (size_t)w * P.N;
                    if (sizeof(T) == 8 || rescaled) {
                        const T *src = sizeof(T) == 8 ? scratch : reinterpret_cast<const T *>(o);
#pragma unroll
                        for (int j = 0; j < WPT; j++) {
                            if (valid[j]) {
                                const int n0 = (gt * WPT + j) * 32;
                                for (int e = 0; e < 32 && n0 + e < P.N; e++) {
                                    T v = src[n0 + e];
                                    if (rescaled) v /= B;
                                    o[n0 + e] = (float)v;
                                }
                            }
                        }
                    }
                    if (gt == 0) outl[w] = (float)(lsb[w] + lsr);
                }
                q = q1;
                post = false;
            }
            const int pn = p + 1; // events of the coming step
            if (pn > m) { pev = 0x7fffffff; return; }
            pev = next_event();
            if (pev == pn) {
                if (pn == m) { tdm = td_last; mph = (td_last && !P.padbit) ? tau : (T)1; }
                const int bp_now = q < P.W ? bpos(q) : -1;
                if (bp_now == (DIR ? pn : pn - 1)) {
                    q1 = q;
                    while (q1 < P.W && bpos(q1) == bp_now) q1++;
                    if (!DIR) { // alpha after step p, post-rescale (:354-374)
                        for (int qq = q; qq < q1; qq++) {
                            store_vec(outv + (size_t)qq * P.N, (T)0, false);
                            if (gt == 0) outl[qq] = (float)(lsb[qq] + lsr);
                        }
                        q = q1;
                    } else { // beta of step pn is b = g_old + R' before the emission multiply (:481-488)
                        post = true;
                        if (sizeof(T) == 8) store_vec(scratch, R, false);
                        else for (int qq = q; qq < q1; qq++) store_vec(outv + (size_t)(P.W - 1 - qq) * P.N, R, false);
                    }
                }
                pev = post ? pn + 1 : (pn == m ? 0x7fffffff : next_event());
                if (!post && pev == pn) pev = pn + 1; // unreachable guard: never re-fire for the same step
            }
        };

        // parks / resumes the chain's state between segments: the team vector (coalesced, L1-bypassing: the next segment
        // usually runs on another SM), the tail elements and the team-uniform scalars
        auto park_io = [&](bool save) {
            char *park = P.segstate + ((size_t)DIR * P.nt + kk) * P.segstride;
            V2 *pv = reinterpret_cast<V2 *>(park);
            double *pd = reinterpret_cast<double *>(park + (size_t)TT * WPT * 16 * sizeof(V2));
            int *pi = reinterpret_cast<int *>(pd + 4);
            if (save) {
#pragma unroll
                for (int j = 0; j < WPT; j++)
#pragma unroll
                    for (int e = 0; e < 16; e++) __stcg(&pv[(size_t)(j * 16 + e) * TT + t], a[j][e]);
                if (t == 0) {
                    __stcg(&pd[0], lsr);
                    __stcg(&pd[1], (double)R);
                    __stcg(&pd[2], (double)xph);
                    __stcg(&pd[3], (double)mph);
                    __stcg(&pi[0], (int)tdm);
                    __stcg(&pi[1], q);
                    __stcg(&pi[2], q1);
                    __stcg(&pi[3], post ? 1 : 0);
                    __stcg(&pi[4], pev);
                    __stcg(&pi[5], __float_as_int(k1));
                    __stcg(&pi[6], __float_as_int(k2));
                }
                __threadfence();
                if (MULTI) __syncthreads(); else __syncwarp();
                if (t == 0) { // push the chain's next segment onto the ready queue
                    const int tp = atomicAdd(queue + 2, 1);
                    atomicExch(P.segready + (size_t)DIR * (P.nt * (nseg - 1)) + tp, (seg + 1) * P.nt + kk + 1);
                }
                if (!MULTI) __syncwarp(); // reconverge here, provably: without it ptxas guards every REDUX / vote of the hot loop with a BRA.DIV
            } else {
#pragma unroll
                for (int j = 0; j < WPT; j++)
#pragma unroll
                    for (int e = 0; e < 16; e++) a[j][e] = __ldcg(&pv[(size_t)(j * 16 + e) * TT + t]);
                lsr = __ldcg(&pd[0]);
                R = (T)__ldcg(&pd[1]);
                xph = (T)__ldcg(&pd[2]);
                mph = (T)__ldcg(&pd[3]);
                tdm = (uint32_t)__ldcg(&pi[0]);
                q = __ldcg(&pi[1]);
                q1 = __ldcg(&pi[2]);
                post = __ldcg(&pi[3]) != 0;
                pev = __ldcg(&pi[4]);
                set_scale_e(P.k1c - __ldcg(&pi[5])); // k1, k2 and the branch thresholds follow from the parked k1
                (void)pi[6];
            }
        };
        if (!CLUSTER && seg > 0) park_io(false);

        // one step computing from set X while filling set Y
        // EV: the handler is known to be due after this step (a pending event); such steps are peeled out of the plain
        // loop below, so that plain steps carry no event test at all
        auto do_step = [&](auto evc, int p, uint32_t (&wX)[WPT], T &cX, int &sX, int &fX, uint32_t (&wY)[WPT], T &cY, int &sY, int &fY) {
            constexpr int EV = decltype(evc)::value; // 0 plain step, 1 event step, 2 test inside the step (multi-warp teams)
            // loads for later steps first: words of step p+1 into Y, entry p+2 (site -> X's next fill, c -> X's next step)
            if (kTma) {
                if (p + 1 < pend) { // the row of step p+1 was requested kRing-1 steps ago
                    const int sl = (p + 1) & (kRing - 1);
                    mbar_wait(&s_mbar[sl], (ring_phase >> sl) & 1u);
                    ring_phase ^= 1u << sl;
                    load_words(wY, reinterpret_cast<const char *>(&s_rows[sl][0]) + (size_t)gt * WPT * 4);
                }
                if (t == 0) { // slot p & 3 held the row of step p, which every thread took before the barrier of step p-1
                    if (p + kRing < pend) ring_issue(s_ahead, p & (kRing - 1));
                    s_ahead = pe[min(p + kRing + 1, m) * ES].site;
                }
            } else
            load_words(wY, gthr + (size_t)(unsigned)sY * rowbytes);
            // The row of step p+2 is loaded at the top of step p+1 and consumed a step later: one step of lead over an
            // L2 latency that, under load, is about a step long (ncu: 8 % of the warps' time on that scoreboard).  So
            // the site of entry p+3 is fetched here (same table line as entry p+2: an L1 hit) and its row is requested
            // from L1 at the top of step p+1 (fX was filled a step ago with the site of entry p+2): two steps of lead.
            if (RP_PF && !MULTI && WPT == 1) { // (measured: +1.4 % at config 2, nothing or a loss for the wider teams)
                prefetch_l1(gthr + (size_t)(unsigned)fX * rowbytes);
                fY = pnx[ES].site;
            }
            // one (vector) load of entry p+2: its site is needed next step (to fetch the words of step p+2), its c at
            // the end of step p+2, which runs on this same register set
            const Ent e2 = *pnx;
            pnx += ES;
            sX = e2.site;
            T Sl = step_local(wX, tdm, R);
            int sa = 64;
            bool rare = false;
            const bool ev = (EV == 2) && (p + 1 == pev);
            const T S = reduce(Sl, p & 1, tlo, sa, rare);
            const T ccur = cX;
            R = S * ccur;
            cX = (T)e2.c;
            if (!MULTI && sizeof(T) == 4) {
                set_scale((float)S, (float)R);
            } else {
                const T lo_e = ev ? (T)INFINITY : band_lo; // (EV == 2) a pending event folded into the threshold
                rare = (S < lo_e) || (S > band_hi);
            }
            // the rare path is taken by all threads or none (S is the team-wide sum): tell the compiler with a vote, so
            // the branch needs no reconvergence bookkeeping
            if (EV == 1) handler(p, S, ccur, Sl, (float)sa < kSumFloor);
            else if (__builtin_expect(__any_sync(0xffffffffu, rare), 0)) handler(p, S, ccur, Sl, (float)sa < kSumFloor);
        };
        auto step_even = [&](auto evc, int p) { do_step(evc, p, wA, cA, sA, fA, wB, cB, sB, fB); }; // computes from set A
        auto step_odd = [&](auto evc, int p) { do_step(evc, p, wB, cB, sB, fB, wA, cA, sA, fA); };  // computes from set B
        using PlainStep = std::integral_constant<int, 0>;
        using EventStep = std::integral_constant<int, 1>;
        using TestStep = std::integral_constant<int, 2>;

        // steps pbeg..pend-1: even steps compute from set A, odd ones from set B.  The handler announces the step after
        // which it must run next (pev = that step + 1: a stepping-stone store, the last step, ...); the steps up to it
        // run in a plain two-step loop, the event step itself is peeled.
        if (MULTI || sizeof(T) == 8) { // (peeling costs these variants registers: they keep the test inside the step)
            for (int p = pbeg; p < pend; p += 2) {
                step_even(TestStep{}, p);
                if (p + 1 >= pend) break;
                step_odd(TestStep{}, p + 1);
            }
        } else for (int p = pbeg; p < pend;) {
            const bool has_ev = pev <= pend;              // the event step is pev-1
            const int plain_end = has_ev ? pev - 1 : pend; // steps [p, plain_end) are plain
            if ((p & 1) && p < plain_end) { step_odd(PlainStep{}, p); p++; }
            for (; p + 1 < plain_end; p += 2) {
                if (RP_PF && !MULTI && WPT == 1) prefetch_l1(pnx + 32 * ES); // site-table line (16 entries) two lines ahead
                step_even(PlainStep{}, p);
                step_odd(PlainStep{}, p + 1);
            }
            if (p < plain_end) { step_even(PlainStep{}, p); p++; }
            if (has_ev && p < pend) { // p == pev - 1
                if (p & 1) step_odd(EventStep{}, p); else step_even(EventStep{}, p);
                p++;
            }
        }
        if (!CLUSTER && seg != nseg - 1) { // park the chain; whichever team takes job (seg+1, chain) continues it
            park_io(true);
            continue;
        }

        if (!DIR) { // alpha stepping stones at the last visited site
            while (q < P.W) {
                store_vec(outv + (size_t)q * P.N, (T)0, false);
                if (gt == 0) outl[q] = (float)(lsb[q] + lsr);
                q++;
            }
        }
    }
}

template <typename T, int WPT, bool MULTI, bool DENSE = false>
__global__ void __launch_bounds__(PaintCfg<T, WPT, MULTI, DENSE>::kMaxThreads, PaintCfg<T, WPT, MULTI, DENSE>::kMinBlocks)
paint_kernel(const PaintParams P)
{
    __shared__ int s_job;
    __shared__ __align__(16) T s_part[2][32];
    if (MULTI) { // partial-sum rows are read as whole vectors: zero the slots no warp writes
        if (threadIdx.x < 64) (&s_part[0][0])[threadIdx.x] = (T)0;
        __syncthreads();
    }
    // even CTAs walk the site lists forwards (alpha), odd CTAs backwards (beta): two instantiations of one loop
    if (blockIdx.x & 1) paint_jobs<T, WPT, MULTI, 1>(P, &s_job, s_part);
    else paint_jobs<T, WPT, MULTI, 0>(P, &s_job, s_part);
}

// Cluster teams: 512 threads x 2 words per CTA, 2..16 CTAs per team (cluster dims set at launch).
__global__ void __launch_bounds__(512, 1) paint_cluster_kernel(const PaintParams P)
{
    __shared__ int s_job;
    __shared__ __align__(16) float s_part[2][32];
    __shared__ float s_cpart[2];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) (&s_part[0][0])[i] = 0.f; // a CTA of a cluster may be one warp
    __syncthreads();
    const unsigned team = blockIdx.x / cooperative_groups::this_cluster().num_blocks();
    if (team & 1) paint_jobs<float, 2, true, 1, true>(P, &s_job, s_part, s_cpart);
    else paint_jobs<float, 2, true, 0, true>(P, &s_job, s_part, s_cpart);
    cooperative_groups::this_cluster().sync(); // no CTA may exit while another still reads its shared memory
}

// =========================================================================================
// Window repaint ("next" row f1): the consumer side of the stepping stones.
//   repaint_kernel   <- FastPainting::RePaintSection            (src/fast_painting.cpp:620-1092)
//   distance_kernel  <- DistanceMeasure::GetMatrix              (src/anc_builder.cpp:108-207)
// For window w every target k is re-painted between its boundary sites bS_k <= wb[w] and eS_k >= wb[w+1]-1 from
// the stored (alpha, beta) pair; ALL visited sites are kept: the forward sweep writes alpha rows to HBM, the
// backward sweep multiplies them in place into the posterior rows top[k][i][:] = alpha_i * beta_i and finishes the
// per-row log-scales.  The window's entries are the slice ia[k][w]..ib[k][w] of the chunk-level site table; only
// the last entry's recombination term differs (x_m = r[eS], :700-709).
//
// Row layout in HBM (`pitch` = T*WPT*32 + 32 floats for a team of T threads): the painter's register order, transposed
// so that one warp instruction touches contiguous memory.  Word w = t*WPT + j of a row belongs to thread t; its 32
// haplotypes, rotated by k & 31 (haplotype n of target k has slot s = (n - k) & 31), form 8 float4 pieces; piece e of
// (t, j) sits at float4 index (j*8 + e)*T + t, so the T threads' pieces e are adjacent (a 16-byte-per-lane access is
// 512 contiguous bytes per warp; with a thread's 128 bytes contiguous instead, every instruction touched 32 different
// lines and the kernel ran at 2.9 TB/s, bound by the load/store unit's tag rate).  The tail haplotypes (N % 32) are
// stored unrotated at float index T*WPT*32 + lane.  HBM-bound: 3 * 4 * N * (sites in window) bytes per target against
// 6 FP32 ops per element.
struct RepaintParams {
    const uint32_t *G;
    int wps, N, L, W, nfw, tailn;
    int w;                   // window
    int nt;                  // targets 0..nt-1 (all N)
    const void *ent;         // chunk-level table (EntF)
    const long long *off;    // [N+1]
    const double *nor;       // [U] nor_i of the chunk-level table
    const double *r;         // [L]
    const int *ia, *ib;      // [N][W]
    const float *alpha_begin, *beta_end; // [N][N] decoded stepping stones of window w (natural order)
    const float *ls_alpha, *ls_beta;     // [N]
    float *top;              // posterior rows, target k at row offset rowoff[k]; `pitch` floats per row
    int pitch;               // T*WPT*32 + 32 for the launch's team size T (see the row layout below)
    float *ls;               // per-row log-scales, same row indexing
    float2 *scal;            // per row: the additive term R that formed the alpha row and the rescaling factor applied to it
    const long long *rowoff; // [N+1] prefix of (ib-ia+1)
    int *queue;
    PaintConsts<float> cf;
    double log_ntheta, log_small, Nm1;
};

// A value stream requested D rows ahead of its use: D registers that shift by one per row.
template <typename V, int D> struct Ahead {
    V q[D];
    __device__ __forceinline__ V pop_push(const V &nv)
    {
        const V r = q[0];
#pragma unroll
        for (int d = 0; d + 1 < D; d++) q[d] = q[d + 1];
        q[D - 1] = nv;
        return r;
    }
};

extern __shared__ __align__(16) unsigned char rp_dyn_smem[];
// Shared memory of a team in the backward sweep: a ring of 2 checkpoint rows (cp.async from HBM) + CK-1 recomputed
// rows, each threads * WPT * 128 bytes, + one 32-float tail line per row.
template <int CK> struct RepaintSmem {
    static constexpr int kSlots = 2 + (CK - 1);
    static size_t bytes(int threads, int wpt) { return (size_t)kSlots * ((size_t)threads * wpt * 128 + 128); }
};

// Checkpointed window repaint.  The forward sweep keeps only every CK-th alpha row in HBM (in its final place in `top`)
// plus two scalars per row: the additive term R that formed the row and the rescaling factor applied to it (1 if
// none).  The backward sweep walks the window in blocks of CK rows from the end: it takes the block's checkpoint from
// a cp.async ring, recomputes the CK-1 rows behind it into shared memory with the stored scalars -- the same two
// roundings per element as in the forward sweep, hence the same bits -- and then runs the backward recursion over the
// block, writing the posterior rows.  HBM traffic per row: 4N (posterior) + 8N/CK (checkpoint written and read back)
// instead of 12N; the recomputation costs 2 FP32 ops per element on a kernel whose FMA pipe was 12 % busy.
template <int WPT, bool MULTI, int CK>
#ifndef RP_REPAINT_MINB
#define RP_REPAINT_MINB 8
#endif
__global__ void __launch_bounds__(MULTI ? (WPT == 1 ? 512 : 256) : 32, MULTI ? 1 : (WPT == 1 ? RP_REPAINT_MINB : 8)) repaint_kernel(const RepaintParams P)
{
    using T = float;
    using RT = Real<T>;
    using V2 = float2;
    __shared__ int s_job;
    __shared__ __align__(16) T s_part[2][32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, TT = blockDim.x, TW = TT >> 5;
    if (MULTI) {
        if (t < 64) (&s_part[0][0])[t] = 0.f;
        __syncthreads();
    }
    const PaintConsts<T> &K = P.cf;
    const T tau = K.tau;
    const EntF *ents = reinterpret_cast<const EntF *>(P.ent);
    const unsigned rowbytes = (unsigned)P.wps * 4u;
    bool valid[WPT];
    T vmul[WPT];
#pragma unroll
    for (int j = 0; j < WPT; j++) {
        valid[j] = (t * WPT + j) < P.nfw;
        vmul[j] = valid[j] ? 1.f : 0.f;
    }
    const char *gthr = reinterpret_cast<const char *>(P.G + ((t * WPT + WPT - 1) < P.wps ? t * WPT : 0));
    const bool has_tail = P.tailn > 0, tail_warp = has_tail && warp == 0, tail_lane = tail_warp && lane < P.tailn;
    const char *gtail = reinterpret_cast<const char *>(P.G + (has_tail ? P.nfw : 0));
    const uint32_t lanebit = 1u << lane;
    const int N = P.N;

    auto reduce = [&](T S, int parity) -> T {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
        if (MULTI) {
            T *part = s_part[parity];
            if (lane == 0) part[warp] = S;
            __syncthreads();
            const float4 *p4 = reinterpret_cast<const float4 *>(part);
            float4 v = p4[0];
            float acc = (v.x + v.y) + (v.z + v.w);
#pragma unroll
            for (int i = 1; i < 4; i++)
                if (4 * i < TW) { v = p4[i]; acc += (v.x + v.y) + (v.z + v.w); }
            S = acc;
        }
        return S;
    };

    // shared-memory rows: piece e of (thread t, word j) of slot s at float4 index ((s*WPT + j)*8 + e)*TT + t (every thread
    // writes and reads only its own pieces: no barrier); tail elements behind the rows
    constexpr int NSLOT = RepaintSmem<CK>::kSlots;
    float4 *srow = reinterpret_cast<float4 *>(rp_dyn_smem);
    float *stail = reinterpret_cast<float *>(srow + (size_t)NSLOT * WPT * 8 * TT);

    for (;;) {
        int k;
        if (MULTI) {
            __syncthreads();
            if (t == 0) s_job = atomicAdd(P.queue, 1);
            __syncthreads();
            k = s_job;
        } else {
            k = 0;
            if (lane == 0) k = atomicAdd(P.queue, 1);
            k = __shfl_sync(0xffffffffu, k, 0);
        }
        if (k >= P.nt) break;
        const int i0w = P.ia[(size_t)k * P.W + P.w], i1w = P.ib[(size_t)k * P.W + P.w];
        const int m = i1w - i0w;                     // rows 0..m
        const EntF *pe = ents + P.off[k] + i0w;      // entry of row i is pe[i]
        const double *pnor = P.nor + P.off[k] + i0w;
        float *top = P.top + (size_t)P.rowoff[k] * P.pitch;
        const size_t pitch = (size_t)P.pitch;
        const int tail_off = TT * WPT * 32;
        float *lsrow = P.ls + P.rowoff[k];
        float2 *sc = P.scal + P.rowoff[k];           // per row: (R that formed it, rescaling factor applied to it)
        const int rot = k & 31, wk = k >> 5;
        T ownmul[WPT];
#pragma unroll
        for (int j = 0; j < WPT; j++) ownmul[j] = ((wk < P.nfw) && (t * WPT + j == wk)) ? 0.f : 1.f;
        const bool tail_live = tail_lane && !((wk == P.nfw) && (lane == rot));
        const T tailmul = tail_live ? 1.f : 0.f;
        // the window's last recombination term: x_m = r[eS] (:700-709)
        const int eS = pe[m].site;
        double nor_last, c_last;
        {
            const double x = P.r[eS];
            nor_last = -x + P.log_ntheta;
            double rho = 1.0 - exp(-x);
            if (rho > 0.99) { rho = 0.99; nor_last = P.log_small + P.log_ntheta; }
            c_last = rho / ((1.0 - rho) * P.Nm1);
        }
        // Inputs of a row: this thread's genotype words, the word holding the target's own allele, the tail word.
        // expand() turns them into the rotated mismatch bits; td = target's own allele mask.
        struct RowIn { uint32_t w[WPT]; uint32_t kw, tw; };
        auto fetch = [&](int site) -> RowIn {
            RowIn in;
            load_words(in.w, gthr + (size_t)(unsigned)site * rowbytes);
            in.kw = P.G[(size_t)site * P.wps + wk];
            in.tw = 0;
            if (tail_warp) in.tw = *reinterpret_cast<const uint32_t *>(gtail + (size_t)(unsigned)site * rowbytes);
            return in;
        };
        auto expand = [&](const RowIn &in, uint32_t (&mw)[WPT], uint32_t &tmw) {
            const uint32_t td = ((in.kw >> rot) & 1u) ? 0xffffffffu : 0u;
#pragma unroll
            for (int j = 0; j < WPT; j++) {
                const uint32_t nb = ~in.w[j] & td;
                mw[j] = __funnelshift_r(nb, nb, rot);
            }
            tmw = ~in.tw & td;
        };
        // one forward step of the state (a, tl):  x <- (x + R) * (mis ? tau : 1), target forced to 0
        V2 a[WPT][16];
        T tl = 0.f;
        auto advance = [&](const RowIn &in, T R, V2 &S0, V2 &S1) {
            uint32_t mw[WPT], tmw;
            expand(in, mw, tmw);
#pragma unroll
            for (int j = 0; j < WPT; j++) {
                const T Rj = R * vmul[j];
                const V2 R2 = make_float2(Rj, Rj);
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    V2 v = RT::add2(a[j][e], R2);
                    if (mw[j] & (1u << (2 * e))) v.x *= tau;
                    if (mw[j] & (2u << (2 * e))) v.y *= tau;
                    if (e == 0) v.x *= ownmul[j];
                    a[j][e] = v;
                    if (e & 1) S1 = RT::add2(S1, v); else S0 = RT::add2(S0, v);
                }
            }
            if (tail_warp) {
                T v = tl + R;
                if (tmw & lanebit) v *= tau;
                tl = v * tailmul;
            }
        };
        auto scale_state = [&](T inv) {
#pragma unroll
            for (int j = 0; j < WPT; j++)
#pragma unroll
                for (int e = 0; e < 16; e++) { a[j][e].x *= inv; a[j][e].y *= inv; }
            tl *= inv;
        };

        // ---------------- forward: checkpoint rows -> top, per-row scalars -> scal ----------------
        {   // row 0 = alpha_begin, target zeroed (:757-779)
            const float *ab = P.alpha_begin + (size_t)k * N;
#pragma unroll
            for (int j = 0; j < WPT; j++)
#pragma unroll
                for (int e = 0; e < 16; e++) {
                    const int n0 = (t * WPT + j) * 32;
                    T x = valid[j] ? ab[n0 + ((2 * e + rot) & 31)] : 0.f;
                    T y = valid[j] ? ab[n0 + ((2 * e + 1 + rot) & 31)] : 0.f;
                    if (e == 0) x *= ownmul[j];
                    a[j][e] = make_float2(x, y);
                }
            if (tail_lane) tl = ab[P.nfw * 32 + lane] * tailmul;
        }
        auto store_row = [&](float *row) {
#pragma unroll
            for (int j = 0; j < WPT; j++)
                if (valid[j]) {
                    float4 *o = reinterpret_cast<float4 *>(row) + (size_t)(j * 8) * TT + t;
#pragma unroll
                    for (int e = 0; e < 8; e++) o[(size_t)e * TT] = make_float4(a[j][2 * e].x, a[j][2 * e].y, a[j][2 * e + 1].x, a[j][2 * e + 1].y);
                }
            if (tail_lane) row[tail_off + lane] = tl;
        };
        int par = 0; // alternates the partial-sum buffer between consecutive reductions
        store_row(top);
        T S;
        {
            V2 S0 = make_float2(0.f, 0.f), S1 = S0;
#pragma unroll
            for (int j = 0; j < WPT; j++)
#pragma unroll
                for (int e = 0; e < 16; e++) { if (e & 1) S1 = RT::add2(S1, a[j][e]); else S0 = RT::add2(S0, a[j][e]); }
            S0 = RT::add2(S0, S1);
            S = reduce(S0.x + S0.y + tl, par ^= 1);
        }
        double prev_ls = (double)P.ls_alpha[k];
        if (t == 0) lsrow[0] = P.ls_alpha[k];
        T R = S * pe[0].c;
        // entries past the window's slice belong to the chunk-level table or its padding; nor is padded too
        Ahead<EntF, 2> entq;
        Ahead<RowIn, 2> inq;
        Ahead<T, 2> cq;
        Ahead<double, 3> norq;
        entq.q[0] = pe[3]; entq.q[1] = pe[4];
        inq.q[0] = fetch(pe[1].site); inq.q[1] = fetch(pe[2].site);
        cq.q[0] = pe[1].c; cq.q[1] = pe[2].c;
        norq.q[0] = pnor[0]; norq.q[1] = pnor[1]; norq.q[2] = pnor[2];
        for (int i = 1; i <= m; i++) {
            if ((i & 15) == 1 && i + 48 <= m) { // table lines (16 entries each) three lines ahead: under load a DRAM round trip is
                prefetch_l2(pe + i + 47);       // longer than the two rows of lead of the streams below
                prefetch_l2(pnor + i + 47);
            }
            const EntF ent = entq.pop_push(pe[i + 4]);       // = pe[i+2]
            const RowIn cin = inq.pop_push(fetch(ent.site)); // inputs of row i; row i+2's are requested
            const T ccur = cq.pop_push(ent.c);               // c of row i
            const double nor_cur = norq.pop_push(pnor[i + 2]); // pnor[i-1]
            V2 S0 = make_float2(0.f, 0.f), S1 = S0;
            const T Rrow = R;
            advance(cin, Rrow, S0, S1);
            S0 = RT::add2(S0, S1);
            T Sl = S0.x + S0.y;
            if (tail_warp) Sl += tl;
            S = reduce(Sl, par ^= 1);
            prev_ls += nor_cur;
            float lsi = (float)prev_ls;
            T inv = 1.f;
            if (S < K.lower || S > K.upper) { // :865-876
                inv = 1.f / S;
                scale_state(inv);
                const double lg = log((double)S);
                prev_ls += lg;
                lsi = (float)((double)lsi + lg);
                R = 1.f;
            } else {
                R = S;
            }
            R *= (i == m) ? (T)c_last : ccur;
            if (i % CK == 0) store_row(top + (size_t)i * pitch);
            if (t == 0) {
                lsrow[i] = lsi;
                sc[i] = make_float2(Rrow, inv);
            }
        }

        // ---------------- backward: top[i] = alpha_i * beta_i ----------------
        // carried in g = b * m_s as in the painter; b_i = g_{i+1} + R' is the plain beta the reference multiplies in.
        // The posterior row is taken BEFORE a rescale of this step while its log-scale receives +log(B) (:1033-1061).
        const T chk = K.ntheta, resc_R = K.inv_ntheta; // B = ntheta * sum g;  R' after a rescale is 1/ntheta
        V2 g[WPT][16];
        T gt = 0.f;
        T Rp = 0.f;
        double prevb = (double)P.ls_beta[k];
        if (MULTI) __syncthreads(); // (sc / lsrow of this target were written by thread 0)
        else __syncwarp();
        auto spiece = [&](int slot, int j, int e) -> float4 * { return srow + ((size_t)(slot * WPT + j) * 8 + e) * TT + t; };
        auto ck_issue = [&](int jb, int slot) { // request the checkpoint row of block jb (nothing if jb < 0); always closes one group
            if (jb >= 0) {
                const float4 *src = reinterpret_cast<const float4 *>(top + (size_t)jb * CK * pitch) + t;
#pragma unroll
                for (int j = 0; j < WPT; j++)
                    if (valid[j]) {
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            const unsigned dst = (unsigned)__cvta_generic_to_shared(spiece(slot, j, e));
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)(j * 8 + e) * TT) : "memory");
                        }
                    }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        const int nblk = m / CK; // index of the last block
        ck_issue(nblk, nblk & 1);
        ck_issue(nblk - 1, (nblk - 1) & 1);
        // per-block inputs, requested one block ahead: the rows' genotype words, their forward scalars, the checkpoint's
        // tail; the rows' site indices (which the word loads depend on) two blocks ahead
        RowIn nin[CK];
        float2 nsc[CK];
        float ntail;
        int nsite[CK]; // sites of the block after next (the one block_request will serve next)
        auto sites_request = [&](int jb) {
            const int b0 = max(jb, 0) * CK;
#pragma unroll
            for (int r = 0; r < CK; r++) nsite[r] = pe[min(b0 + r, m)].site;
        };
        auto block_request = [&](int jb) { // uses nsite = the sites of block jb
            const int b0 = max(jb, 0) * CK;
#pragma unroll
            for (int r = 0; r < CK; r++) {
                nin[r] = fetch(nsite[r]); // (rows past m are never used)
                nsc[r] = sc[min(b0 + r, m)];
            }
            ntail = tail_lane ? top[(size_t)b0 * pitch + tail_off + lane] : 0.f;
        };
        sites_request(nblk);
        block_request(nblk);
        sites_request(nblk - 1);
        // small per-row streams walking down: c of the row, its log-scale (written by the forward sweep) and pnor[i+1]
        Ahead<float, 3> lsq;
        Ahead<double, 3> norbq;
        cq.q[0] = pe[m].c; cq.q[1] = pe[m - 1].c;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            lsq.q[d] = lsrow[max(m - d, 0)];
            norbq.q[d] = pnor[max(m - d + 1, 0)];
        }
        for (int jb = nblk; jb >= 0; jb--) {
            const int b0 = jb * CK;
            const int nb = min(CK, m - b0 + 1); // rows b0 .. b0+nb-1
            RowIn cin[CK];
            float2 csc[CK];
#pragma unroll
            for (int r = 0; r < CK; r++) { cin[r] = nin[r]; csc[r] = nsc[r]; }
            const float ctail = ntail;
            block_request(jb - 1);
            sites_request(jb - 2);
            if (jb >= 6) { // the small per-row streams of the block five blocks down, into L2 (one line each covers several blocks)
                const int pb = (jb - 5) * CK;
                prefetch_l2(pe + pb);
                prefetch_l2(sc + pb);
                prefetch_l2(pnor + pb);
                prefetch_l2(lsrow + pb);
            }
            const int cs = jb & 1;
            asm volatile("cp.async.wait_group 1;" ::: "memory"); // this block's checkpoint has landed (the next one may be in flight)
            // ---- recompute rows b0+1 .. b0+nb-1 from the checkpoint ----
#pragma unroll
            for (int j = 0; j < WPT; j++)
#pragma unroll
                for (int e2 = 0; e2 < 8; e2++) {
                    const float4 v = valid[j] ? *spiece(cs, j, e2) : make_float4(0.f, 0.f, 0.f, 0.f);
                    a[j][2 * e2] = make_float2(v.x, v.y);
                    a[j][2 * e2 + 1] = make_float2(v.z, v.w);
                }
            tl = ctail;
            if (tail_lane) stail[cs * 32 + lane] = tl;
#pragma unroll
            for (int r = 1; r < CK; r++) {
                if (r < nb) {
                    V2 d0 = make_float2(0.f, 0.f), d1 = d0;
                    advance(cin[r], csc[r].x, d0, d1);
                    if (csc[r].y != 1.f) scale_state(csc[r].y);
#pragma unroll
                    for (int j = 0; j < WPT; j++)
                        if (valid[j]) {
#pragma unroll
                            for (int e2 = 0; e2 < 8; e2++)
                                *spiece(2 + (r - 1), j, e2) = make_float4(a[j][2 * e2].x, a[j][2 * e2].y, a[j][2 * e2 + 1].x, a[j][2 * e2 + 1].y);
                        }
                    if (tail_lane) stail[(2 + (r - 1)) * 32 + lane] = tl;
                }
            }
            // ---- backward over the block ----
#pragma unroll
            for (int r = CK - 1; r >= 0; r--) {
                if (r < nb) {
                    const int i = b0 + r;
                    const int slot = (r == 0) ? cs : 2 + (r - 1);
                    float *row = top + (size_t)i * pitch;
                    const T ccur = cq.pop_push(pe[i - 2].c);
                    const float ls_cur = lsq.pop_push(lsrow[max(i - 3, 0)]);
                    const double norb_cur = norbq.pop_push(pnor[max(i - 2, 0)]); // pnor[i+1]
                    uint32_t mw[WPT], tmw;
                    expand(cin[r], mw, tmw);
                    V2 S0 = make_float2(0.f, 0.f), S1 = S0;
                    if (i == m) { // b_m = beta_end (:895-909)
                        const float *be = P.beta_end + (size_t)k * N;
#pragma unroll
                        for (int j = 0; j < WPT; j++)
#pragma unroll
                            for (int e = 0; e < 16; e++) {
                                const int n0 = (t * WPT + j) * 32;
                                g[j][e] = make_float2(valid[j] ? be[n0 + ((2 * e + rot) & 31)] : 0.f,
                                                      valid[j] ? be[n0 + ((2 * e + 1 + rot) & 31)] : 0.f);
                            }
                        if (tail_lane) gt = be[P.nfw * 32 + lane];
                    }
#pragma unroll
                    for (int j = 0; j < WPT; j++) {
                        const T Rj = (i == m) ? 0.f : Rp * vmul[j];
                        const V2 R2 = make_float2(Rj, Rj);
                        float4 *o = reinterpret_cast<float4 *>(row) + (size_t)(j * 8) * TT + t;
#pragma unroll
                        for (int e2 = 0; e2 < 8; e2++) {
                            const float4 av = valid[j] ? *spiece(slot, j, e2) : make_float4(0.f, 0.f, 0.f, 0.f); // alpha_i
                            V2 b0v = RT::add2(g[j][2 * e2], R2), b1v = RT::add2(g[j][2 * e2 + 1], R2);
                            if (e2 == 0) b0v.x *= ownmul[j]; // b[k] = 0
                            if (valid[j]) o[(size_t)e2 * TT] = make_float4(av.x * b0v.x, av.y * b0v.y, av.z * b1v.x, av.w * b1v.y);
                            if (mw[j] & (1u << (4 * e2))) b0v.x *= tau;
                            if (mw[j] & (2u << (4 * e2))) b0v.y *= tau;
                            if (mw[j] & (4u << (4 * e2))) b1v.x *= tau;
                            if (mw[j] & (8u << (4 * e2))) b1v.y *= tau;
                            g[j][2 * e2] = b0v;
                            g[j][2 * e2 + 1] = b1v;
                            S0 = RT::add2(S0, b0v);
                            S1 = RT::add2(S1, b1v);
                        }
                    }
                    S0 = RT::add2(S0, S1);
                    T Sl = S0.x + S0.y;
                    if (tail_warp) {
                        const float tail_a = tail_lane ? stail[slot * 32 + lane] : 0.f;
                        T b = ((i == m) ? gt : gt + Rp) * tailmul;
                        if (tail_lane) row[tail_off + lane] = tail_a * b;
                        if (tmw & lanebit) b *= tau;
                        gt = b;
                        Sl += b;
                    }
                    const T G = reduce(Sl, par ^= 1);
                    const T B = chk * G;
                    // log-scale of the row (thread 0): float accumulations exactly as the reference orders them
                    float lsi = 0.f;
                    if (t == 0) {
                        lsi = ls_cur;
                        if (i == m) lsi = lsi + P.ls_beta[k];
                        else {
                            prevb += (i + 1 == m) ? nor_last : norb_cur;
                            lsi = (float)((double)lsi + prevb);
                        }
                    }
                    if (i < m && (B < K.lower || B > K.upper)) {
                        const T inv = 1.f / B;
#pragma unroll
                        for (int j = 0; j < WPT; j++)
#pragma unroll
                            for (int e = 0; e < 16; e++) { g[j][e].x *= inv; g[j][e].y *= inv; }
                        gt *= inv;
                        const double lg = log((double)B);
                        prevb += lg;
                        lsi = (float)((double)lsi + lg);
                        Rp = resc_R;
                    } else {
                        Rp = G;
                    }
                    Rp *= (i == m) ? (T)c_last : ccur;
                    if (t == 0) lsrow[i] = lsi;
                }
            }
            ck_issue(jb - 2, cs); // the ring slot just consumed receives the checkpoint of block jb-2
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
}

// DistanceMeasure::GetMatrix (src/anc_builder.cpp:108-207) for one SNP of the open window: one CTA per target row n.
// v = number of n-derived sites in [start, snp] (row index into top[n], row 0 = bS_n), rpos_prev / rpos_next = genetic
// positions of the last n-derived site <= snp (else SNP 0) and the first one >= snp (else SNP L-1); all three are
// recomputed from the haplotype-major bit rows, which makes the call stateless (the reference carries them along).
struct DistanceParams {
    const uint32_t *GT;  // [N][lw]
    int lw, N, L, W, w, start, snp;
    const double *rpos;  // [L+1]
    const float *top, *ls;
    const long long *rowoff;
    float *d;            // [N][N]
    int pitch, tt, wpt;  // row layout of `top`: the repaint launch's pitch, team size and words per thread
};

__global__ void __launch_bounds__(256) distance_kernel(const DistanceParams P)
{
    const int n = blockIdx.x, t = threadIdx.x, N = P.N;
    __shared__ int s_v, s_prev, s_next, s_der;
    __shared__ float s_min[8];
    if (t == 0) {
        const uint32_t *row = P.GT + (size_t)n * P.lw;
        const int snp = P.snp;
        auto bit = [&](int s) { return (row[s >> 5] >> (s & 31)) & 1u; };
        // v: derived sites in [start, snp], not counting SNP 0 (the caller increments only for snp > start, and the
        // initial count at snp == start skips snp == 0; anc_builder.cpp:79-90, 487-495)
        int v = 0;
        const int lo = P.start > 0 ? P.start : 1;
        for (int wd = lo >> 5; wd <= (snp >> 5) && lo <= snp; wd++) {
            uint32_t x = row[wd];
            if (wd == (lo >> 5)) x &= 0xffffffffu << (lo & 31);
            if (wd == (snp >> 5)) x &= (snp & 31) == 31 ? 0xffffffffu : ((1u << ((snp & 31) + 1)) - 1u);
            v += __popc(x);
        }
        int p = snp;
        while (!bit(p) && p > 0) p--;
        int q = snp;
        while (!bit(q) && q < P.L - 1) q++;
        s_v = v; s_prev = p; s_next = q; s_der = (int)bit(snp);
    }
    __syncthreads();
    const int v = s_v, rot = n & 31;
    const int nfw32 = (N >> 5) << 5;
    const float *top_n = P.top + (size_t)P.rowoff[n] * P.pitch;
    const float *ls_n = P.ls + P.rowoff[n];
    auto slot = [&](int j) -> int { // the repaint kernel's row layout (see there)
        if (j >= nfw32) return P.tt * P.wpt * 32 + (j - nfw32);
        const int w = j >> 5, s = (j - rot) & 31;
        return (((w % P.wpt) * 8 + (s >> 2)) * P.tt + w / P.wpt) * 4 + (s & 3);
    };
    float *out = P.d + (size_t)n * N;
    const float scale = -1.0f;
    float mn = INFINITY;
    if (s_der || P.snp == 0 || P.snp == P.L - 1) {
        const float *tr = top_n + (size_t)v * P.pitch;
        const float lsp = ls_n[v];
        for (int j = t; j < N; j += blockDim.x) {
            const float m = __fmul_rn(__fadd_rn(fast_log_dev(tr[slot(j)]), lsp), scale);
            out[j] = m;
            mn = fminf(mn, m);
        }
    } else {
        const double rp = P.rpos[s_prev], rn = P.rpos[s_next], rs = P.rpos[P.snp];
        double wl, wr;
        if (rp == rn) { wl = 0.5; wr = 0.5; }
        else { const double den = rn - rp; wl = (rn - rs) / den; wr = (rs - rp) / den; }
        const float *tp = top_n + (size_t)v * P.pitch, *tn = top_n + (size_t)(v + 1) * P.pitch;
        const float lsp = ls_n[v], lsn = ls_n[v + 1];
        const float e_pn = expf(lsp - lsn), e_np = expf(lsn - lsp);
        for (int j = t; j < N; j += blockDim.x) {
            const int sj = slot(j);
            float m;
            if (lsp <= lsn) {
                const float arg = (float)__dadd_rn(__dmul_rn(__dmul_rn(wl, (double)tp[sj]), (double)e_pn), __dmul_rn(wr, (double)tn[sj]));
                m = __fmul_rn(__fadd_rn(fast_log_dev(arg), lsn), scale);
            } else {
                const float arg = (float)__dadd_rn(__dmul_rn(wl, (double)tp[sj]), __dmul_rn(__dmul_rn(wr, (double)tn[sj]), (double)e_np));
                m = __fmul_rn(__fadd_rn(fast_log_dev(arg), lsp), scale);
            }
            out[j] = m;
            mn = fminf(mn, m);
        }
    }
    // row minimum over all j (taken before the diagonal is zeroed, :187-192)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if ((t & 31) == 0) s_min[t >> 5] = mn;
    __syncthreads();
    mn = s_min[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); i++) mn = fminf(mn, s_min[i]);
    for (int j = t; j < N; j += blockDim.x) out[j] = (j == n) ? 0.0f : __fsub_rn(out[j], mn);
}

// =========================================================================================
// Stepping-stone record codec on the device: CollapsedMatrix<float>::DumpToFile (src/collapsed_matrix.hpp:228-265).
// A value joins the current run when fabs(head - v) < 1e-3 * min(head, v) (float difference, comparison in double);
// otherwise it becomes the head of a new run (see rle_kernel below for the parallelisation).
// Record: size_t 1; size_t N; int site; float logscale; int K; float val[K]; int len[K]   (28 + 8K bytes)
// File image of window w, per target: int wb[w]; int wb[w+1]-1; alpha record; beta record (fast_painting.cpp:589-601)
struct RleParams {
    const float *alpha, *beta;        // [T][W][N]
    const float *ls_alpha, *ls_beta;  // [T][W]
    const int *site_begin, *site_end; // [T][W]
    const int *wb;                    // [W+1]
    int T, W, N;
    int *K;                           // [T][W][2] run counts
    const long long *rec_off;         // [T][W] byte offset of the target's block within window w's image (EMIT)
    const long long *img_off;         // [W] byte offset of window w's image within `image` (EMIT)
    char *image;
};

// Byte layout of the W file images of one batch: rec_off[k*W+w] = offset of target k's block inside window w's image
// (exclusive scan over targets of 64 + 8*(Ka+Kb)), win_bytes[w] = size of the image.  One CTA per window.
__global__ void __launch_bounds__(256) rle_offsets_kernel(const int *__restrict__ K, int T, int W,
                                                          long long *__restrict__ rec_off,
                                                          long long *__restrict__ win_bytes)
{
    __shared__ long long wsum[8];
    __shared__ long long base_s;
    const int w = blockIdx.x, t = threadIdx.x, lane = t & 31, wid = t >> 5;
    if (t == 0) base_s = 0;
    __syncthreads();
    for (int k0 = 0; k0 < T; k0 += 256) {
        const int k = k0 + t;
        long long sz = 0;
        if (k < T) {
            const size_t tw = (size_t)k * W + w;
            sz = 64 + 8ll * (K[2 * tw] + K[2 * tw + 1]);
        }
        long long incl = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        long long pre = base_s;
        for (int i = 0; i < wid; i++) pre += wsum[i];
        if (k < T) rec_off[(size_t)k * W + w] = pre + incl - sz;
        __syncthreads();
        if (t == 255) base_s = pre + incl;
        __syncthreads();
    }
    if (t == 0) win_bytes[w] = base_s;
}

// img_off[w] = exclusive scan of win_bytes (img_off[W] = total).  W <= a few hundred: one thread.
__global__ void rle_image_scan_kernel(const long long *__restrict__ win_bytes, int W, long long *__restrict__ img_off)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        long long acc = 0;
        for (int w = 0; w < W; w++) { img_off[w] = acc; acc += win_bytes[w]; }
        img_off[W] = acc;
    }
}

// One WARP per vector, 32 segments scanned at once.  The rule is sequential only through the current run head, and a
// scan that starts from a wrong head falls into step with the true one at the first position where both start a run
// (about every other element does).  Per block of 32 x SEG elements (staged in shared memory, segment stride SEG+1 floats:
// conflict-free both ways): (A) every lane scans its SEG-element segment assuming its first element starts a run
// (lane 0 starts from the true head carried over from the previous block) and keeps a 64-bit mask of run heads and its
// final head; (B) lanes 1..31 take the final head of the lane before them and rescan from their segment's start until
// they meet a head of their own speculative scan, correcting the mask on the way; repeated while any lane's final head
// changed (normally once); (C) popcounts and a warp scan turn the masks into run indices; `EMIT` writes each head's
// value and closes the previous run's length.  A warp-per-vector scan with ballot/ffs/shuffle per head spent ~50
// cycles per element; a thread per vector is fine at config 2 (12 000 vectors) but leaves 420 warps for the 13 000
// vectors of a batch at N = 10 000.
__device__ __forceinline__ bool rle_joins(float head, float x)
{
    const float mn = fminf(head, x); // std::min for non-NaN inputs
    return (double)fabsf(head - x) < 1e-3 * (double)mn;
}

template <bool EMIT, int SEG> __global__ void __launch_bounds__(128) rle_kernel(const RleParams P)
{
    static_assert(SEG == 32 || SEG == 64, "segments of 32 (short vectors) or 64 elements");
    constexpr int STRIDE = SEG + 1, BLK = 32 * SEG, SHIFT = SEG == 64 ? 6 : 5;
    __shared__ float tile[4][32 * STRIDE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *T = tile[wid];
    const long long nvec = (long long)P.T * P.W * 2;
    const int N = P.N;
    for (long long vec = (long long)blockIdx.x * 4 + wid; vec < nvec; vec += (long long)gridDim.x * 4) {
        const int ab = (int)(vec & 1);
        const long long tw = vec >> 1; // k*W + w
        const float *v = (ab ? P.beta : P.alpha) + (size_t)tw * N;
        float *vals = nullptr;
        int *lens = nullptr;
        if (EMIT) {
            const int w = (int)(tw % P.W);
            const int Ka = P.K[2 * tw], Kb = P.K[2 * tw + 1];
            char *blk = P.image + P.img_off[w] + P.rec_off[tw];
            char *rec = blk + 8 + (ab ? 28 + 8 * (size_t)Ka : 0);
            const int K = ab ? Kb : Ka;
            if (lane == 0) {
                if (!ab) {
                    reinterpret_cast<int *>(blk)[0] = P.wb[w];
                    reinterpret_cast<int *>(blk)[1] = P.wb[w + 1] - 1;
                }
                // records are 4-byte aligned only: write the two size_t fields as 32-bit halves
                int *h = reinterpret_cast<int *>(rec);
                h[0] = 1; h[1] = 0; h[2] = N; h[3] = 0;
                h[4] = (ab ? P.site_end : P.site_begin)[tw];
                reinterpret_cast<float *>(rec)[5] = (ab ? P.ls_beta : P.ls_alpha)[tw];
                h[6] = K;
            }
            vals = reinterpret_cast<float *>(rec + 28);
            lens = reinterpret_cast<int *>(rec + 28 + 4 * (size_t)K);
        }
        float head_in = 0.f; // true head entering the block (block 0: element 0 is a head by definition)
        int kbase = 0;       // runs started before this block
        int last_head = 0;   // position of the latest head before this block
        for (int b0 = 0; b0 < N; b0 += BLK) {
            const int nblk = min(BLK, N - b0);
            __syncwarp();
            for (int g = lane; g < nblk; g += 32) T[(g >> SHIFT) * STRIDE + (g & (SEG - 1))] = v[b0 + g];
            __syncwarp();
            const int s0 = lane * SEG;                      // segment start within the block
            const int sn = max(0, min(SEG, nblk - s0));     // elements in this lane's segment
            const float *seg = T + lane * STRIDE;
            // ---- A: speculative scan ----
            unsigned long long heads = 0ull;
            float head = head_in;
            const bool first_is_head = (lane > 0) || (b0 == 0);
            if (sn > 0) {
                int i = 0;
                if (first_is_head) { head = seg[0]; heads = 1ull; i = 1; }
                for (; i < sn; i++) {
                    const float x = seg[i];
                    if (!rle_joins(head, x)) { head = x; heads |= 1ull << i; }
                }
            }
            float hout = head; // for empty segments: passes the incoming head through (fixed up below)
            // ---- B: replace the speculation by the true incoming head, until nothing changes ----
            for (;;) {
                float hin = __shfl_up_sync(0xffffffffu, hout, 1);
                bool changed = false;
                if (lane > 0) {
                    if (sn == 0) {
                        changed = hout != hin;
                        hout = hin;
                    } else {
                        float h = hin;
                        unsigned long long nh = heads;
                        bool met = false;
                        for (int i = 0; i < sn; i++) {
                            const float x = seg[i];
                            const bool is_head = !rle_joins(h, x);
                            const bool was_head = (heads >> i) & 1ull;
                            if (is_head) {
                                h = x;
                                nh |= 1ull << i;
                                if (was_head) { met = true; break; } // both scans restart here: identical from now on
                            } else {
                                nh &= ~(1ull << i);
                            }
                        }
                        heads = nh;
                        if (!met) { // the whole segment was rescanned: its final head may have changed
                            changed = hout != h;
                            hout = h;
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, changed)) break;
            }
            // ---- C: run indices, output ----
            const int cnt = __popcll(heads);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t2;
            }
            if (EMIT) {
                // position of the latest head before this lane's segment
                int mylast = heads ? b0 + s0 + (63 - __clzll(heads)) : -1;
                int prevlast = mylast;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t2 = __shfl_up_sync(0xffffffffu, prevlast, o);
                    if (lane >= o) prevlast = max(prevlast, t2);
                }
                int before = __shfl_up_sync(0xffffffffu, prevlast, 1);
                if (lane == 0) before = -1;
                int prevpos = max(before, last_head);
                int k = kbase + incl - cnt;
                unsigned long long hm = heads;
                while (hm) {
                    const int i = __ffsll((long long)hm) - 1;
                    hm &= hm - 1;
                    const int pos = b0 + s0 + i;
                    vals[k] = seg[i];
                    if (k > 0) lens[k - 1] = pos - prevpos;
                    prevpos = pos;
                    k++;
                }
                last_head = max(last_head, __shfl_sync(0xffffffffu, prevlast, 31));
            }
            kbase += __shfl_sync(0xffffffffu, incl, 31);
            head_in = __shfl_sync(0xffffffffu, hout, 31);
        }
        if (lane == 0) {
            if (EMIT) lens[kbase - 1] = N - last_head;
            else P.K[vec] = kbase;
        }
    }
}

// What a reader of the paint files sees (CollapsedMatrix<float>::ReadFromFile, collapsed_matrix.hpp:268-296): every
// element replaced by the head of its run under DumpToFile's rule.  Applied here to the stepping stones of window w
// while they are still in HBM (alpha/beta [T][W][N] of all T = N targets), so the window repaint can start from them
// without the round trip through the paint files and yet from bit-identical inputs.  One warp per vector.
__global__ void __launch_bounds__(256) collapse_kernel(const float *__restrict__ alpha, const float *__restrict__ beta,
                                                       const float *__restrict__ lsa, const float *__restrict__ lsb, int N,
                                                       int W, int w, float *__restrict__ oa, float *__restrict__ ob,
                                                       float *__restrict__ olsa, float *__restrict__ olsb)
{
    const int lane = threadIdx.x & 31;
    for (int vec = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; vec < 2 * N; vec += (gridDim.x * blockDim.x) >> 5) {
        const int ab = vec & 1, k = vec >> 1;
        const float *v = (ab ? beta : alpha) + ((size_t)k * W + w) * N;
        float *o = (ab ? ob : oa) + (size_t)k * N;
        if (lane == 0) (ab ? olsb : olsa)[k] = (ab ? lsb : lsa)[(size_t)k * W + w];
        float head = v[0];
        if (lane == 0) o[0] = head;
        for (int j0 = 1; j0 < N; j0 += 32) {
            const int idx = j0 + lane;
            const bool valid = idx < N;
            const float x = valid ? v[idx] : 0.f;
            float mine = 0.f;
            int cur = 0;
            for (;;) {
                const float mn = fminf(head, x);
                const bool merge = (double)fabsf(head - x) < 1e-3 * (double)mn;
                const unsigned mask = __ballot_sync(0xffffffffu, valid && lane >= cur && !merge);
                const int f = mask ? __ffs(mask) - 1 : 32;
                if (lane >= cur && lane < f) mine = head; // joins the current run
                if (mask == 0) break;
                head = __shfl_sync(0xffffffffu, x, f);
                if (lane == f) mine = head;               // heads a new run
                cur = f + 1;
            }
            if (valid) o[idx] = mine;
        }
    }
}

} // namespace rp
