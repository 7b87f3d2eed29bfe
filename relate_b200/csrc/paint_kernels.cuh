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
__global__ void pack_snp_major_kernel(const unsigned char *__restrict__ hap, int N, int L, uint32_t *__restrict__ G,
                                      int wps)
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

template <typename ENT>
__global__ void tables_kernel(ENT *__restrict__ ent, const long long *__restrict__ off, int nt, int L, int W,
                              const double *__restrict__ r, const double *__restrict__ Phi,
                              const double *__restrict__ Plo, TableConsts tc, const int *__restrict__ ia,
                              const int *__restrict__ ib, double *__restrict__ lsA, double *__restrict__ lsB)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nt) return;
    ENT *e = ent + off[warp];
    const int D = (int)(off[warp + 1] - off[warp]);
    const int *ja = ia + (size_t)warp * W, *jb = ib + (size_t)warp * W;
    double *oa = lsA + (size_t)warp * W, *ob = lsB + (size_t)warp * W;
    int qa = 0, qb = 0;
    // windows whose forward boundary is index 0 have an empty sum
    while (qa < W && ja[qa] == 0) { if (lane == 0) oa[qa] = 0.0; qa++; }
    double carry = 0.0;
    for (int i0 = 0; i0 < D; i0 += 32) {
        const int i = i0 + lane;
        double nor = 0.0;
        if (i < D) {
            const int a = e[i].site;
            const int b = (i + 1 < D) ? e[i + 1].site : L;
            double x;
            if (b - a <= 16) {
                x = r[a];
                for (int s = a + 1; s < b; s++) x += r[s];
            } else {
                x = (Phi[b] - Phi[a]) + (Plo[b] - Plo[a]);
            }
            nor = -x + tc.log_ntheta;
            double rho = 1.0 - exp(-x);
            if (rho > 0.99) {
                rho = 0.99;
                nor = tc.log_small + tc.log_ntheta;
            }
            e[i].set_c(rho / ((1.0 - rho) * tc.Nm1));
        }
        double incl = nor;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        incl += carry; // cum_i = sum_{j<=i} nor_j
        // forward bases need cum_{ia-1}; backward ones cum_{ib} (finalised below)
        while (qa < W && ja[qa] - 1 < i0 + 32) {
            double v = __shfl_sync(0xffffffffu, incl, ja[qa] - 1 - i0);
            if (lane == 0) oa[qa] = v;
            qa++;
        }
        while (qb < W && jb[qb] < i0 + 32) {
            double v = __shfl_sync(0xffffffffu, incl, jb[qb] - i0);
            if (lane == 0) ob[qb] = v; // provisional: cum_{ib}
            qb++;
        }
        carry = __shfl_sync(0xffffffffu, incl, 31);
    }
    // carry == cum_{m}; lsB = norm + (cum_m - cum_ib)
    const double norm = log(tc.Nm1) - (double)D * tc.log_ntheta;
    __syncwarp();
    for (int w = lane; w < W; w += 32) ob[w] = norm + (carry - ob[w]);
}

// ---------------------------------------------------------------------------------------
// Per-step table entries.
struct EntF { // fp32 mode
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

struct PaintParams {
    const uint32_t *G;     // SNP-major bits
    int wps;               // words per SNP row
    int N, L, W;
    int nfw;               // full 32-haplotype words: N / 32
    int tailn;             // N % 32
    int k0, nt;            // targets [k0, k0+nt)
    int njobs;             // 2*nt
    const void *ent;       // EntF[] or EntD[]
    const long long *off;  // [nt+1]
    const int *ia, *ib;    // [nt][W] boundary indices into the site list
    const double *lsA, *lsB; // [nt][W] log-scale bases
    float *alpha, *beta;   // [nt][W][N]
    float *ls_alpha, *ls_beta; // [nt][W]
    int *queue;            // job counter
    double *scratch;       // fp64 mode only: [gridDim.x][N] staging rows for backward stepping stones
    // model constants, fp64 (fast_painting.hpp:26-39)
    double tau_mul;        // theta_ratio + 1.0  (the mismatch multiplier)
    double prior_n;        // ntheta/(N-1)
    double ntheta;
    double inv_ntheta;
    double lower, upper;
};

// The team of T = blockDim.x threads owns the job's vector.  Thread t owns genotype words
// j*T + t (j < WPT); a word's 32 haplotypes sit in 32 registers, rotated by rot = k & 31.
//
// Register budget: 32*WPT state registers (x2 for fp64) + ~25.  MAXT bounds blockDim so that
// ptxas keeps the state in registers: fp32 WPT=1 -> 1024 threads (<=64 regs), fp32 WPT=2 and
// fp64 WPT=1 -> 512 threads (<=128 regs).
template <typename T, int WPT, bool MULTI>
struct PaintCfg {
    static constexpr int kStateRegs = 32 * WPT * (int)(sizeof(T) / 4);
    static constexpr int kMaxThreads = MULTI ? (kStateRegs <= 32 ? 1024 : 512) : 32;
    static constexpr int kMinBlocks = MULTI ? 1 : (kStateRegs <= 32 ? 16 : 8);
};

template <typename T, int WPT, bool MULTI>
__global__ void __launch_bounds__(PaintCfg<T, WPT, MULTI>::kMaxThreads, PaintCfg<T, WPT, MULTI>::kMinBlocks)
paint_kernel(const PaintParams P)
{
    using RT = Real<T>;
    using V2 = typename RT::V2;
    using Ent = typename RT::Ent;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int TT = blockDim.x, TW = TT >> 5;
    __shared__ int s_job;
    __shared__ T s_part[2][32];

    const T tau = (T)P.tau_mul;
    const T lower = (T)P.lower, upper = (T)P.upper;
    const Ent *ents = reinterpret_cast<const Ent *>(P.ent);
    T *scratch = reinterpret_cast<T *>(P.scratch) + (size_t)blockIdx.x * P.N; // dereferenced in fp64 mode only

    bool valid[WPT];
#pragma unroll
    for (int j = 0; j < WPT; j++) valid[j] = (j * TT + t) < P.nfw;
    const bool tail_warp = (P.tailn > 0) && (warp == 0);
    const bool tail_valid = tail_warp && (lane < P.tailn);

    for (;;) {
        int job;
        if (MULTI) {
            __syncthreads();
            if (t == 0) s_job = atomicAdd(P.queue, 1);
            __syncthreads();
            job = s_job;
        } else {
            job = 0;
            if (lane == 0) job = atomicAdd(P.queue, 1);
            job = __shfl_sync(0xffffffffu, job, 0);
        }
        if (job >= P.njobs) break;
        const int kk = job >> 1, dir = job & 1, k = P.k0 + kk;
        const long long base = P.off[kk];
        const int m = (int)(P.off[kk + 1] - base) - 1;
        const Ent *ent = ents + base;
        const int rot = k & 31, wk = k >> 5;
        bool own[WPT];
#pragma unroll
        for (int j = 0; j < WPT; j++) own[j] = (wk < P.nfw) && (j * TT + t == wk);
        const bool own_tail = tail_valid && (wk == P.nfw) && (lane == rot);

        // boundary bookkeeping, in step order
        const int *bidx = (dir ? P.ib : P.ia) + (size_t)kk * P.W;
        const double *lsb = (dir ? P.lsB : P.lsA) + (size_t)kk * P.W;
        float *outv = (dir ? P.beta : P.alpha) + (size_t)kk * P.W * P.N;
        float *outl = (dir ? P.ls_beta : P.ls_alpha) + (size_t)kk * P.W;
        // q walks windows in the order the job meets them: forward w = q, backward w = W-1-q
        int q = 0;
        auto bpos = [&](int qq) -> int { // step position of the qq-th boundary
            return dir ? (m - bidx[P.W - 1 - qq]) : bidx[qq];
        };
        int nextb = bpos(0);

        // state
        V2 a[WPT][16];
#pragma unroll
        for (int j = 0; j < WPT; j++)
#pragma unroll
            for (int e = 0; e < 16; e++) a[j][e] = RT::mk((T)0, (T)0);
        T tl = (T)0;
        T R = dir ? (T)1 : (T)P.prior_n;  // first step is "(0 + R0) * m"
        const T chk = dir ? (T)P.ntheta : (T)1;      // band is tested on chk*S
        const T resc_R = dir ? (T)P.inv_ntheta : (T)1; // R after a rescale (before *c_i)
        double lsr = 0.0; // log-scale added by rescaling

        // software pipeline: entries 3 steps ahead, genotype words 2 steps ahead
        auto eidx = [&](int p) -> int { p = p > m ? m : p; return dir ? m - p : p; };
        Ent eC = ent[eidx(0)], eN = ent[eidx(1)], eNN = ent[eidx(2)];
        uint32_t wC[WPT], wN[WPT], twC = 0, twN = 0;
#pragma unroll
        for (int j = 0; j < WPT; j++) {
            wC[j] = valid[j] ? P.G[(size_t)eC.site * P.wps + j * TT + t] : 0u;
            wN[j] = valid[j] ? P.G[(size_t)eN.site * P.wps + j * TT + t] : 0u;
        }
        if (tail_warp) {
            twC = P.G[(size_t)eC.site * P.wps + P.nfw];
            twN = P.G[(size_t)eN.site * P.wps + P.nfw];
        }

        for (int p = 0; p <= m + 1; p++) {
            // ---- stepping-stone stores (rare) -------------------------------------------
            // forward: the state after step p-1 (post-rescale) is the stored alpha (:354-374).
            // backward: the stored beta at step p is b = g_old + R' *before* the emission
            // multiply (:481-488), divided by B if this step rescales (:538-551 precede :559-578);
            // it is written in T precision first (fp32: straight into the output row, fp64:
            // into this CTA's scratch row) and finalised after the chain below.
            const bool bnd = dir ? (nextb == p) : (nextb == p - 1);
            int q1 = q;
            if (bnd) {
                while (q1 < P.W && bpos(q1) == nextb) q1++;
                const bool ones = dir && (p == 0);
                const T addR = dir ? R : (T)0;
                auto store_vec = [&](auto *o) {
                    using O = typename std::remove_pointer<decltype(o)>::type;
#pragma unroll
                    for (int j = 0; j < WPT; j++) {
                        if (valid[j]) {
                            const int n0 = (j * TT + t) * 32;
#pragma unroll
                            for (int e = 0; e < 16; e++) {
                                T vx = a[j][e].x + addR, vy = a[j][e].y + addR;
                                if (ones) { vx = (T)1; vy = (T)1; }
                                else if (e == 0 && own[j]) vx = (T)0;
                                o[n0 + ((2 * e + rot) & 31)] = (O)vx;
                                o[n0 + ((2 * e + 1 + rot) & 31)] = (O)vy;
                            }
                        }
                    }
                    if (tail_valid) {
                        T v = tl + addR;
                        if (ones) v = (T)1; else if (own_tail) v = (T)0;
                        o[P.nfw * 32 + lane] = (O)v;
                    }
                };
                if (dir && sizeof(T) == 8) {
                    store_vec(scratch);
                } else {
                    for (int qq = q; qq < q1; qq++) {
                        const int w = dir ? P.W - 1 - qq : qq;
                        store_vec(outv + (size_t)w * P.N);
                        if (!dir && t == 0) outl[w] = (float)(lsb[w] + lsr);
                    }
                }
                if (!dir) { q = q1; nextb = q < P.W ? bpos(q) : 0x7fffffff; }
            }
            if (p > m) break;

            // ---- prefetch --------------------------------------------------------------
            const Ent eNNN = ent[eidx(p + 3)];
            uint32_t wNN[WPT], twNN = 0;
#pragma unroll
            for (int j = 0; j < WPT; j++) wNN[j] = valid[j] ? P.G[(size_t)eNN.site * P.wps + j * TT + t] : 0u;
            if (tail_warp) twNN = P.G[(size_t)eNN.site * P.wps + P.nfw];

            // ---- the step: x <- (x + R) * (mis ? tau : 1), S = sum x -----------------------
            // mis = target derived && reference ancestral.  Interior sites are derived by
            // construction; SNP 0 and SNP L-1 are visited regardless (fast_painting.cpp:52-59,150).
            uint32_t tdm = 0xffffffffu;
            if (p == 0 || p == m) {
                const uint32_t kw = P.G[(size_t)eC.site * P.wps + wk];
                tdm = ((kw >> rot) & 1u) ? 0xffffffffu : 0u;
            }
            V2 S0 = RT::mk((T)0, (T)0), S1 = RT::mk((T)0, (T)0);
            const V2 R2 = RT::mk(R, R);
#pragma unroll
            for (int j = 0; j < WPT; j++) {
                if (valid[j]) {
                    const uint32_t nb = ~wC[j] & tdm;
                    const uint32_t mw = __funnelshift_r(nb, nb, rot);
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        V2 v = RT::add2(a[j][e], R2);
                        if (mw & (1u << (2 * e))) v.x *= tau;
                        if (mw & (2u << (2 * e))) v.y *= tau;
                        if (e == 0 && own[j]) v.x = (T)0;
                        a[j][e] = v;
                        if (e & 1) S1 = RT::add2(S1, v); else S0 = RT::add2(S0, v);
                    }
                }
            }
            S0 = RT::add2(S0, S1);
            T S = S0.x + S0.y;
            if (tail_warp) {
                T v = tl + R;
                if ((~twC & tdm) >> lane & 1u) v *= tau;
                if (!tail_valid || own_tail) v = (T)0;
                tl = v;
                S += v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) S += __shfl_xor_sync(0xffffffffu, S, o);
            if (MULTI) {
                T *part = s_part[p & 1];
                if (lane == 0) part[warp] = S;
                __syncthreads();
                S = part[0];
                for (int w = 1; w < TW; w++) S += part[w];
            }

            // ---- scalar chain: rescale test, next R (fast_painting.cpp:331-352, 536-556) ----
            const T B = chk * S;
            bool rescaled = false;
            if (p > 0 && (B < lower || B > upper)) {
                rescaled = true;
#pragma unroll
                for (int j = 0; j < WPT; j++)
#pragma unroll
                    for (int e = 0; e < 16; e++) { a[j][e].x /= B; a[j][e].y /= B; }
                tl /= B;
                lsr += dir ? (double)fast_log_dev((float)B) : log((double)B);
                R = resc_R;
            } else {
                R = S;
            }
            R *= (T)eC.c;

            if (dir && bnd) { // finalise the backward stepping stone(s) of this step
                for (int qq = q; qq < q1; qq++) {
                    const int w = P.W - 1 - qq;
                    float *o = outv + (size_t)w * P.N;
                    if (sizeof(T) == 8 || rescaled) {
                        const T *src = sizeof(T) == 8 ? scratch : reinterpret_cast<const T *>(o);
#pragma unroll
                        for (int j = 0; j < WPT; j++) {
                            if (valid[j]) {
                                const int n0 = (j * TT + t) * 32;
                                for (int e = 0; e < 32; e++) {
                                    T v = src[n0 + e];
                                    if (rescaled) v /= B;
                                    o[n0 + e] = (float)v;
                                }
                            }
                        }
                        if (tail_valid) {
                            T v = src[P.nfw * 32 + lane];
                            if (rescaled) v /= B;
                            o[P.nfw * 32 + lane] = (float)v;
                        }
                    }
                    if (t == 0) outl[w] = (float)(lsb[w] + lsr);
                }
                q = q1;
                nextb = q < P.W ? bpos(q) : 0x7fffffff;
            }

            // ---- rotate the pipeline -----------------------------------------------------
            eC = eN; eN = eNN; eNN = eNNN;
#pragma unroll
            for (int j = 0; j < WPT; j++) { wC[j] = wN[j]; wN[j] = wNN[j]; }
            twC = twN; twN = twNN;
        }
    }
}

} // namespace rp
