// paint_api.cu — C ABI (include/relate_paint.h) over the kernels in paint_kernels.cuh.
//
// Host side of the B200-native `Relate --mode Paint`: what pipeline/Paint.cpp:17-108 and the
// loop around FastPainting::PaintSteppingStones do in the reference, restructured as
//   chunk files -> HBM-resident bit matrices -> per-target site tables -> persistent paint kernel
//   -> stepping stones (device or host) -> RLE records -> chunk_<c>/paint/relate_<w>.bin.
// No CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/relate_paint.h"

#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_io.hpp"
#include "make_chunks.hpp"
#include "paint_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

} // namespace
namespace rp {
int api_fail(int code, const std::string &msg) { return fail(code, msg); } // for the other translation units (minmatch.cu)
}
namespace {

#define RP_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? RP_ENODEVICE        \
                                                                                        : RP_ECUDA,       \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                               \
    } while (0)

#define RP_TRY(expr)                \
    do {                            \
        int rc_ = (expr);           \
        if (rc_ != RP_OK) return rc_; \
    } while (0)

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// RP_TRACE=1: wall-clock marks of the stage driver on stderr (cold-start analysis)
void trace(const char *what, int dev = -1)
{
    static const bool on = getenv("RP_TRACE") != nullptr;
    static const double t_first = now_ms();
    if (on) fprintf(stderr, "[rp %9.2f ms] dev %d: %s\n", now_ms() - t_first, dev, what);
}

// Device memory released by a DevBuf is parked per device and handed out again to the next request it fits: cudaMalloc /
// cudaFree of the GB-sized buffers of this path (a window's posterior is 5 GB at config 2 and 100 GB at config 4) cost tens
// to hundreds of milliseconds each and synchronise the device.  rp_release_cache() returns everything to the driver; an
// allocation that fails empties the pool and tries once more.
class DevicePool {
  public:
    void *take(int dev, size_t want, size_t *got)
    {
        std::lock_guard<std::mutex> lk(mu_);
        auto &m = free_[dev];
        auto it = m.lower_bound(want);
        if (it == m.end() || it->first > want + want / 2 + (1u << 20)) return nullptr; // (do not burn a much larger block)
        void *p = it->second;
        *got = it->first;
        held_ -= it->first;
        m.erase(it);
        return p;
    }
    void give(int dev, void *p, size_t cap)
    {
        std::lock_guard<std::mutex> lk(mu_);
        free_[dev].emplace(cap, p);
        held_ += cap;
    }
    void flush() // frees every parked block (on whatever device it lives)
    {
        std::lock_guard<std::mutex> lk(mu_);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto &kv : free_) {
            cudaSetDevice(kv.first);
            for (auto &b : kv.second) cudaFree(b.second);
            kv.second.clear();
        }
        cudaSetDevice(cur);
        held_ = 0;
    }

  private:
    std::mutex mu_;
    std::map<int, std::multimap<size_t, void *>> free_;
    size_t held_ = 0;
};
DevicePool g_pool;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int dev = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return RP_OK;
        release();
        cudaGetDevice(&dev);
        size_t want = bytes + bytes / 8 + 256;
        if ((p = g_pool.take(dev, bytes, &cap)) != nullptr) return RP_OK;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            g_pool.flush();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(RP_ENOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
        }
        cap = want;
        return RP_OK;
    }
    void release()
    {
        if (p) g_pool.give(dev, p, cap);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

} // namespace

struct rp_chunk {
    int device = 0;
    int N = 0, L = 0, W = 0;
    int wps = 0, lw = 0, nfw = 0, tailn = 0;
    int padbit = 0; // pad bits of a partial last word in G: 1 when theta > 1/2 (the painter's phantoms take the smaller multiplier)
    unsigned flags = 0;
    double theta = 0.001;
    int sm_count = 148;
    std::vector<int> wb;
    std::vector<long long> site_prefix; // [N+1] prefix of D_k = visited sites per target (counted once when the chunk is loaded)
    rp_tune tune{};
    // resident
    DevBuf G, GT, r, Phi, Plo, wbdev, chars, cnt_all; // cnt_all: D_k of every target
    // per-paint work buffers (grown on demand, reused)
    DevBuf counts, off, ent, ia, ib, lsA, lsB, sb, se, alpha, beta, lsa, lsb, queue, scratch, nor, segstate, segdone;
    // record encoder (device RLE): run counts, byte offsets, the W file images of the last batch
    DevBuf rleK, rec_off, win_bytes, img_off, image, image_alt; // image_alt: the stage driver alternates the two
    std::vector<long long> h_img_off; // W+1 offsets of the last encoded batch
    int enc_targets = 0;              // targets in c->image (0: none)
    long long last_sites = 0;         // visited sites (sum of chain lengths) of the batch being painted
    bool resident_all = false;        // alpha/beta/lsa/lsb hold the stepping stones of ALL targets (last paint was 0..N)
    long long *h_total = nullptr; // pinned
    cudaStream_t stream = nullptr;       // the stream every copy and kernel of this chunk is issued on
    cudaStream_t own_stream = nullptr;   // created by the library; `stream` may be replaced by a caller's
    cudaStream_t copy_stream = nullptr;  // device->host copies of encoded records (stage driver), overlapping the next batch
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<rp_chunk **> window_refs; // the `c` fields of the windows open on this chunk (cleared by rp_chunk_free)
    // buffers of the last closed window, kept for the next one (BuildTopology opens a chunk's windows one after the other;
    // allocating the posterior of a window, GBs, costs more than repainting it)
    DevBuf park_top, park_ls, park_scal, park_d, park_rowoff, park_rpos, park_lsa, park_lsb;
};

namespace {

struct LaunchPlan {
    int wpt = 1;
    bool multi = false;
    int threads = 32;
    int cluster = 1; // CTAs per team (thread-block cluster with distributed shared memory when > 1)
};

int plan_launch(const rp_chunk *c, LaunchPlan &lp)
{
    const int nfw = (c->N + 31) / 32; // genotype words per row, a partial last word included (the painter pads it with phantoms)
    const bool fp64 = (c->flags & RP_FP64) != 0;
    int wpt = c->tune.words_per_thread;
    if (wpt != 0 && wpt != 1 && wpt != 2) return fail(RP_EINVAL, "words_per_thread must be 0, 1 or 2");
    if (fp64) wpt = 1;
    if (wpt == 0) wpt = (nfw <= 32) ? 1 : 2;
    if (!fp64 && wpt == 1 && nfw > 384) wpt = 2; // one word per thread: teams of at most 384 threads
    const int need = std::max(1, (nfw + wpt - 1) / wpt);
    lp.wpt = wpt;
    lp.multi = need > 32;
    lp.threads = ((need + 31) / 32) * 32;
    const int maxt = (fp64 || wpt == 1) ? 384 : 512; // PaintCfg::kMaxThreads
    const int forced = c->tune.reserved[3];          // tests: force a cluster of this many CTAs per team
    if (!fp64 && lp.multi && wpt == 2 && (lp.threads > maxt || forced > 1)) {
        // a team of several CTAs: thread-block cluster, sums combined through distributed shared memory
        int cs = std::max(forced, (need + 511) / 512);
        if (cs > 16) return fail(RP_EUNSUPPORTED, "N=" + std::to_string(c->N) + " is beyond a 16-CTA cluster (524288 haplotypes)");
        lp.cluster = cs;
        lp.threads = (((need + cs - 1) / cs + 31) / 32) * 32;
        return RP_OK;
    }
    if (lp.threads > maxt)
        return fail(RP_EUNSUPPORTED, "N=" + std::to_string(c->N) + " needs " + std::to_string(lp.threads) +
                                         " threads per team; one CTA owns at most " + std::to_string(maxt * wpt * 32) +
                                         " haplotypes in this mode (fp64 verification mode has no cluster variant)");
    return RP_OK;
}

template <typename T, int WPT, bool MULTI, bool DENSE = false>
int launch_paint_t(rp_chunk *c, rp::PaintParams &P, int threads, int &ctas)
{
    auto kern = rp::paint_kernel<T, WPT, MULTI, DENSE>;
    int occ = 0;
    RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0));
    if (occ < 1) return fail(RP_ECUDA, "paint kernel does not fit on an SM");
    if (c->tune.ctas_per_sm > 0) occ = std::min(occ, c->tune.ctas_per_sm);
    const int slots = std::max(1, occ * c->sm_count / 2); // resident teams per direction
    // Chain segments: with whole chains as jobs, a kernel whose 2*nt chains do not fill the SM sub-partitions evenly
    // (config 2: 2000 chains on 592 schedulers = 3 or 4 each) ends when the fullest sub-partitions do.  Cut into
    // segments whose state is parked in HBM in between, chains migrate to whichever team is free, so lightly loaded
    // sub-partitions get through more segments; the same cut shortens the tail of multi-wave launches.
    // Off for a launch that cannot fill half the slots (nothing to balance) and in the fp64 verification mode
    // (its backward stepping-stone staging rows are per CTA).
    int nseg = 1;
    if (sizeof(T) == 4) {
        const int want = c->tune.reserved[1];
        if (want > 0) nseg = want;
        else if (2 * P.nt > slots) // a park/resume costs a few microseconds: keep segments at >= ~512 steps
            nseg = (int)std::max<long long>(1, std::min<long long>(8, c->last_sites / std::max(1, P.nt) / 512));
    }
    P.nseg = nseg;
    P.segready = nullptr;
    P.segstate = nullptr;
    P.segstride = 0;
    if (nseg > 1) {
        P.segstride = (((size_t)threads * WPT * 32 + 32) * sizeof(T) + 64 + 15) & ~(size_t)15;
        RP_TRY(c->segstate.ensure(2 * (size_t)P.nt * P.segstride));
        const size_t rq = 2 * (size_t)P.nt * (nseg - 1) * 4;
        RP_TRY(c->segdone.ensure(rq));
        RP_CUDA(cudaMemsetAsync(c->segdone.p, 0, rq, c->stream));
        P.segready = c->segdone.as<int>();
        P.segstate = c->segstate.as<char>();
    }
    // with segments the grid is the full occupancy even when there are fewer chains than slots: the idle teams are what
    // lets chains migrate towards lightly loaded sub-partitions
    ctas = 2 * (int)std::min<long long>((long long)P.nt * nseg, slots); // even CTAs paint forwards, odd ones backwards
    kern<<<ctas, threads, 0, c->stream>>>(P);
    RP_CUDA(cudaGetLastError());
    return RP_OK;
}

// grid size is needed before the launch to size the fp64 scratch; compute it the same way
template <typename T, int WPT, bool MULTI> int grid_for(const rp_chunk *c, int nt, int threads, int &ctas)
{
    auto kern = rp::paint_kernel<T, WPT, MULTI>;
    int occ = 0;
    RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0));
    if (occ < 1) return fail(RP_ECUDA, "paint kernel does not fit on an SM");
    if (c->tune.ctas_per_sm > 0) occ = std::min(occ, c->tune.ctas_per_sm);
    ctas = 2 * std::min(nt, std::max(1, occ * c->sm_count / 2));
    return RP_OK;
}

int launch_paint(rp_chunk *c, rp::PaintParams &P, const LaunchPlan &lp, DevBuf &scratch, int &ctas)
{
    const bool fp64 = (c->flags & RP_FP64) != 0;
    if (fp64) {
        if (lp.multi) RP_TRY((grid_for<double, 1, true>(c, P.nt, lp.threads, ctas)));
        else RP_TRY((grid_for<double, 1, false>(c, P.nt, lp.threads, ctas)));
        RP_TRY(scratch.ensure((size_t)ctas * c->N * sizeof(double)));
        P.scratch = scratch.as<double>();
        return lp.multi ? launch_paint_t<double, 1, true>(c, P, lp.threads, ctas)
                        : launch_paint_t<double, 1, false>(c, P, lp.threads, ctas);
    }
    P.scratch = nullptr;
    if (lp.cluster > 1) {
        auto kern = rp::paint_cluster_kernel;
        if (lp.cluster > 8) RP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)lp.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.blockDim = dim3((unsigned)lp.threads);
        cfg.gridDim = dim3((unsigned)lp.cluster); // provisional, for the occupancy query
        cfg.dynamicSmemBytes = 0;
        cfg.stream = c->stream;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int max_clusters = 0;
        RP_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
        if (max_clusters < 1) return fail(RP_EUNSUPPORTED, "a cluster of " + std::to_string(lp.cluster) + " CTAs cannot be co-scheduled on this device");
        const int teams = 2 * std::min(P.nt, std::max(1, max_clusters / 2)); // even teams paint forwards, odd ones backwards
        ctas = teams * lp.cluster;
        cfg.gridDim = dim3((unsigned)ctas);
        RP_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
        return RP_OK;
    }
    if (lp.wpt == 1)
        return lp.multi ? launch_paint_t<float, 1, true>(c, P, lp.threads, ctas)
                        : launch_paint_t<float, 1, false>(c, P, lp.threads, ctas);
    if (lp.multi && lp.threads <= 160 && c->tune.reserved[0] == 2) // opt-in: measured no faster than the 128-register variant
        return launch_paint_t<float, 2, true, true>(c, P, lp.threads, ctas); // 96-register variant: more CTAs per SM
    return lp.multi ? launch_paint_t<float, 2, true>(c, P, lp.threads, ctas)
                    : launch_paint_t<float, 2, false>(c, P, lp.threads, ctas);
}

// double-double prefix of r: P[s] = sum_{j<s} r[j] as hi+lo
void dd_prefix(const double *r, int L, std::vector<double> &hi, std::vector<double> &lo)
{
    hi.assign((size_t)L + 1, 0.0);
    lo.assign((size_t)L + 1, 0.0);
    double h = 0.0, l = 0.0;
    for (int s = 0; s < L; s++) {
        const double a = h, b = r[s];
        const double sum = a + b;
        const double bb = sum - a;
        const double err = (a - (sum - bb)) + (b - bb); // TwoSum
        l += err;
        const double nh = sum + l; // renormalise (FastTwoSum)
        l = l - (nh - sum);
        h = nh;
        hi[s + 1] = h;
        lo[s + 1] = l;
    }
}

// Genotype bytes that are still being read from disk: slice i of `slice` bytes may be copied once ready[i] != 0
// (1 = read, -1 = read failed).  Lets the host->device copy run behind the file reads.
// The slices pass through a ring of `nslots` pinned slots (slice i sits in slot i % nslots): pinning the whole chunk
// cost ~0.3 s per 0.5 GB at every cold start.  When the ring is shorter than the chunk, a slot is refilled only after
// every device has copied the slice it held (consumed[i] == ndev).
// The reader threads bit-pack what they read (SNP-major rows of `wps` words, the layout of rp_chunk::G): slice i is
// `rows` whole SNP rows, `slice` = rows * wps * 4 bytes in the ring.  The devices then copy 1 bit per genotype instead
// of 1 byte (8 devices painting one chunk pull 1/8 of the bytes through the host's memory and PCIe), the chars never
// reach HBM and the pack kernel is not launched.
struct HapFeed {
    size_t slice = 0; // bytes of one ring slot
    int rows = 0;     // SNP rows per slice
    int nsl = 0, nslots = 0, ndev = 1;
    int padbit = 0;   // genotype bit of the slots past N in a partial last word (1 when theta > 1/2)
    const char *ring = nullptr;
    std::unique_ptr<std::atomic<int>[]> ready, consumed;
    std::atomic<int> abort{0};
    const char *src(int i) const { return ring + (size_t)(i % nslots) * slice; }
    bool wraps() const { return nsl > nslots; }
};

// chars '0'/'1' of one SNP row -> bits (bit n&31 of word n>>5 = hap[n] == '1'), words beyond the row zeroed;
// what pack_snp_major_kernel does on the device
inline void pack_row_host(const char *p, int N, uint32_t *out, int wps, int padbit = 0)
{
    const __m128i one = _mm_set1_epi8('1');
    int n = 0, w = 0;
    for (; n + 32 <= N; n += 32, w++) {
        const unsigned lo = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(p + n)), one));
        const unsigned hi = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i *>(p + n + 16)), one));
        out[w] = lo | (hi << 16);
    }
    if (n < N) {
        uint32_t bits = 0;
        for (int j = 0; n + j < N; j++) bits |= (uint32_t)(p[n + j] == '1') << j;
        if (padbit) bits |= ~0u << (N - n); // phantom slots of the partial last word (tau > 1)
        out[w++] = bits;
    }
    for (; w < wps; w++) out[w] = 0;
}

int chunk_from_host(int device, int N, int L, const char *hap, const double *r, const int *wb, int n_wb,
                    double theta, unsigned flags, rp_chunk **out, rp_stats *st, HapFeed *feed = nullptr)
{
    if (!out || !hap || !r || !wb) return fail(RP_EINVAL, "null argument");
    if (N < 2 || L < 2 || n_wb < 2) return fail(RP_EINVAL, "need N>=2, L>=2 and at least one window");
    if (wb[0] != 0 || wb[n_wb - 1] != L) return fail(RP_EINVAL, "window boundaries must start at 0 and end at L");
    for (int i = 1; i < n_wb; i++)
        if (wb[i] <= wb[i - 1]) return fail(RP_EINVAL, "window boundaries must be strictly increasing");
    if (!(theta > 0.0 && theta < 1.0)) return fail(RP_EINVAL, "theta must be in (0,1)");
    int ndev = 0;
    RP_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(RP_ENODEVICE, "no CUDA device");
    if (device < 0 || device >= ndev) return fail(RP_EINVAL, "device index out of range");
    RP_CUDA(cudaSetDevice(device));

    // *out may carry a parked chunk of the same device whose buffers, stream and events are reused
    rp_chunk *c = (*out && (*out)->device == device) ? *out : nullptr;
    const bool reused = c != nullptr;
    if (*out && !reused) rp_chunk_free(*out);
    *out = nullptr;
    if (!c) c = new rp_chunk();
    auto bail = [&](int rc) {
        std::string keep = g_err;
        rp_chunk_free(c);
        g_err = keep;
        return rc;
    };
#define RP_TRYB(expr)                      \
    do {                                   \
        int rc_ = (expr);                  \
        if (rc_ != RP_OK) return bail(rc_); \
    } while (0)
#define RP_CUDAB(call)                                                                               \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return bail(fail(RP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); \
    } while (0)

    c->device = device;
    c->N = N;
    c->L = L;
    c->W = n_wb - 1;
    c->flags = flags;
    c->theta = theta;
    c->wb.assign(wb, wb + n_wb);
    c->nfw = N / 32;
    c->tailn = N % 32;
    c->padbit = theta > 0.5 ? 1 : 0;
    c->wps = (((N + 31) / 32) + 3) / 4 * 4; // rows padded to 16 bytes
    c->lw = (L + 31) / 32;
    if (!reused) {
        int major = 0, sms = 0;
        RP_CUDAB(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
        RP_CUDAB(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        c->sm_count = sms;
        if (major < 10) return bail(fail(RP_ENODEVICE, "device is not sm_100-class; this library ships sm_100a code only"));
        RP_CUDAB(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        RP_CUDAB(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (auto &e : c->ev) RP_CUDAB(cudaEventCreate(&e));
        RP_CUDAB(cudaMallocHost(&c->h_total, sizeof(long long)));
    }
    c->stream = c->own_stream;
    c->tune = rp_tune{};
    LaunchPlan lp;
    RP_TRYB(plan_launch(c, lp));

    const double t0 = now_ms();
    DevBuf &chars = c->chars;
    const size_t nchar = (size_t)L * N;
    if (!feed) RP_TRYB(chars.ensure(nchar)); // (a feed delivers packed rows straight into G)
    RP_TRYB(c->G.ensure((size_t)L * c->wps * 4));
    RP_TRYB(c->GT.ensure((size_t)N * c->lw * 4));
    RP_TRYB(c->r.ensure((size_t)L * 8));
    RP_TRYB(c->Phi.ensure((size_t)(L + 1) * 8));
    RP_TRYB(c->Plo.ensure((size_t)(L + 1) * 8));
    RP_TRYB(c->wbdev.ensure((size_t)n_wb * 4));
    RP_CUDAB(cudaEventRecord(c->ev[0], c->stream));
    std::vector<double> hi, lo;
    dd_prefix(r, L, hi, lo);
    if (!feed) {
        RP_CUDAB(cudaMemcpyAsync(chars.p, hap, nchar, cudaMemcpyHostToDevice, c->stream));
    } else {
        for (int i = 0; i < feed->nsl; i++) {
            int state;
            while ((state = feed->ready[i].load(std::memory_order_acquire)) == 0) std::this_thread::yield();
            if (state < 0) return bail(fail(RP_EIO, "short read in the chunk's .hap file"));
            const size_t rowb = (size_t)c->wps * 4;
            const size_t at = (size_t)i * feed->rows * rowb, n = (size_t)std::min(feed->rows, L - i * feed->rows) * rowb;
            RP_CUDAB(cudaMemcpyAsync(c->G.as<char>() + at, feed->src(i), n, cudaMemcpyHostToDevice, c->stream));
            if (feed->wraps()) { // the slot is reused: tell the readers once this device has its copy
                RP_CUDAB(cudaStreamSynchronize(c->stream));
                feed->consumed[i].fetch_add(1, std::memory_order_release);
            }
        }
    }
    RP_CUDAB(cudaMemcpyAsync(c->r.p, r, (size_t)L * 8, cudaMemcpyHostToDevice, c->stream));
    RP_CUDAB(cudaMemcpyAsync(c->Phi.p, hi.data(), (size_t)(L + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    RP_CUDAB(cudaMemcpyAsync(c->Plo.p, lo.data(), (size_t)(L + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    RP_CUDAB(cudaMemcpyAsync(c->wbdev.p, wb, (size_t)n_wb * 4, cudaMemcpyHostToDevice, c->stream));
    RP_CUDAB(cudaEventRecord(c->ev[1], c->stream));
    {
        const long long total = (long long)L * c->wps;
        const int th = 256;
        if (!feed) {
            rp::pack_snp_major_kernel<<<(unsigned)((total + th - 1) / th), th, 0, c->stream>>>(
                chars.as<unsigned char>(), N, L, c->G.as<uint32_t>(), c->wps, c->padbit);
            RP_CUDAB(cudaGetLastError());
        }
        dim3 grid((N + 31) / 32, (c->lw + 31) / 32), block(32, 32);
        rp::transpose_bits_kernel<<<grid, block, 0, c->stream>>>(c->G.as<uint32_t>(), c->wps, N, L,
                                                                 c->GT.as<uint32_t>(), c->lw);
        RP_CUDAB(cudaGetLastError());
    }
    // D_k of every target, once per chunk: a paint call then knows its table size without counting and without a
    // device->host round trip in front of its kernels
    RP_TRYB(c->cnt_all.ensure((size_t)N * 4));
    rp::count_sites_kernel<<<(unsigned)((N + 7) / 8), 256, 0, c->stream>>>(c->GT.as<uint32_t>(), c->lw, L, 0, N, c->cnt_all.as<int>());
    RP_CUDAB(cudaGetLastError());
    std::vector<int> h_cnt((size_t)N);
    RP_CUDAB(cudaMemcpyAsync(h_cnt.data(), c->cnt_all.p, (size_t)N * 4, cudaMemcpyDeviceToHost, c->stream));
    RP_CUDAB(cudaEventRecord(c->ev[2], c->stream));
    RP_CUDAB(cudaStreamSynchronize(c->stream));
    c->site_prefix.assign((size_t)N + 1, 0);
    for (int k = 0; k < N; k++) c->site_prefix[(size_t)k + 1] = c->site_prefix[(size_t)k] + h_cnt[(size_t)k];
    if (!reused) chars.release(); // a parked workspace keeps its staging buffer for the next chunk
    if (st) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
        st->ms_h2d += a;
        st->ms_prep += b;
        st->h2d_bytes += (feed ? (long long)L * c->wps * 4 : (long long)nchar) + (long long)L * 8 + (long long)(L + 1) * 16 + (long long)n_wb * 4;
        st->launches += feed ? 2 : 3;
        st->ms_total += now_ms() - t0;
    }
    *out = c;
    return RP_OK;
#undef RP_TRYB
#undef RP_CUDAB
}

// Runs prep + paint for targets [k0,k1) on the chunk's stream; results stay in c->alpha etc.
int paint_device(rp_chunk *c, int k0, int k1, rp_stats *st, bool run_paint = true, bool want_nor = false)
{
    if (!c) return fail(RP_EINVAL, "null chunk");
    if (k0 < 0 || k1 > c->N || k0 >= k1) return fail(RP_EINVAL, "bad target range");
    RP_CUDA(cudaSetDevice(c->device));
    const bool fp64 = (c->flags & RP_FP64) != 0;
    const int nt = k1 - k0, W = c->W, N = c->N;
    LaunchPlan lp;
    RP_TRY(plan_launch(c, lp));
    const size_t nw = (size_t)nt * W;
    RP_TRY(c->counts.ensure((size_t)nt * 4));
    RP_TRY(c->off.ensure((size_t)(nt + 1) * 8));
    RP_TRY(c->ia.ensure(nw * 4));
    RP_TRY(c->ib.ensure(nw * 4));
    RP_TRY(c->lsA.ensure(nw * 8));
    RP_TRY(c->lsB.ensure(nw * 8));
    RP_TRY(c->sb.ensure(nw * 4));
    RP_TRY(c->se.ensure(nw * 4));
    RP_TRY(c->lsa.ensure(nw * 4));
    RP_TRY(c->lsb.ensure(nw * 4));
    if (run_paint) {
        RP_TRY(c->alpha.ensure(nw * N * 4));
        RP_TRY(c->beta.ensure(nw * N * 4));
    }
    RP_TRY(c->queue.ensure(16));

    cudaStream_t s = c->stream;
    int launches = 0;
    RP_CUDA(cudaEventRecord(c->ev[0], s));
    const int th = 256, wpb = th / 32;
    const unsigned gw = (unsigned)((nt + wpb - 1) / wpb);
    rp::scan_counts_kernel<<<1, 1024, 0, s>>>(c->cnt_all.as<int>() + k0, nt, c->off.as<long long>());
    RP_CUDA(cudaGetLastError());
    launches += 1;
    const long long U = c->site_prefix[(size_t)k1] - c->site_prefix[(size_t)k0];
    c->last_sites = U;
    // the paint kernel's prefetch reads up to 3 entries past either end of a target's list: pad both ends
    const size_t entsz = fp64 ? sizeof(rp::EntD) : sizeof(rp::EntF);
    const size_t pad = 4;
    RP_TRY(c->ent.ensure(((size_t)U + 2 * pad) * entsz));
    RP_CUDA(cudaMemsetAsync(c->ent.p, 0, pad * entsz, s));
    RP_CUDA(cudaMemsetAsync(c->ent.as<char>() + (pad + (size_t)U) * entsz, 0, pad * entsz, s));
    double *nor_out = nullptr;
    if (want_nor) {
        RP_TRY(c->nor.ensure(((size_t)U + 8) * 8)); // the window repaint reads up to 2 entries past a slice
        nor_out = c->nor.as<double>();
    }

    rp::TableConsts tc;
    const double ntheta = 1.0 - c->theta;
    tc.log_ntheta = log(ntheta);
    tc.log_small = log(0.01);
    tc.Nm1 = N - 1.0;
    const unsigned gb = (unsigned)((nw + th - 1) / th);
    if (fp64) {
        auto *e = c->ent.as<rp::EntD>() + pad;
        rp::fill_sites_kernel<<<gw, th, 0, s>>>(c->GT.as<uint32_t>(), c->lw, c->L, k0, nt, c->off.as<long long>(), e);
        rp::boundaries_kernel<<<gb, th, 0, s>>>(e, c->off.as<long long>(), nt, W, c->wbdev.as<int>(), c->ia.as<int>(),
                                                c->ib.as<int>(), c->sb.as<int>(), c->se.as<int>());
        rp::tables_kernel<rp::EntD, 16><<<gw, th, 0, s>>>(e, c->off.as<long long>(), nt, c->L, W, c->r.as<double>(),
                                            c->Phi.as<double>(), c->Plo.as<double>(), tc, c->ia.as<int>(),
                                            c->ib.as<int>(), c->lsA.as<double>(), c->lsB.as<double>(), nor_out);
    } else {
        auto *e = c->ent.as<rp::EntF>() + pad;
        rp::fill_sites_cta_kernel<rp::EntF, 8><<<nt, 256, 0, s>>>(c->GT.as<uint32_t>(), c->lw, c->L, k0, nt, c->off.as<long long>(), e);
        rp::boundaries_kernel<<<gb, th, 0, s>>>(e, c->off.as<long long>(), nt, W, c->wbdev.as<int>(), c->ia.as<int>(),
                                                c->ib.as<int>(), c->sb.as<int>(), c->se.as<int>());
        rp::tables_cta_kernel<rp::EntF, 8><<<nt, 256, 0, s>>>(e, c->off.as<long long>(), nt, c->L, W, c->Phi.as<double>(),
                                                             c->Plo.as<double>(), tc, c->ia.as<int>(), c->ib.as<int>(),
                                                             c->lsA.as<double>(), c->lsB.as<double>(), nor_out);
    }
    RP_CUDA(cudaGetLastError());
    launches += 3;
    RP_CUDA(cudaMemsetAsync(c->queue.p, 0, 16, s));
    RP_CUDA(cudaEventRecord(c->ev[1], s));
    if (!run_paint) { // site tables only (the window repaint builds on them)
        RP_CUDA(cudaStreamSynchronize(s));
        if (st) {
            float a = 0;
            cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
            st->ms_prep += a;
            st->sites += U;
            st->launches += launches;
        }
        return RP_OK;
    }

    rp::PaintParams P{};
    P.G = c->G.as<uint32_t>();
    P.wps = c->wps;
    P.N = N;
    P.L = c->L;
    P.W = W;
    P.nfw = (N + 31) / 32; // the partial last word counts as a word of the team
    P.tailn = c->tailn;
    P.padbit = c->padbit;
    P.k0 = k0;
    P.nt = nt;
    P.ent = c->ent.as<char>() + pad * entsz;
    P.off = c->off.as<long long>();
    P.ia = c->ia.as<int>();
    P.ib = c->ib.as<int>();
    P.lsA = c->lsA.as<double>();
    P.lsB = c->lsB.as<double>();
    P.alpha = c->alpha.as<float>();
    P.beta = c->beta.as<float>();
    P.ls_alpha = c->lsa.as<float>();
    P.ls_beta = c->lsb.as<float>();
    P.queue = c->queue.as<int>();
    // fast_painting.hpp:26-39, evaluated in fp64 exactly as written there, then rounded once for the fp32 kernel
    const double theta_ratio = c->theta / (1.0 - c->theta) - 1.0;
    P.cd.tau = 1.0 * theta_ratio + 1.0;
    P.cd.prior_n = ntheta / (N - 1.0);
    P.cd.ntheta = ntheta;
    P.cd.inv_ntheta = 1.0 / ntheta;
    P.cd.lower = 1e-10;
    P.cd.upper = 1.0 / P.cd.lower;
    P.cf.tau = (float)P.cd.tau;
    P.cf.prior_n = (float)P.cd.prior_n;
    P.cf.ntheta = (float)P.cd.ntheta;
    P.cf.inv_ntheta = (float)P.cd.inv_ntheta;
    P.cf.lower = (float)P.cd.lower;
    P.cf.upper = (float)P.cd.upper;
    {   // Fixed-point unit of the single-REDUX team sum: 2^(E - 29 + h), E the exponent of the bound S + N*R the kernel
        // forms per step.  That bound assumes multipliers <= 1; for theta > 1/2 the mismatch multiplier tau = theta/(1-theta)
        // exceeds 1 and the sum can reach tau*(S + N*R): h = ceil(log2(tau)) bits of headroom cover it.
        const double tau = c->theta / ntheta;
        P.hshift = tau > 1.0 ? (int)ceil(log2(tau)) : 0;
        if (P.hshift > 6) return fail(RP_EUNSUPPORTED, "theta too close to 1 for the fp32 painter (use RP_FP64)");
        P.k1c = (283 - P.hshift) << 23;       // k1 = as_float(k1c - exponent bits of the bound) = 2^(29 - h - E)
        P.k2c = (-29 + P.hshift) * (1 << 23); // k2 = as_float(exponent bits + k2c)            = 2^(E + h - 29)
        // band edges for the kernel's rare-path pre-filter (it multiplies them into fixed-point units): forward B = S,
        // backward B = ntheta*S
        for (int dir = 0; dir < 2; dir++) {
            const float chk = dir ? P.cf.ntheta : 1.0f;
            const float lo = P.cf.lower / chk * 1.000002f, hi = P.cf.upper / chk * 0.999998f;
            int lob, hib;
            memcpy(&lob, &lo, 4);
            memcpy(&hib, &hi, 4);
            P.xlo[dir] = lob;
            P.xhi[dir] = hib;
        }
    }
    int ctas = 0;
    c->resident_all = false;
    RP_TRY(launch_paint(c, P, lp, c->scratch, ctas));
    c->resident_all = (k0 == 0 && k1 == N);
    launches += 1;
    RP_CUDA(cudaEventRecord(c->ev[2], s));
    RP_CUDA(cudaStreamSynchronize(s));
    if (st) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
        st->ms_prep += a;
        st->ms_paint += b;
        st->sites += U;
        st->cells += (long long)nt * N * c->L;
        st->launches += launches;
        st->n_targets += nt;
        st->team_threads = lp.threads;
        st->words_per_thread = lp.wpt;
        st->ctas = ctas;
    }
    return RP_OK;
}

// Device side of CollapsedMatrix<float>::DumpToFile for the batch painted last: counts runs, lays out the W file
// images (window-major, targets in order) and writes the records into c->image.  c->h_img_off[w] = byte offset of
// window w's image, [W] = total.
int encode_device(rp_chunk *c, int nt, rp_stats *st)
{
    const int W = c->W, N = c->N;
    const size_t nw = (size_t)nt * W;
    cudaStream_t s = c->stream;
    RP_TRY(c->rleK.ensure(nw * 2 * 4));
    RP_TRY(c->rec_off.ensure(nw * 8));
    RP_TRY(c->win_bytes.ensure((size_t)W * 8));
    RP_TRY(c->img_off.ensure((size_t)(W + 1) * 8));
    rp::RleParams P{};
    P.alpha = c->alpha.as<float>();
    P.beta = c->beta.as<float>();
    P.ls_alpha = c->lsa.as<float>();
    P.ls_beta = c->lsb.as<float>();
    P.site_begin = c->sb.as<int>();
    P.site_end = c->se.as<int>();
    P.wb = c->wbdev.as<int>();
    P.T = nt;
    P.W = W;
    P.N = N;
    P.K = c->rleK.as<int>();
    P.rec_off = c->rec_off.as<long long>();
    P.img_off = c->img_off.as<long long>();
    const long long nvec = (long long)nw * 2;
    const unsigned grid = (unsigned)std::min<long long>((nvec + 3) / 4, (long long)c->sm_count * 16); // 4 vectors (warps) per CTA
    RP_CUDA(cudaEventRecord(c->ev[0], s));
    const bool short_vec = N <= 4096; // 32 x 32-element blocks keep all lanes busy on short vectors
    if (short_vec) rp::rle_kernel<false, 32><<<grid, 128, 0, s>>>(P);
    else rp::rle_kernel<false, 64><<<grid, 128, 0, s>>>(P);
    rp::rle_offsets_kernel<<<W, 256, 0, s>>>(c->rleK.as<int>(), nt, W, c->rec_off.as<long long>(),
                                             c->win_bytes.as<long long>());
    rp::rle_image_scan_kernel<<<1, 32, 0, s>>>(c->win_bytes.as<long long>(), W, c->img_off.as<long long>());
    RP_CUDA(cudaGetLastError());
    c->h_img_off.resize(W + 1);
    RP_CUDA(cudaMemcpyAsync(c->h_img_off.data(), c->img_off.p, (size_t)(W + 1) * 8, cudaMemcpyDeviceToHost, s));
    RP_CUDA(cudaStreamSynchronize(s));
    const long long total = c->h_img_off[W];
    RP_TRY(c->image.ensure((size_t)total));
    P.image = c->image.as<char>();
    if (short_vec) rp::rle_kernel<true, 32><<<grid, 128, 0, s>>>(P);
    else rp::rle_kernel<true, 64><<<grid, 128, 0, s>>>(P);
    RP_CUDA(cudaGetLastError());
    RP_CUDA(cudaEventRecord(c->ev[1], s));
    RP_CUDA(cudaStreamSynchronize(s));
    c->enc_targets = nt;
    if (st) {
        float a = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        st->ms_rle += a;
        st->launches += 4;
        st->d2h_bytes += (long long)(W + 1) * 8;
    }
    return RP_OK;
}

int copy_out(rp_chunk *c, int nt, float *alpha, float *beta, float *ls_alpha, float *ls_beta, int *site_begin,
             int *site_end, rp_stats *st)
{
    cudaStream_t s = c->stream;
    const size_t nw = (size_t)nt * c->W;
    long long bytes = 0;
    RP_CUDA(cudaEventRecord(c->ev[0], s));
    if (alpha) { RP_CUDA(cudaMemcpyAsync(alpha, c->alpha.p, nw * c->N * 4, cudaMemcpyDeviceToHost, s)); bytes += nw * c->N * 4; }
    if (beta) { RP_CUDA(cudaMemcpyAsync(beta, c->beta.p, nw * c->N * 4, cudaMemcpyDeviceToHost, s)); bytes += nw * c->N * 4; }
    if (ls_alpha) { RP_CUDA(cudaMemcpyAsync(ls_alpha, c->lsa.p, nw * 4, cudaMemcpyDeviceToHost, s)); bytes += nw * 4; }
    if (ls_beta) { RP_CUDA(cudaMemcpyAsync(ls_beta, c->lsb.p, nw * 4, cudaMemcpyDeviceToHost, s)); bytes += nw * 4; }
    if (site_begin) { RP_CUDA(cudaMemcpyAsync(site_begin, c->sb.p, nw * 4, cudaMemcpyDeviceToHost, s)); bytes += nw * 4; }
    if (site_end) { RP_CUDA(cudaMemcpyAsync(site_end, c->se.p, nw * 4, cudaMemcpyDeviceToHost, s)); bytes += nw * 4; }
    RP_CUDA(cudaEventRecord(c->ev[1], s));
    RP_CUDA(cudaStreamSynchronize(s));
    if (st) {
        float a = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        st->ms_d2h += a;
        st->d2h_bytes += bytes;
    }
    return RP_OK;
}

} // namespace

// =========================================================================================
extern "C" {

const char *rp_last_error(void) { return g_err.c_str(); }

const char *rp_version(void) { return "relate_b200 paint 0.1 (sm_100a)"; }

int rp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rp_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(RP_EINVAL, "null argument");
    RP_CUDA(cudaMallocHost(out, bytes));
    return RP_OK;
}

void rp_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

int rp_chunk_create(int device, int N, int L, const char *hap, const double *r, const int *wb, int n_wb,
                    double theta, unsigned flags, rp_chunk **out)
{
    if (out) *out = nullptr;
    return chunk_from_host(device, N, L, hap, r, wb, n_wb, theta, flags, out, nullptr);
}

int rp_chunk_info(const rp_chunk *c, rp_info *info)
{
    if (!c || !info) return fail(RP_EINVAL, "null argument");
    info->N = c->N;
    info->L = c->L;
    info->W = c->W;
    info->device = c->device;
    info->words_per_snp = c->wps;
    info->hbm_bytes = (long long)(c->G.cap + c->GT.cap + c->r.cap + c->Phi.cap + c->Plo.cap + c->wbdev.cap);
    return RP_OK;
}

int rp_chunk_set_tune(rp_chunk *c, const rp_tune *t)
{
    if (!c) return fail(RP_EINVAL, "null chunk");
    rp_tune old = c->tune;
    c->tune = t ? *t : rp_tune{};
    LaunchPlan lp;
    int rc = plan_launch(c, lp);
    if (rc != RP_OK) c->tune = old;
    return rc;
}

int rp_chunk_set_stream(rp_chunk *c, void *cuda_stream)
{
    if (!c) return fail(RP_EINVAL, "null chunk");
    RP_CUDA(cudaSetDevice(c->device));
    RP_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return RP_OK;
}

void rp_chunk_free(rp_chunk *c)
{
    if (!c) return;
    for (rp_chunk **ref : c->window_refs) *ref = nullptr; // windows left open fail with RP_EINVAL instead of dangling
    cudaSetDevice(c->device);
    for (DevBuf *b : {&c->G, &c->GT, &c->r, &c->Phi, &c->Plo, &c->wbdev, &c->chars, &c->cnt_all, &c->counts, &c->off, &c->ent, &c->ia, &c->ib,
                      &c->lsA, &c->lsB, &c->sb, &c->se, &c->alpha, &c->beta, &c->lsa, &c->lsb, &c->queue, &c->scratch, &c->nor, &c->segstate, &c->segdone,
                      &c->rleK, &c->rec_off, &c->win_bytes, &c->img_off, &c->image, &c->image_alt, &c->park_top, &c->park_ls, &c->park_scal,
                      &c->park_d, &c->park_rowoff, &c->park_rpos, &c->park_lsa, &c->park_lsb})
        b->release();
    if (c->h_total) cudaFreeHost(c->h_total);
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int rp_paint_targets_device(rp_chunk *c, int k_begin, int k_end, const float **dev_alpha, const float **dev_beta,
                            const float **dev_ls_alpha, const float **dev_ls_beta, const int **dev_site_begin,
                            const int **dev_site_end, rp_stats *stats)
{
    const double t0 = now_ms();
    RP_TRY(paint_device(c, k_begin, k_end, stats));
    if (dev_alpha) *dev_alpha = c->alpha.as<float>();
    if (dev_beta) *dev_beta = c->beta.as<float>();
    if (dev_ls_alpha) *dev_ls_alpha = c->lsa.as<float>();
    if (dev_ls_beta) *dev_ls_beta = c->lsb.as<float>();
    if (dev_site_begin) *dev_site_begin = c->sb.as<int>();
    if (dev_site_end) *dev_site_end = c->se.as<int>();
    if (stats) stats->ms_total += now_ms() - t0;
    return RP_OK;
}

int rp_paint_targets(rp_chunk *c, int k_begin, int k_end, float *alpha, float *beta, float *ls_alpha,
                     float *ls_beta, int *site_begin, int *site_end, rp_stats *stats)
{
    const double t0 = now_ms();
    RP_TRY(paint_device(c, k_begin, k_end, stats));
    RP_TRY(copy_out(c, k_end - k_begin, alpha, beta, ls_alpha, ls_beta, site_begin, site_end, stats));
    if (stats) stats->ms_total += now_ms() - t0;
    return RP_OK;
}

int rp_paint_records(rp_chunk *c, int k_begin, int k_end, long long *win_off, rp_stats *stats)
{
    const double t0 = now_ms();
    if (c) c->enc_targets = 0;
    RP_TRY(paint_device(c, k_begin, k_end, stats));
    RP_TRY(encode_device(c, k_end - k_begin, stats));
    if (win_off) memcpy(win_off, c->h_img_off.data(), (size_t)(c->W + 1) * 8);
    if (stats) stats->ms_total += now_ms() - t0;
    return RP_OK;
}

int rp_records_copy(rp_chunk *c, long long offset, long long bytes, void *host_image, rp_stats *stats)
{
    if (!c || !host_image) return fail(RP_EINVAL, "null argument");
    if (c->enc_targets == 0) return fail(RP_EINVAL, "no encoded batch on this chunk (call rp_paint_records first)");
    if (offset < 0 || bytes < 0 || offset + bytes > c->h_img_off[c->W]) return fail(RP_EINVAL, "range outside the encoded images");
    RP_CUDA(cudaSetDevice(c->device));
    RP_CUDA(cudaEventRecord(c->ev[0], c->stream));
    RP_CUDA(cudaMemcpyAsync(host_image, c->image.as<char>() + offset, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    RP_CUDA(cudaEventRecord(c->ev[1], c->stream));
    RP_CUDA(cudaStreamSynchronize(c->stream));
    if (stats) {
        float a = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        stats->ms_d2h += a;
        stats->d2h_bytes += bytes;
    }
    return RP_OK;
}

int rp_paint_from_host(int device, int N, int L, const char *hap, const double *r, const int *wb, int n_wb,
                       double theta, unsigned flags, const rp_tune *tune, int k_begin, int k_end, float *alpha,
                       float *beta, float *ls_alpha, float *ls_beta, int *site_begin, int *site_end,
                       rp_stats *stats)
{
    const double t0 = now_ms();
    rp_chunk *c = nullptr;
    RP_TRY(chunk_from_host(device, N, L, hap, r, wb, n_wb, theta, flags, &c, stats));
    int rc = RP_OK;
    if (tune) rc = rp_chunk_set_tune(c, tune);
    if (rc == RP_OK) rc = paint_device(c, k_begin, k_end, stats);
    if (rc == RP_OK) rc = copy_out(c, k_end - k_begin, alpha, beta, ls_alpha, ls_beta, site_begin, site_end, stats);
    std::string keep = g_err;
    rp_chunk_free(c);
    g_err = keep;
    if (stats) stats->ms_total = now_ms() - t0;
    return rc;
}

int rp_make_chunks(const char *haps, const char *sample, const char *map, const char *dist, const char *out_dir,
                   int transversion, float memory_gb, int *n_chunks, char *warnings, size_t warnings_cap)
{
    const char *hb = getenv("RELATE_HAPBITS");
    return rp_make_chunks_ex(haps, sample, map, dist, out_dir, transversion, memory_gb, (hb && atoi(hb) != 0) ? RP_MC_HAPBITS : 0u,
                             n_chunks, warnings, warnings_cap);
}

int rp_make_chunks_ex(const char *haps, const char *sample, const char *map, const char *dist, const char *out_dir,
                      int transversion, float memory_gb, unsigned mc_flags, int *n_chunks, char *warnings, size_t warnings_cap)
{
    if (!haps || !sample || !map || !out_dir) return fail(RP_EINVAL, "Needed: haps, sample, map, output.");
    const std::string out = out_dir;
    struct stat sb;
    if (stat((out + "/").c_str(), &sb) == 0) // pipeline/MakeChunks.cpp:38-43
        return fail(RP_EINVAL, "Error: Directory " + out + " already exists. Relate will use this directory to store temporary files.");
    if (mkdir(out.c_str(), 0700) != 0) return fail(RP_EIO, "could not create directory " + out);
    rp::MakeChunksInfo info;
    const std::string err = rp::make_chunks(haps, sample, map, dist ? dist : "unspecified", out, transversion == 0, memory_gb, &info,
                                            (mc_flags & RP_MC_HAPBITS) != 0);
    if (!err.empty()) return fail(err.rfind("Failed to open", 0) == 0 || err.rfind("cannot", 0) == 0 ? RP_EIO : RP_EINVAL, err);
    if (n_chunks) *n_chunks = info.num_chunks;
    if (warnings && warnings_cap > 0) {
        const size_t n = std::min(warnings_cap - 1, info.warnings.size());
        memcpy(warnings, info.warnings.data(), n);
        warnings[n] = 0;
    }
    return RP_OK;
}

int rp_rle_encode(const float *v, int n, float *vals, int *lens)
{
    if (!v || !vals || !lens || n < 1) return fail(RP_EINVAL, "bad argument");
    return rp::rle_encode(v, n, vals, lens);
}

int rp_fast_log_device(int device, const float *in, float *out, int n)
{
    if (!in || !out || n < 1) return fail(RP_EINVAL, "bad argument");
    RP_CUDA(cudaSetDevice(device));
    float *d = nullptr;
    RP_CUDA(cudaMalloc(&d, (size_t)n * 8));
    cudaError_t e = cudaMemcpy(d, in, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rp::fast_log_kernel<<<(n + 255) / 256, 256>>>(d, d + n, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d + n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(RP_ECUDA, cudaGetErrorString(e));
    return RP_OK;
}

// FP32 issue-rate microbenchmark: the same instruction mix as the paint step (packed adds + scalar multiplies,
// no memory), to give the roofline a measured denominator on the box it runs on.
int rp_peak_fp32(int device, double *lane_ops_per_s_mix, double *lane_ops_per_s_scalar)
{
    RP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RP_CUDA(cudaGetDeviceProperties(&prop, device));
    float *d = nullptr;
    RP_CUDA(cudaMalloc(&d, 4096 * sizeof(float)));
    cudaEvent_t e0, e1;
    RP_CUDA(cudaEventCreate(&e0));
    RP_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, threads = 256, blocks = prop.multiProcessorCount * 8;
    double res[2] = {0, 0};
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            RP_CUDA(cudaEventRecord(e0));
            if (mode == 0) rp::peak_fp32_kernel<true><<<blocks, threads>>>(d, iters, 0.999f, 1e-3f);
            else rp::peak_fp32_kernel<false><<<blocks, threads>>>(d, iters, 0.999f, 1e-3f);
            RP_CUDA(cudaEventRecord(e1));
            RP_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            RP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        // per thread and iteration: 16 pairs x (add, mul, add) = 96 lane-ops
        res[mode] = (double)blocks * threads * iters * 96.0 / (best * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (lane_ops_per_s_mix) *lane_ops_per_s_mix = res[0];
    if (lane_ops_per_s_scalar) *lane_ops_per_s_scalar = res[1];
    return RP_OK;
}

int rp_debug_pack_host(int N, int L, const char *hap, uint32_t *snp_major, int words_per_snp)
{
    if (!hap || !snp_major || N < 1 || L < 1 || words_per_snp < (N + 31) / 32) return fail(RP_EINVAL, "bad argument");
    for (int s = 0; s < L; s++) pack_row_host(hap + (size_t)s * N, N, snp_major + (size_t)s * words_per_snp, words_per_snp);
    return RP_OK;
}

int rp_debug_pack(int device, int N, int L, const char *hap, uint32_t *snp_major, int *words_per_snp,
                  uint32_t *hap_major, int *words_per_hap)
{
    std::vector<double> r((size_t)L, 1e-4);
    int wb[2] = {0, L};
    rp_chunk *c = nullptr;
    RP_TRY(chunk_from_host(device, N, L, hap, r.data(), wb, 2, 0.001, 0, &c, nullptr));
    cudaError_t e = cudaSuccess;
    if (snp_major) e = cudaMemcpy(snp_major, c->G.p, (size_t)L * c->wps * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && hap_major) e = cudaMemcpy(hap_major, c->GT.p, (size_t)N * c->lw * 4, cudaMemcpyDeviceToHost);
    if (words_per_snp) *words_per_snp = c->wps;
    if (words_per_hap) *words_per_hap = c->lw;
    rp_chunk_free(c);
    if (e != cudaSuccess) return fail(RP_ECUDA, cudaGetErrorString(e));
    return RP_OK;
}

// ---- window repaint + distance matrices ("next" row f1) -------------------------------------------
struct rp_window {
    rp_chunk *c = nullptr;
    int device = 0;
    int w = 0, start = 0, end = 0;
    long long rows = 0;
    int pitch = 0, tt = 0, wpt = 0; // row layout of `top` (RepaintParams::pitch)
    DevBuf top, ls, scal, rowoff, rpos, d, ab, be, lsa, lsb;
};

// alpha == nullptr: take the stepping stones of window w from the chunk's HBM buffers (last paint covered all targets)
// and apply the codec's collapse (every element becomes the head of its run) on the device.
static int window_open_impl(rp_chunk *c, int w, const float *alpha, const float *beta, const float *ls_alpha,
                            const float *ls_beta, const double *rpos, rp_window **out, rp_stats *stats)
{
    const bool resident = alpha == nullptr;
    if (!c || !rpos || !out) return fail(RP_EINVAL, "null argument");
    if (!resident && (!beta || !ls_alpha || !ls_beta)) return fail(RP_EINVAL, "null argument");
    if (resident && !c->resident_all)
        return fail(RP_EINVAL, "the chunk does not hold the stepping stones of all targets (paint targets 0..N first)");
    if (w < 0 || w >= c->W) return fail(RP_EINVAL, "window index out of range");
    if (c->flags & RP_FP64) return fail(RP_EUNSUPPORTED, "window repaint runs with fp32 state");
    *out = nullptr;
    const double t0 = now_ms();
    const int N = c->N, W = c->W, L = c->L;
    RP_CUDA(cudaSetDevice(c->device));
    RP_TRY(paint_device(c, 0, N, stats, /*run_paint=*/false, /*want_nor=*/true));
    LaunchPlan lp;
    RP_TRY(plan_launch(c, lp));
    if (lp.cluster > 1 || (lp.multi && lp.threads > (lp.wpt == 1 ? 512 : 256)))
        return fail(RP_EUNSUPPORTED, "N too large for the window repaint kernel");
    cudaStream_t s = c->stream;
    // rows per target = ib - ia + 1; prefix on the host (N ints each way)
    std::vector<int> ia((size_t)N * W), ib((size_t)N * W);
    RP_CUDA(cudaMemcpyAsync(ia.data(), c->ia.p, ia.size() * 4, cudaMemcpyDeviceToHost, s));
    RP_CUDA(cudaMemcpyAsync(ib.data(), c->ib.p, ib.size() * 4, cudaMemcpyDeviceToHost, s));
    RP_CUDA(cudaStreamSynchronize(s));
    std::vector<long long> rowoff((size_t)N + 1, 0);
    for (int k = 0; k < N; k++) rowoff[k + 1] = rowoff[k] + (ib[(size_t)k * W + w] - ia[(size_t)k * W + w] + 1);
    rp_window *win = new rp_window();
    win->c = c;
    win->device = c->device;
    win->w = w;
    win->start = c->wb[w];
    win->end = (w < W - 1) ? c->wb[w + 1] - 1 : L - 1;
    win->rows = rowoff[N];
    auto bail = [&](int rc) {
        std::string keep = g_err;
        rp_window_close(win);
        g_err = keep;
        return rc;
    };
    int rc = RP_OK;
    const size_t nn = (size_t)N * N;
    std::swap(win->top, c->park_top); // (the buffers of the window closed last, if any)
    std::swap(win->ls, c->park_ls);
    std::swap(win->scal, c->park_scal);
    std::swap(win->d, c->park_d);
    std::swap(win->rowoff, c->park_rowoff);
    std::swap(win->rpos, c->park_rpos);
    std::swap(win->lsa, c->park_lsa);
    std::swap(win->lsb, c->park_lsb);
    win->tt = lp.threads;
    win->wpt = lp.wpt;
    win->pitch = lp.threads * lp.wpt * 32 + 32;
    if ((rc = win->top.ensure((size_t)win->rows * win->pitch * 4)) || (rc = win->ls.ensure((size_t)win->rows * 4)) ||
        (rc = win->scal.ensure((size_t)win->rows * 8)) ||
        (rc = win->rowoff.ensure(((size_t)N + 1) * 8)) || (rc = win->rpos.ensure(((size_t)L + 1) * 8)) ||
        (rc = win->d.ensure(nn * 4)) || (rc = win->ab.ensure(nn * 4)) || (rc = win->be.ensure(nn * 4)) ||
        (rc = win->lsa.ensure((size_t)N * 4)) || (rc = win->lsb.ensure((size_t)N * 4)))
        return bail(rc);
#define RP_CUDAW(call)                                                                                          \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) return bail(fail(RP_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); \
    } while (0)
    RP_CUDAW(cudaEventRecord(c->ev[0], s));
    RP_CUDAW(cudaMemcpyAsync(win->rowoff.p, rowoff.data(), rowoff.size() * 8, cudaMemcpyHostToDevice, s));
    RP_CUDAW(cudaMemcpyAsync(win->rpos.p, rpos, ((size_t)L + 1) * 8, cudaMemcpyHostToDevice, s));
    if (!resident) {
        RP_CUDAW(cudaMemcpyAsync(win->ab.p, alpha, nn * 4, cudaMemcpyHostToDevice, s));
        RP_CUDAW(cudaMemcpyAsync(win->be.p, beta, nn * 4, cudaMemcpyHostToDevice, s));
        RP_CUDAW(cudaMemcpyAsync(win->lsa.p, ls_alpha, (size_t)N * 4, cudaMemcpyHostToDevice, s));
        RP_CUDAW(cudaMemcpyAsync(win->lsb.p, ls_beta, (size_t)N * 4, cudaMemcpyHostToDevice, s));
    } else {
        const unsigned grid = (unsigned)std::min<long long>(((long long)2 * N + 7) / 8, (long long)c->sm_count * 64);
        rp::collapse_kernel<<<grid, 256, 0, s>>>(c->alpha.as<float>(), c->beta.as<float>(), c->lsa.as<float>(), c->lsb.as<float>(), N, W, w,
                                                 win->ab.as<float>(), win->be.as<float>(), win->lsa.as<float>(), win->lsb.as<float>());
        RP_CUDAW(cudaGetLastError());
    }
    RP_CUDAW(cudaMemsetAsync(c->queue.p, 0, 16, s));
    RP_CUDAW(cudaEventRecord(c->ev[1], s));
    rp::RepaintParams P{};
    P.G = c->G.as<uint32_t>();
    P.wps = c->wps; P.N = N; P.L = L; P.W = W; P.nfw = c->nfw; P.tailn = c->tailn;
    P.w = w; P.nt = N;
    P.ent = c->ent.as<char>() + 4 * sizeof(rp::EntF);
    P.off = c->off.as<long long>();
    P.nor = c->nor.as<double>();
    P.r = c->r.as<double>();
    P.ia = c->ia.as<int>(); P.ib = c->ib.as<int>();
    P.alpha_begin = win->ab.as<float>(); P.beta_end = win->be.as<float>();
    P.ls_alpha = win->lsa.as<float>(); P.ls_beta = win->lsb.as<float>();
    P.top = win->top.as<float>(); P.ls = win->ls.as<float>();
    P.scal = win->scal.as<float2>();
    P.pitch = win->pitch;
    P.rowoff = win->rowoff.as<long long>();
    P.queue = c->queue.as<int>();
    const double ntheta = 1.0 - c->theta;
    const double theta_ratio = c->theta / (1.0 - c->theta) - 1.0;
    P.cf.tau = (float)(1.0 * theta_ratio + 1.0);
    P.cf.prior_n = (float)(ntheta / (N - 1.0));
    P.cf.ntheta = (float)ntheta;
    P.cf.inv_ntheta = (float)(1.0 / ntheta);
    P.cf.lower = (float)1e-10;
    P.cf.upper = (float)(1.0 / 1e-10);
    P.log_ntheta = log(ntheta); P.log_small = log(0.01); P.Nm1 = N - 1.0;
    int occ = 0, ctas = 0;
    auto launch = [&](auto kern, size_t smem) -> int { // smem: checkpoint ring + recomputed rows of one team
        RP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, lp.threads, smem));
        if (occ < 1) return fail(RP_ECUDA, "repaint kernel does not fit on an SM");
        ctas = std::min(N, occ * c->sm_count);
        kern<<<ctas, lp.threads, smem, s>>>(P);
        RP_CUDA(cudaGetLastError());
        return RP_OK;
    };
    // checkpoint spacing: 4 rows for one-word single-warp teams (HBM traffic 12N -> 6N bytes per row), 2 for two-word and
    // multi-warp teams (registers; rows of tens of KB of shared memory each): 12N -> 8N.  RP_REPAINT_CK=8 selects 8 for one-word single-warp teams.
    const int ck_env = getenv("RP_REPAINT_CK") ? atoi(getenv("RP_REPAINT_CK")) : 0;
    const bool ck8 = ck_env == 8;
    if (lp.wpt == 1) {
        if (lp.multi) rc = launch(rp::repaint_kernel<1, true, 2>, rp::RepaintSmem<2>::bytes(lp.threads, 1));
        else if (ck8) rc = launch(rp::repaint_kernel<1, false, 8>, rp::RepaintSmem<8>::bytes(32, 1));
        else if (ck_env == 2) rc = launch(rp::repaint_kernel<1, false, 2>, rp::RepaintSmem<2>::bytes(32, 1));
        else if (ck_env == 3) rc = launch(rp::repaint_kernel<1, false, 3>, rp::RepaintSmem<3>::bytes(32, 1));
        else rc = launch(rp::repaint_kernel<1, false, 4>, rp::RepaintSmem<4>::bytes(32, 1));
    } else {
        if (lp.multi) rc = launch(rp::repaint_kernel<2, true, 2>, rp::RepaintSmem<2>::bytes(lp.threads, 2));
        else rc = launch(rp::repaint_kernel<2, false, 2>, rp::RepaintSmem<2>::bytes(32, 2)); // (4 rows spill: 2 x 64 state registers)
    }
    if (rc != RP_OK) return bail(rc);
    RP_CUDAW(cudaEventRecord(c->ev[2], s));
    RP_CUDAW(cudaStreamSynchronize(s));
    for (DevBuf *b : {&win->ab, &win->be}) b->release();
    if (stats) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
        stats->ms_h2d += a;
        stats->ms_paint += b; // the repaint kernel
        stats->h2d_bytes += (long long)((resident ? 0 : 2 * nn * 4) + ((size_t)L + 1) * 8);
        stats->launches += resident ? 2 : 1;
        stats->n_targets = N;
        stats->team_threads = lp.threads;
        stats->words_per_thread = lp.wpt;
        stats->ctas = ctas;
        stats->cells = win->rows; // rows of the posterior held in HBM
        stats->ms_total += now_ms() - t0;
    }
    c->window_refs.push_back(&win->c);
    *out = win;
    return RP_OK;
#undef RP_CUDAW
}

int rp_window_open(rp_chunk *c, int w, const float *alpha, const float *beta, const float *ls_alpha,
                   const float *ls_beta, const double *rpos, rp_window **out, rp_stats *stats)
{
    if (!alpha) return fail(RP_EINVAL, "null argument");
    return window_open_impl(c, w, alpha, beta, ls_alpha, ls_beta, rpos, out, stats);
}

int rp_window_open_resident(rp_chunk *c, int w, const double *rpos, rp_window **out, rp_stats *stats)
{
    return window_open_impl(c, w, nullptr, nullptr, nullptr, nullptr, rpos, out, stats);
}

int rp_window_open_files(rp_chunk *c, const char *out_dir, int chunk_index, int w, rp_window **out, rp_stats *stats)
{
    if (!c || !out_dir || !out) return fail(RP_EINVAL, "null argument");
    if (w < 0 || w >= c->W) return fail(RP_EINVAL, "window index out of range");
    const int N = c->N, L = c->L;
    const std::string base = std::string(out_dir) + "/chunk_" + std::to_string(chunk_index);
    std::vector<double> rpos((size_t)L + 1);
    {
        FILE *fp = fopen((base + ".rpos").c_str(), "rb");
        if (!fp) return fail(RP_EIO, "cannot open " + base + ".rpos");
        unsigned n = 0;
        bool ok = fread(&n, 4, 1, fp) == 1 && (int)n == L + 1 && fread(rpos.data(), 8, (size_t)L + 1, fp) == (size_t)L + 1;
        fclose(fp);
        if (!ok) return fail(RP_EIO, "short read in " + base + ".rpos");
    }
    std::vector<float> alpha((size_t)N * N), beta((size_t)N * N), lsa(N), lsb(N);
    const std::string pf = base + "/paint/relate_" + std::to_string(w) + ".bin";
    std::string err = rp::decode_paint_file(pf, N, c->wb[w], c->wb[w + 1] - 1, alpha.data(), beta.data(), lsa.data(), lsb.data());
    if (!err.empty()) return fail(RP_EIO, err);
    return rp_window_open(c, w, alpha.data(), beta.data(), lsa.data(), lsb.data(), rpos.data(), out, stats);
}

int rp_window_distance(rp_window *win, int snp, float *d)
{
    if (!win || !d) return fail(RP_EINVAL, "null argument");
    if (snp < win->start || snp > win->end) return fail(RP_EINVAL, "snp outside the window");
    rp_chunk *c = win->c;
    if (!c) return fail(RP_EINVAL, "the window's chunk has been freed");
    RP_CUDA(cudaSetDevice(c->device));
    rp::DistanceParams P{};
    P.GT = c->GT.as<uint32_t>();
    P.lw = c->lw; P.N = c->N; P.L = c->L; P.W = c->W; P.w = win->w; P.start = win->start; P.snp = snp;
    P.rpos = win->rpos.as<double>();
    P.top = win->top.as<float>(); P.ls = win->ls.as<float>();
    P.rowoff = win->rowoff.as<long long>();
    P.d = win->d.as<float>();
    P.pitch = win->pitch; P.tt = win->tt; P.wpt = win->wpt;
    rp::distance_kernel<<<c->N, 256, 0, c->stream>>>(P);
    RP_CUDA(cudaGetLastError());
    RP_CUDA(cudaMemcpyAsync(d, win->d.p, (size_t)c->N * c->N * 4, cudaMemcpyDeviceToHost, c->stream));
    RP_CUDA(cudaStreamSynchronize(c->stream));
    return RP_OK;
}

long long rp_window_rows(const rp_window *win) { return win ? win->rows : 0; }

void rp_window_close(rp_window *win)
{
    if (!win) return;
    if (win->c) {
        rp_chunk *c = win->c;
        cudaSetDevice(c->device);
        auto &refs = c->window_refs;
        refs.erase(std::remove(refs.begin(), refs.end(), &win->c), refs.end());
        if (!c->park_top.p) { // park this window's buffers in the chunk for the next window
            std::swap(win->top, c->park_top);
            std::swap(win->ls, c->park_ls);
            std::swap(win->scal, c->park_scal);
            std::swap(win->d, c->park_d);
            std::swap(win->rowoff, c->park_rowoff);
            std::swap(win->rpos, c->park_rpos);
            std::swap(win->lsa, c->park_lsa);
            std::swap(win->lsb, c->park_lsb);
        }
    } else {
        cudaSetDevice(win->device);
    }
    for (DevBuf *b : {&win->top, &win->ls, &win->scal, &win->rowoff, &win->rpos, &win->d, &win->ab, &win->be, &win->lsa, &win->lsb}) b->release();
    delete win;
}

} // extern "C"

// ---- the whole stage (pipeline/Paint.cpp:17-108) ----------------------------------------
// chunk files -> reader threads bit-pack into a pinned ring -> every GPU's HBM (replica, own H2D, no collective) ->
// paint in batches of targets pulled from one counter (dynamic balance over GPUs) -> records encoded on the device
// (rle_kernel) -> pinned pieces -> pwrite at absolute file offsets (a batch's offsets are fixed as soon as the sizes
// of all earlier batches are published, so pieces, batches and devices reach the files in any order).
// Device buffers, pinned staging and streams are parked in a per-device cache between calls
// (rp_release_cache() frees them), so a process painting many chunks pays for them once.
namespace {

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return RP_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(RP_ENOMEM, "cudaHostAlloc of " + std::to_string(bytes) + " bytes failed");
        }
        cap = bytes;
        return RP_OK;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct DeviceWorkspace { // parked between rp_paint_chunk calls
    rp_chunk *shell = nullptr; // keeps its DevBufs, stream and events
    PinnedBuf out;             // ring of pinned pieces the encoded records pass through on their way to the files
    PinnedBuf hap_in;          // genotype bytes of the chunk being painted with this device as devs[0]
};

// Pinned output staging is a small ring of pieces (pinning costs ~0.5 s/GB, which dominated a one-shot
// `relate --mode Paint`): a batch's W file images are copied in pieces taken round-robin from the W windows (buffered
// writes to one file serialise on its inode lock, so the pieces in flight should belong to different files), every
// piece is one pwrite task at an absolute file offset (the order in which pieces, batches and devices reach the disk
// does not matter), and a piece returns to the ring when its task is done.
constexpr size_t kOutPiece = (size_t)4 << 20;
constexpr int kOutPieces = 24;

struct OutRing {
    char *base = nullptr;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int> free_slots;
    std::atomic<int> pending[kOutPieces];
    void init(char *p)
    {
        base = p;
        free_slots.clear();
        for (int i = 0; i < kOutPieces; i++) {
            free_slots.push_back(i);
            pending[i].store(0);
        }
    }
    int acquire()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !free_slots.empty(); });
        const int sl = free_slots.back();
        free_slots.pop_back();
        return sl;
    }
    void task_done(int sl)
    {
        if (pending[sl].fetch_sub(1) == 1) {
            std::lock_guard<std::mutex> lk(mu);
            free_slots.push_back(sl);
            cv.notify_one();
        }
    }
    void wait_all()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return (int)free_slots.size() == kOutPieces; });
    }
};

struct WriteTask {
    const char *src;
    size_t len;
    int fd;
    long long off;
    OutRing *ring;
    int slot;
};

class WritePool {
  public:
    explicit WritePool(int nthreads)
    {
        for (int i = 0; i < nthreads; i++) th_.emplace_back([this] { run(); });
    }
    ~WritePool() { finish(); }
    void push(const WriteTask &t)
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            if (t_first_ == 0) t_first_ = now_ms();
            q_.push_back(t);
        }
        cv_.notify_one();
    }
    void finish()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_)
            if (t.joinable()) t.join();
        th_.clear();
    }
    bool failed() const { return err_.load() != 0; }
    double span_ms() const { return t_first_ == 0 ? 0.0 : t_last_ - t_first_; }

  private:
    void run()
    {
        for (;;) {
            WriteTask t;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                t = q_.front();
                q_.pop_front();
            }
            const char *p = t.src;
            size_t left = t.len;
            long long off = t.off;
            while (left > 0) {
                const ssize_t put = pwrite(t.fd, p, left, (off_t)off);
                if (put <= 0) {
                    err_ = 1;
                    break;
                }
                p += put;
                off += put;
                left -= (size_t)put;
            }
            {
                std::lock_guard<std::mutex> lk(m_);
                t_last_ = now_ms();
            }
            t.ring->task_done(t.slot);
        }
    }
    std::vector<std::thread> th_;
    std::deque<WriteTask> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
    std::atomic<int> err_{0};
    double t_first_ = 0, t_last_ = 0;
};


std::mutex g_stage_mu; // rp_paint_chunk calls are serialised (they share the cache)
std::map<int, DeviceWorkspace> g_ws;
std::vector<rp_stats> g_last_dstats; // per-device statistics of the last rp_paint_chunk call (rp_stage_device_stats)

// Host threads this process may use for file I/O (chunk readers + paint-file writers).  Several processes painting at
// once on one host (one rank per GPU) must share the cores: RP_IO_THREADS sets the budget; otherwise it is the core
// count divided by LOCAL_WORLD_SIZE (set by torchrun) when that is present.
unsigned io_thread_budget()
{
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (const char *e = getenv("RP_IO_THREADS")) return (unsigned)std::max(1, atoi(e));
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) return std::max(2u, hw / (unsigned)std::max(1, atoi(e)));
    return hw;
}

template <typename F> void parallel_for(int n, int nthreads, F f)
{
    std::atomic<int> next{0};
    auto body = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) break;
            f(i);
        }
    };
    nthreads = std::max(1, std::min(nthreads, n));
    std::vector<std::thread> ts;
    for (int i = 1; i < nthreads; i++) ts.emplace_back(body);
    body();
    for (auto &t : ts) t.join();
}

// Reader threads that deliver a chunk's genotype rows, bit-packed, through a ring of pinned slots (see HapFeed): from
// chunk_<c>.hapbits when this library's MakeChunks left one (the rows already in the painter's layout: 1/8 of the bytes
// and nothing to pack), else from the chars of chunk_<c>.hap, packed by the thread that read them.  Used by the stage
// driver (several devices copy from one ring) and by rp_chunk_load.
struct ChunkReaders {
    HapFeed feed;
    std::vector<std::thread> threads;
    std::atomic<int> next_slice{0}, readers_left{0};
    int hap_fd = -1, bits_fd = -1;
    bool have_sidecar = false; // the rows come from chunk_<c>.hapbits
    std::string bits_path;
    double t_loaded = 0;
    int N = 0, L = 0, wps = 0;

    // takes ownership of hap_fd (the open chunk_<c>.hap, header already checked)
    int start(const char *out_dir, int chunk_index, const rp::HostChunk &hc, int hap_fd_in, int ndev, PinnedBuf &ring)
    {
        hap_fd = hap_fd_in;
        N = hc.N;
        L = hc.L;
        wps = (((N + 31) / 32) + 3) / 4 * 4; // as rp_chunk::wps
        t_loaded = now_ms();
        bits_path = std::string(out_dir) + "/chunk_" + std::to_string(chunk_index) + ".hapbits";
        bits_fd = open(bits_path.c_str(), O_RDONLY);
        if (bits_fd >= 0) {
            rp::HapBitsHeader h{};
            struct stat sb;
            const bool ok = pread(bits_fd, &h, sizeof h, 0) == (ssize_t)sizeof h && memcmp(h.magic, rp::hapbits_magic(), 8) == 0 && h.N == N &&
                            h.L == L && h.wps == wps && fstat(bits_fd, &sb) == 0 && (size_t)sb.st_size >= sizeof h + (size_t)h.L * h.wps * 4;
            if (!ok) {
                close(bits_fd);
                bits_fd = -1;
            }
        }
        have_sidecar = bits_fd >= 0;
        const size_t nchar = (size_t)L * N;
        const size_t ring_cap = (size_t)(getenv("RP_RING_KB") ? atoi(getenv("RP_RING_KB")) : 65536) << 10; // tests shrink it
        // a slice = whole SNP rows worth about RP_SLICE_KB of genotype chars on disk; in the ring it is 1/8 of that
        const size_t slice_chars = (size_t)(getenv("RP_SLICE_KB") ? atoi(getenv("RP_SLICE_KB")) : (nchar <= 8 * ring_cap ? 1024 : 4096)) << 10;
        feed.rows = (int)std::max<size_t>(1, slice_chars / (size_t)N);
        feed.slice = (size_t)feed.rows * wps * 4;
        feed.nsl = (L + feed.rows - 1) / feed.rows;
        feed.nslots = (int)std::min<size_t>((size_t)feed.nsl, std::max<size_t>(2, ring_cap / feed.slice));
        feed.ndev = ndev;
        feed.padbit = hc.theta > 0.5 ? 1 : 0; // as chunk_from_host sets rp_chunk::padbit
        if (ring.ensure((size_t)feed.nslots * feed.slice) != RP_OK) {
            close_files();
            return RP_ENOMEM;
        }
        feed.ring = static_cast<const char *>(ring.p);
        feed.ready.reset(new std::atomic<int>[feed.nsl]);
        feed.consumed.reset(new std::atomic<int>[feed.nsl]);
        for (int i = 0; i < feed.nsl; i++) {
            feed.ready[i].store(0);
            feed.consumed[i].store(0);
        }
        const unsigned want_readers = getenv("RP_READERS") ? (unsigned)atoi(getenv("RP_READERS")) : 8u;
        // (a chunk is read, painted and written one after the other, so readers and writers may each use the whole budget)
        const int nread = (int)std::max(1u, std::min<unsigned>({want_readers, io_thread_budget(), (unsigned)feed.nslots}));
        readers_left = nread;
        for (int t = 0; t < nread; t++) threads.emplace_back([this]() { run(); });
        return RP_OK;
    }
    void run()
    {
        std::vector<char> raw; // the slice as it is on disk (pageable: it never meets the GPU)
        for (;;) {
            const int i = next_slice.fetch_add(1);
            if (i >= feed.nsl) break;
            bool ok = true;
            if (i >= feed.nslots) // wait until every device has copied the slice that occupies the slot
                while (feed.consumed[i - feed.nslots].load(std::memory_order_acquire) < feed.ndev) {
                    if (feed.abort.load()) { ok = false; break; }
                    std::this_thread::yield();
                }
            const int row0 = i * feed.rows, nrows = std::min(feed.rows, L - row0);
            uint32_t *dst = reinterpret_cast<uint32_t *>(const_cast<char *>(feed.src(i)));
            if (bits_fd >= 0) { // packed rows straight from the sidecar
                char *d8 = reinterpret_cast<char *>(dst);
                size_t left = (size_t)nrows * wps * 4, off = sizeof(rp::HapBitsHeader) + (size_t)row0 * wps * 4;
                while (ok && left > 0) {
                    const ssize_t got = pread(bits_fd, d8, left, (off_t)off);
                    if (got <= 0) ok = false;
                    else { d8 += got; off += (size_t)got; left -= (size_t)got; }
                }
                if (ok && feed.padbit && (N & 31)) // phantom slots of the partial last word (tau > 1)
                    for (int rr = 0; rr < nrows; rr++) dst[(size_t)rr * wps + (N >> 5)] |= ~0u << (N & 31);
            } else {
                raw.resize((size_t)feed.rows * N);
                ok = ok && rp::read_hap_range(hap_fd, (size_t)row0 * N, (size_t)nrows * N, raw.data());
                if (ok)
                    for (int rr = 0; rr < nrows; rr++) pack_row_host(raw.data() + (size_t)rr * N, N, dst + (size_t)rr * wps, wps, feed.padbit);
            }
            feed.ready[i].store(ok ? 1 : -1, std::memory_order_release);
        }
        if (readers_left.fetch_sub(1) == 1) t_loaded = now_ms();
    }
    void close_files()
    {
        if (hap_fd >= 0) close(hap_fd);
        if (bits_fd >= 0) close(bits_fd);
        hap_fd = bits_fd = -1;
    }
    void join()
    {
        for (auto &t : threads) t.join();
        threads.clear();
        close_files();
    }
    // the sidecar has served its purpose (the reference's Finalize cannot delete a directory that holds a file it does
    // not know); call after join()
    void remove_sidecar()
    {
        if (have_sidecar) unlink(bits_path.c_str());
        have_sidecar = false;
    }
    ~ChunkReaders()
    {
        feed.abort.store(1);
        join();
    }
};

} // namespace

extern "C" void rp_release_cache(void)
{
    std::lock_guard<std::mutex> lk(g_stage_mu);
    for (auto &kv : g_ws) {
        cudaSetDevice(kv.first);
        if (kv.second.shell) rp_chunk_free(kv.second.shell);
        kv.second.out.release();
        kv.second.hap_in.release();
    }
    g_ws.clear();
    g_pool.flush(); // device memory parked by freed chunks and windows goes back to the driver
}

namespace {

int check_devices(const int *devices, int n_devices, std::vector<int> &devs)
{
    int ndev = rp_device_count();
    if (ndev < 1) return fail(RP_ENODEVICE, "no CUDA device (there is no CPU fallback)");
    devs.clear();
    if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
    else devs.push_back(0);
    for (size_t i = 0; i < devs.size(); i++) {
        if (devs[i] < 0 || devs[i] >= ndev) return fail(RP_EINVAL, "device index out of range");
        // two workers on one device would share its parked workspace (shell chunk, pinned output ring) and corrupt the files
        for (size_t j = 0; j < i; j++)
            if (devs[j] == devs[i]) return fail(RP_EINVAL, "device " + std::to_string(devs[i]) + " listed twice");
    }
    return RP_OK;
}

// Sizes the per-batch work buffers of a chunk once, for batches of B targets: growing one of them later means cudaFree +
// cudaMalloc, which synchronises the device and so serialises a batch's painting with the previous batch's copies.
int reserve_for_batches(rp_chunk *c, int B)
{
    const int N = c->N, W = c->W;
    RP_CUDA(cudaSetDevice(c->device));
    long long umax = 0;
    for (int k0 = 0; k0 < N; k0 += B) umax = std::max(umax, c->site_prefix[(size_t)std::min(N, k0 + B)] - c->site_prefix[(size_t)k0]);
    const bool fp64 = (c->flags & RP_FP64) != 0;
    const size_t entsz = fp64 ? sizeof(rp::EntD) : sizeof(rp::EntF);
    const size_t nw = (size_t)std::min(B, N) * W;
    RP_TRY(c->ent.ensure(((size_t)umax + 8) * entsz));
    RP_TRY(c->off.ensure(((size_t)B + 1) * 8));
    for (DevBuf *b : {&c->ia, &c->ib, &c->sb, &c->se, &c->lsa, &c->lsb}) RP_TRY(b->ensure(nw * 4));
    for (DevBuf *b : {&c->lsA, &c->lsB}) RP_TRY(b->ensure(nw * 8));
    RP_TRY(c->alpha.ensure(nw * N * 4));
    RP_TRY(c->beta.ensure(nw * N * 4));
    RP_TRY(c->rleK.ensure(nw * 2 * 4));
    RP_TRY(c->rec_off.ensure(nw * 8));
    // record images: 64 + 8*(Ka+Kb) bytes per (target, window); Ka+Kb is typically 0.6-1.2 N (a larger batch grows the buffer)
    const size_t img = nw * (64 + (size_t)10 * N);
    RP_TRY(c->image.ensure(img));
    RP_TRY(c->image_alt.ensure(img));
    return RP_OK;
}

// One chunk on the given devices.  The caller holds g_stage_mu and has created the g_ws entries of `devs`; concurrent
// calls must use disjoint device sets (rp_paint_chunks: one device each).
int paint_chunk_stage(const char *out_dir, int chunk_index, const char *painting, const std::vector<int> &devs, unsigned flags,
                      rp_stats *stats, std::vector<rp_stats> *per_device = nullptr)
{
    const double t0 = now_ms();

    // ---- input: small files now, genotype bytes by reader threads while the devices already copy ----
    rp::HostChunk hc;
    int hap_fd = -1;
    trace("stage begin");
    RP_CUDA(cudaSetDevice(devs[0]));
    trace("cudaSetDevice done", devs[0]);
    {
        std::string err = rp::load_chunk_small(out_dir, chunk_index, painting, hc, &hap_fd);
        if (!err.empty()) return fail(RP_EIO, err);
    }
    const unsigned hw = io_thread_budget();
    ChunkReaders in;
    {
        const int rc = in.start(out_dir, chunk_index, hc, hap_fd, (int)devs.size(), g_ws.at(devs[0]).hap_in);
        if (rc != RP_OK) return rc;
    }
    HapFeed &feed = in.feed;
    hc.hap = const_cast<char *>(feed.ring); // (only slices in the ring are addressable through it)
    trace("input ring pinned", devs[0]);

    // ---- output files: created (and old ones truncated) in the background; the first write waits for it ----
    const int N = hc.N, W = (int)hc.wb.size() - 1;
    std::vector<int> fds(W, -1);
    std::shared_future<std::string> files_open = std::async(std::launch::async, [&]() -> std::string {
        const std::string cdir = std::string(out_dir) + "/chunk_" + std::to_string(chunk_index);
        const std::string pdir = cdir + "/paint";
        // filesys::MakeDir semantics (src/filesystem.cpp:4-24): create if absent, mode 0700
        for (const std::string &d : {cdir, pdir}) {
            struct stat sb;
            if (stat(d.c_str(), &sb) != 0 && mkdir(d.c_str(), 0700) != 0) return "could not create directory " + d;
        }
        std::atomic<int> bad{-1};
        parallel_for(W, 8, [&](int w) {
            const std::string p = pdir + "/relate_" + std::to_string(w) + ".bin";
            fds[w] = open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
            if (fds[w] < 0) bad = w;
        });
        if (bad >= 0) return "cannot create " + pdir + "/relate_" + std::to_string(bad.load()) + ".bin";
        return "";
    }).share();

    // Batch size.  A batch is one launch of the paint kernel on one device: it should hold a few waves of chains (a
    // multi-warp team is one CTA: 444-740 resident teams per device; 63 batches of 160 targets left config 4 at 0.45 of the
    // nominal issue rate against 0.57 for a well filled grid), there should be up to 4 batches per device so that copying
    // and writing one batch overlaps painting the next, and its stepping stones (2*W*N floats per target) plus the two
    // record images should not take more than ~24 GB of HBM.
    const size_t per_target = (size_t)2 * W * N * 4;
    const long long ndev_ll = (long long)devs.size();
    const long long fill = N > 2048 ? 700 : 1000; // targets for ~3 waves of teams (two chains per target)
    const long long per_dev = std::max<long long>(1, std::min<long long>(4, N / (ndev_ll * fill)));
    long long bsz = (N + ndev_ll * per_dev - 1) / (ndev_ll * per_dev);
    bsz = std::min<long long>(bsz, (long long)((8ull << 30) / per_target));
    bsz = std::max<long long>(bsz, 64);
    bsz = std::min<long long>(bsz, N);
    if (const char *fb = getenv("RP_BATCH_TARGETS")) bsz = std::max(1, std::min(N, atoi(fb))); // tests: force small batches
    const int B = (int)bsz;
    const int nbatch = (N + B - 1) / B;

    std::atomic<int> next_batch{0};
    std::mutex mu;
    std::condition_variable cv;
    int first_rc = RP_OK;
    std::string first_err;
    std::vector<rp_stats> dstats(devs.size());
    auto set_error = [&](int rc, const std::string &msg) { // call with mu held
        if (first_rc == RP_OK) {
            first_rc = rc;
            first_err = msg;
        }
        cv.notify_all();
    };
    // Where batch b's records start in file w = the sizes of all earlier batches' records: published as soon as a batch
    // has been encoded, so the data itself can travel in any order.
    std::vector<std::vector<long long>> bsize(nbatch), bbase(nbatch);
    std::vector<char> have(nbatch, 0);
    std::vector<long long> running(W, 0);
    int upto = 0;
    auto publish_sizes = [&](int b, const std::vector<long long> &off) -> bool { // false if the stage has failed
        std::unique_lock<std::mutex> lk(mu);
        bsize[b].resize(W);
        for (int w = 0; w < W; w++) bsize[b][w] = off[w + 1] - off[w];
        have[b] = 1;
        while (upto < nbatch && have[upto]) {
            bbase[upto] = running;
            for (int w = 0; w < W; w++) running[w] += bsize[upto][w];
            upto++;
        }
        cv.notify_all();
        cv.wait(lk, [&] { return upto > b || first_rc != RP_OK; });
        return first_rc == RP_OK;
    };
    // writers: each pwrite is one thread's copy into the page cache (a few GB/s); pieces in flight belong to different files
    const unsigned want_writers = getenv("RP_WRITERS") ? (unsigned)atoi(getenv("RP_WRITERS")) : 16u;
    WritePool pool((int)std::max(2u, std::min(want_writers, hw)));
    std::vector<std::unique_ptr<OutRing>> rings(devs.size());

    auto worker = [&](int di) {
        rp_stats &st = dstats[di];
        memset(&st, 0, sizeof st);
        DeviceWorkspace &ws = g_ws.at(devs[di]);
        rp_chunk *c = ws.shell;
        ws.shell = nullptr;
        trace("worker start", devs[di]);
        int rc = chunk_from_host(devs[di], hc.N, hc.L, hc.hap, hc.r.data(), hc.wb.data(), (int)hc.wb.size(), hc.theta,
                                 flags, &c, &st, &feed);
        trace("chunk resident (H2D + bit-pack)", devs[di]);
        if (rc != RP_OK) feed.abort.store(1); // readers must not wait for this device's copies
        if (rc == RP_OK) rc = reserve_for_batches(c, B);
        if (rc == RP_OK) rc = ws.out.ensure(kOutPiece * kOutPieces);
        rings[di].reset(new OutRing());
        OutRing &ring = *rings[di];
        ring.init(static_cast<char *>(ws.out.p)); // (a null base is never used: rc != RP_OK skips the batches)
        cudaEvent_t pev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; // one per copy kept in flight (+1)
        for (auto &ev : pev)
            if (rc == RP_OK && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) rc = fail(RP_ECUDA, "cudaEventCreate failed");
        trace("output ring pinned", devs[di]);
        cudaEvent_t tev[2] = {nullptr, nullptr}; // timing of a batch's copies (on the copy stream)
        for (auto &ev : tev)
            if (rc == RP_OK && cudaEventCreate(&ev) != cudaSuccess) rc = fail(RP_ECUDA, "cudaEventCreate failed");
        // Drains one encoded batch: device image -> pinned pieces -> write tasks.  Runs on its own thread and the copy
        // stream while this worker already paints the next batch into the other image buffer.  Its time and byte counts
        // go to accumulators of its own (the worker updates `st` concurrently through paint_device / encode_device);
        // they are folded into `st` once the last copier has been joined.
        double drain_ms = 0.0;
        long long drain_bytes = 0;
        auto drain = [&, di](int b, const char *img, std::vector<long long> off) {
            if (!publish_sizes(b, off)) return;
            const std::string oerr = files_open.get();
            if (!oerr.empty()) {
                std::unique_lock<std::mutex> lk(mu);
                set_error(RP_EIO, oerr);
                return;
            }
            cudaSetDevice(devs[di]);
            cudaStream_t cs = c->copy_stream;
            // a few copies are kept queued on the stream (one event each); a piece goes to the writers as soon as its
            // event has fired, so the copy engine never waits for this thread
            std::vector<long long> cur(off.begin(), off.end() - 1); // next byte of every window's image to be copied
            std::deque<std::pair<int, WriteTask>> inflight;         // (event index, task), in stream order
            constexpr int kDepth = 4;
            int evn = 0;
            long long bytes = 0;
            cudaError_t e = cudaEventRecord(tev[0], cs);
            auto retire = [&](size_t keep) {
                while (inflight.size() > keep && e == cudaSuccess) {
                    e = cudaEventSynchronize(pev[inflight.front().first]);
                    pool.push(inflight.front().second);
                    inflight.pop_front();
                }
            };
            for (bool more = true; more && e == cudaSuccess;) {
                more = false;
                for (int w = 0; w < W && e == cudaSuccess; w++) {
                    if (cur[w] >= off[w + 1]) continue;
                    const long long lo = cur[w], n = std::min<long long>((long long)kOutPiece, off[w + 1] - lo);
                    cur[w] += n;
                    more = true;
                    const int sl = ring.acquire();
                    char *dst = ring.base + (size_t)sl * kOutPiece;
                    ring.pending[sl].store(1);
                    e = cudaMemcpyAsync(dst, img + lo, (size_t)n, cudaMemcpyDeviceToHost, cs);
                    if (e == cudaSuccess) e = cudaEventRecord(pev[evn], cs);
                    inflight.emplace_back(evn, WriteTask{dst, (size_t)n, fds[w], bbase[b][w] + (lo - off[w]), &ring, sl});
                    evn = (evn + 1) % (kDepth + 1);
                    bytes += n;
                    retire(kDepth);
                }
            }
            if (e == cudaSuccess) e = cudaEventRecord(tev[1], cs);
            retire(0);
            if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
            if (e != cudaSuccess) {
                cudaStreamSynchronize(cs);
                for (auto &it : inflight) ring.task_done(it.second.slot); // pieces that never reached the writers
                std::unique_lock<std::mutex> lk(mu);
                set_error(RP_ECUDA, std::string("copying the encoded records to the host: ") + cudaGetErrorString(e));
                return;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, tev[0], tev[1]);
            drain_ms += ms;
            drain_bytes += bytes;
            trace("batch copied to host", devs[di]);
        };
        std::thread copier;
        while (rc == RP_OK) {
            {
                std::lock_guard<std::mutex> lk(mu);
                if (first_rc != RP_OK) break;
            }
            const int b = next_batch.fetch_add(1);
            if (b >= nbatch) break;
            const int k0 = b * B, k1 = std::min(N, k0 + B);
            rc = paint_device(c, k0, k1, &st);
            trace("batch painted", devs[di]);
            if (rc == RP_OK) rc = encode_device(c, k1 - k0, &st);
            trace("batch encoded", devs[di]);
            if (rc != RP_OK) break;
            if (copier.joinable()) copier.join(); // the previous batch has left the other image buffer
            const char *img = c->image.as<char>();
            std::swap(c->image, c->image_alt);    // the next batch is encoded into the other buffer
            copier = std::thread(drain, b, img, c->h_img_off);
        }
        if (copier.joinable()) copier.join();
        st.ms_d2h += drain_ms;
        st.d2h_bytes += drain_bytes;
        if (rc != RP_OK) {
            std::unique_lock<std::mutex> lk(mu);
            set_error(rc, g_err);
        }
        ring.wait_all(); // every piece of this device has reached its file
        for (auto &ev : pev)
            if (ev) cudaEventDestroy(ev);
        for (auto &ev : tev)
            if (ev) cudaEventDestroy(ev);
        ws.shell = c; // park the workspace (may be nullptr if creation failed)
    };
    std::vector<std::thread> threads;
    for (int di = 1; di < (int)devs.size(); di++) threads.emplace_back(worker, di);
    worker(0);
    for (auto &t : threads) t.join();
    in.join();
    pool.finish();
    if (pool.failed() && first_rc == RP_OK) {
        first_rc = RP_EIO;
        first_err = "short write to the paint files";
    }
    const double ms_write = pool.span_ms();
    {
        const std::string err = files_open.get();
        for (int fd : fds)
            if (fd >= 0) close(fd);
        if (first_rc == RP_OK && !err.empty()) return fail(RP_EIO, err);
    }
    if (first_rc != RP_OK) return fail(first_rc, first_err);
    in.remove_sidecar();
    if (per_device) *per_device = dstats;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (const rp_stats &s : dstats) {
            stats->ms_h2d = std::max(stats->ms_h2d, s.ms_h2d);
            stats->ms_prep = std::max(stats->ms_prep, s.ms_prep);
            stats->ms_paint = std::max(stats->ms_paint, s.ms_paint);
            stats->ms_d2h = std::max(stats->ms_d2h, s.ms_d2h);
            stats->ms_rle = std::max(stats->ms_rle, s.ms_rle);
            stats->sites += s.sites;
            stats->cells += s.cells;
            stats->h2d_bytes += s.h2d_bytes;
            stats->d2h_bytes += s.d2h_bytes;
            stats->launches += s.launches;
            stats->n_targets += s.n_targets;
            stats->team_threads = s.team_threads;
            stats->words_per_thread = s.words_per_thread;
            stats->ctas = s.ctas;
        }
        stats->ms_write = ms_write;
        stats->ms_load = in.t_loaded - t0;
        stats->ms_total = now_ms() - t0;
    }
    return RP_OK;
}

} // namespace

// Data::Data(6 files) + the --painting handling of Paint.cpp:38-61, with the genotype rows bit-packed by reader threads on
// their way to the device (ChunkReaders: 1 bit per genotype crosses PCIe, the chars never reach HBM)
extern "C" int rp_chunk_load(int device, const char *out_dir, int chunk_index, const char *painting, unsigned flags,
                             rp_chunk **out)
{
    if (!out_dir || !out) return fail(RP_EINVAL, "null argument");
    *out = nullptr;
    rp::HostChunk hc;
    int hap_fd = -1;
    {
        std::string err = rp::load_chunk_small(out_dir, chunk_index, painting, hc, &hap_fd);
        if (!err.empty()) return fail(RP_EIO, err);
    }
    const int ndev = rp_device_count();
    if (ndev < 1 || device < 0 || device >= ndev) {
        close(hap_fd);
        return ndev < 1 ? fail(RP_ENODEVICE, "no CUDA device (there is no CPU fallback)") : fail(RP_EINVAL, "device index out of range");
    }
    std::lock_guard<std::mutex> stage_lock(g_stage_mu); // (the pinned input ring is the device's parked one)
    RP_CUDA(cudaSetDevice(device));
    ChunkReaders in;
    RP_TRY(in.start(out_dir, chunk_index, hc, hap_fd, 1, g_ws[device].hap_in));
    hc.hap = const_cast<char *>(in.feed.ring);
    const int rc = chunk_from_host(device, hc.N, hc.L, hc.hap, hc.r.data(), hc.wb.data(), (int)hc.wb.size(), hc.theta, flags, out,
                                   nullptr, &in.feed);
    if (rc != RP_OK) in.feed.abort.store(1);
    in.join(); // (the sidecar, if any, stays: only the stage, which paints the whole chunk, removes it)
    return rc;
}

extern "C" int rp_paint_chunk(const char *out_dir, int chunk_index, const char *painting, const int *devices,
                              int n_devices, unsigned flags, rp_stats *stats)
{
    if (!out_dir) return fail(RP_EINVAL, "null out_dir");
    std::vector<int> devs;
    RP_TRY(check_devices(devices, n_devices, devs));
    std::lock_guard<std::mutex> stage_lock(g_stage_mu);
    for (int d : devs) g_ws[d];
    return paint_chunk_stage(out_dir, chunk_index, painting, devs, flags, stats, &g_last_dstats);
}

extern "C" int rp_stage_device_stats(int i, rp_stats *out)
{
    if (!out) return fail(RP_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(g_stage_mu);
    if (i < 0 || i >= (int)g_last_dstats.size()) return fail(RP_EINVAL, "no such device in the last rp_paint_chunk call");
    *out = g_last_dstats[i];
    return RP_OK;
}

// Chunks first_chunk..last_chunk, whole chunks per device, largest first: one host thread per device
// pulls the next chunk from the sorted list and runs the single-device stage on it.  No collective; the only shared
// state is the list cursor.
extern "C" int rp_paint_chunks(const char *out_dir, int first_chunk, int last_chunk, const char *painting, const int *devices,
                               int n_devices, unsigned flags, rp_stats *stats)
{
    if (!out_dir) return fail(RP_EINVAL, "null out_dir");
    if (first_chunk < 0 || last_chunk < first_chunk) return fail(RP_EINVAL, "bad chunk range");
    const double t0 = now_ms();
    std::vector<int> devs;
    RP_TRY(check_devices(devices, n_devices, devs));
    std::vector<std::pair<long long, int>> order; // (-cost, chunk)
    for (int c = first_chunk; c <= last_chunk; c++) {
        const std::string p = std::string(out_dir) + "/parameters_c" + std::to_string(c) + ".bin";
        FILE *fp = fopen(p.c_str(), "rb");
        int hdr[2] = {0, 0};
        const bool ok = fp && fread(hdr, 4, 2, fp) == 2;
        if (fp) fclose(fp);
        if (!ok) return fail(RP_EIO, "cannot read " + p);
        order.emplace_back(-(long long)hdr[0] * hdr[0] * hdr[1], c); // painted cells N*N*L
    }
    std::sort(order.begin(), order.end());
    std::lock_guard<std::mutex> stage_lock(g_stage_mu);
    for (int d : devs) g_ws[d];
    std::atomic<int> cursor{0};
    std::mutex mu;
    int first_rc = RP_OK;
    std::string first_err;
    std::vector<rp_stats> acc(devs.size());
    auto worker = [&](int di) {
        rp_stats &a = acc[di];
        memset(&a, 0, sizeof a);
        for (;;) {
            const int i = cursor.fetch_add(1);
            if (i >= (int)order.size()) break;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (first_rc != RP_OK) break;
            }
            rp_stats st;
            memset(&st, 0, sizeof st);
            const int rc = paint_chunk_stage(out_dir, order[i].second, painting, std::vector<int>{devs[di]}, flags, &st);
            if (rc != RP_OK) {
                std::lock_guard<std::mutex> lk(mu);
                if (first_rc == RP_OK) {
                    first_rc = rc;
                    first_err = "chunk " + std::to_string(order[i].second) + ": " + g_err;
                }
                break;
            }
            a.ms_h2d += st.ms_h2d; a.ms_prep += st.ms_prep; a.ms_paint += st.ms_paint; a.ms_d2h += st.ms_d2h;
            a.ms_rle += st.ms_rle; a.ms_write += st.ms_write; a.ms_load += st.ms_load;
            a.sites += st.sites; a.cells += st.cells; a.h2d_bytes += st.h2d_bytes; a.d2h_bytes += st.d2h_bytes;
            a.launches += st.launches; a.n_targets += st.n_targets;
            a.team_threads = st.team_threads; a.words_per_thread = st.words_per_thread; a.ctas = st.ctas;
        }
    };
    std::vector<std::thread> threads;
    for (int di = 1; di < (int)devs.size(); di++) threads.emplace_back(worker, di);
    worker(0);
    for (auto &t : threads) t.join();
    if (first_rc != RP_OK) return fail(first_rc, first_err);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (const rp_stats &a : acc) { // per-device times: the busiest device; counts: summed
            stats->ms_h2d = std::max(stats->ms_h2d, a.ms_h2d); stats->ms_prep = std::max(stats->ms_prep, a.ms_prep);
            stats->ms_paint = std::max(stats->ms_paint, a.ms_paint); stats->ms_d2h = std::max(stats->ms_d2h, a.ms_d2h);
            stats->ms_rle = std::max(stats->ms_rle, a.ms_rle); stats->ms_write = std::max(stats->ms_write, a.ms_write);
            stats->ms_load = std::max(stats->ms_load, a.ms_load);
            stats->sites += a.sites; stats->cells += a.cells; stats->h2d_bytes += a.h2d_bytes; stats->d2h_bytes += a.d2h_bytes;
            stats->launches += a.launches; stats->n_targets += a.n_targets;
            if (a.n_targets) { stats->team_threads = a.team_threads; stats->words_per_thread = a.words_per_thread; stats->ctas = a.ctas; }
        }
        stats->ms_total = now_ms() - t0;
    }
    return RP_OK;
}
