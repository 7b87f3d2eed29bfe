/*
 * relate_paint.h — C ABI of the B200-native replacement for Relate's chromosome-painting
 * hot path (`Relate --mode Paint`).
 *
 * The reference has no plugin/FFI interface: the path sits behind one in-process C++ call
 * and a set of files.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference checkout, include/...):
 *
 *   rp_paint_chunk      <- int Paint(cxxopts::Options&, int chunk_index)      pipeline/Paint.cpp:17-108
 *   rp_chunk_load       <- Data::Data(6 files) + --painting handling          src/data.cpp:86-97, pipeline/Paint.cpp:21-61
 *   rp_chunk_create     <- the in-memory Data the painter is handed            src/data.hpp (sequence, r, theta)
 *   rp_paint_targets    <- for(hap) FastPainting::PaintSteppingStones(...)     pipeline/Paint.cpp:81-87, src/fast_painting.cpp:18-618
 *   rp_make_chunks      <- int MakeChunks(cxxopts::Options&) + Data::MakeChunks pipeline/MakeChunks.cpp:14-117, src/data.cpp:117-518
 *   rp_paint_records    <- fast_painting.cpp:589-601 + DumpToFile, on the device
 *   rp_rle_encode       <- CollapsedMatrix<float>::DumpToFile (stepping stone) src/collapsed_matrix.hpp:228-265
 *   rp_fast_log_device  <- fast_log                                            src/fast_log.hpp:6-22
 *
 * Conventions: every function returns 0 on success and a negative RP_E* code on failure and
 * never calls exit()/abort(); rp_last_error() gives the message of the calling thread's last
 * failure.  All pointers are plain host pointers unless the name says `dev`.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with RP_ENODEVICE.
 */
#ifndef RELATE_PAINT_H
#define RELATE_PAINT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RP_OK 0
#define RP_EINVAL (-1)    /* bad argument / inconsistent input            */
#define RP_EIO (-2)       /* file missing, short read, cannot create       */
#define RP_ECUDA (-3)     /* CUDA runtime error (message has the details)  */
#define RP_ENODEVICE (-4) /* no usable CUDA device                         */
#define RP_ENOMEM (-5)
#define RP_EUNSUPPORTED (-6) /* e.g. N beyond a 16-CTA cluster (524288), fp64 mode beyond one CTA */

/* flags */
#define RP_FP64 1u /* verification mode: fp64 state and sums (default fp32 state, as north_star) */

typedef struct rp_chunk rp_chunk; /* opaque: a chunk resident in one GPU's HBM */

typedef struct rp_info {
    int N, L, W;           /* haplotypes, SNPs, windows                                */
    int device;
    int words_per_snp;     /* row stride of the SNP-major bit matrix, in 32-bit words  */
    long long hbm_bytes;   /* bytes of HBM held by the chunk (bit matrices, r, plan)   */
} rp_info;

typedef struct rp_stats {
    double ms_h2d;       /* host->device copies issued by the call                       */
    double ms_prep;      /* bit-pack / transpose / site tables (CUDA events)              */
    double ms_paint;     /* the forward/backward kernel (CUDA events, launching stream)   */
    double ms_d2h;       /* device->host copies                                           */
    double ms_encode;    /* host RLE + file writes (wall)                                 */
    double ms_total;     /* wall time of the call                                         */
    long long sites;     /* U = sum_k D_k over the painted targets (visited sites)        */
    long long cells;     /* painted cells: (#targets) * N * L                             */
    long long h2d_bytes, d2h_bytes;
    int launches;        /* kernel launches issued by the call                            */
    int n_targets;
    int team_threads;    /* threads cooperating on one (target,direction) job             */
    int words_per_thread;
    int ctas;            /* grid size of the paint kernel                                 */
    int reserved;
    double ms_load;      /* rp_paint_chunk: reading the chunk files (wall)                */
    double ms_rle;       /* device record encoder (CUDA events)                            */
    double ms_write;     /* rp_paint_chunk: writing the paint files (wall, overlaps painting) */
} rp_stats;

typedef struct rp_tune { /* all zero = automatic */
    int words_per_thread; /* 1 or 2: 32-bit genotype words (32 haplotypes each) per thread  */
    int ctas_per_sm;      /* persistent CTAs per SM                                          */
    int reserved[6];      /* [1]: chain segments of the paint kernel (0 = automatic, 1 = whole chains as jobs, n = every
                           * chain cut into n segments parked in HBM in between: load balance, results bit-identical);
                           * [3]: force teams of that many CTAs (thread-block cluster); others 0                    */
} rp_tune;

const char *rp_last_error(void);
int rp_device_count(void);
const char *rp_version(void);

/* Pinned host memory for callers that want full-speed PCIe copies (cudaMallocHost / cudaFreeHost). */
int rp_host_alloc(size_t bytes, void **out);
void rp_host_free(void *p);

/* ---- loader (SURVEY.md 8 "next" row f2) -------------------------------------------------------------
 * `Relate --mode MakeChunks`: SHAPEIT haps/sample (plain or gzip) + genetic map -> <out_dir>/parameters.bin,
 * props.bin, parameters_c<c>.bin and chunk_<c>.{hap,state,bp,dist,rpos,r}, byte-identical to the reference's
 * (same chunk / window boundary rules, data.cpp:129-229).  dist may be NULL ("unspecified"); transversion != 0
 * is the --transversion flag; memory_gb is --memory (reference default 5).  Like the reference it refuses an
 * existing out_dir and creates it with mode 0700.  Host-only.  n_chunks / warnings (a caller-provided buffer
 * for the lines the reference prints to stderr) may be NULL. */
int rp_make_chunks(const char *haps, const char *sample, const char *map, const char *dist, const char *out_dir,
                   int transversion, float memory_gb, int *n_chunks, char *warnings, size_t warnings_cap);
/* Same with options.  RP_MC_HAPBITS: also write <out_dir>/chunk_<c>.hapbits, the genotype rows of the chunk in the
 * painter's bit layout (header: "RPHBITS1", int N, L, words_per_snp, 0; then L rows of words_per_snp 32-bit words, bit
 * n&31 of word n>>5 = allele of haplotype n).  rp_paint_chunk reads it instead of the 8x larger chunk_<c>.hap when its
 * header matches, and deletes it after painting the chunk, so that the directory holds exactly the reference's files
 * again (its Finalize refuses to remove a directory with unknown files).  rp_make_chunks itself takes the option from
 * the environment (RELATE_HAPBITS=1); `relate --mode All`, where Paint is certain to follow, sets it. */
#define RP_MC_HAPBITS 1u
int rp_make_chunks_ex(const char *haps, const char *sample, const char *map, const char *dist, const char *out_dir,
                      int transversion, float memory_gb, unsigned mc_flags, int *n_chunks, char *warnings,
                      size_t warnings_cap);

/* ---- chunk lifetime ---------------------------------------------------------------- */
/* hap: L*N chars '0'/'1', SNP-major (Data::sequence); r: L doubles, already multiplied by rho;
 * wb: n_wb = W+1 window boundaries (wb[0]=0, wb[W]=L); theta as Data::theta.
 * The chars are copied to the device, bit-packed there (SNP-major and haplotype-major). */
int rp_chunk_create(int device, int N, int L, const char *hap, const double *r, const int *wb, int n_wb,
                    double theta, unsigned flags, rp_chunk **out);
/* Reads <out_dir>/parameters_c<c>.bin and chunk_<c>.{hap,r} (and checks .bp,.dist,.rpos,.state exist, as
 * the reference loader opens them); painting = the --painting string "theta,rho" or NULL. */
int rp_chunk_load(int device, const char *out_dir, int chunk_index, const char *painting, unsigned flags,
                  rp_chunk **out);
int rp_chunk_info(const rp_chunk *c, rp_info *info);
void rp_chunk_free(rp_chunk *c);
int rp_chunk_set_tune(rp_chunk *c, const rp_tune *t);
/* Issue this chunk's copies and kernels on the caller's CUDA stream (a cudaStream_t passed as void*; NULL restores
 * the chunk's own stream).  Lets a caller bracket the work with events of its own. */
int rp_chunk_set_stream(rp_chunk *c, void *cuda_stream);

/* ---- painting ---------------------------------------------------------------------- */
/* Paint targets k in [k_begin, k_end).  Host outputs (pre-RLE), T = k_end-k_begin:
 *   alpha, beta        float [T][W][N]   forward / backward stepping-stone vectors
 *   ls_alpha, ls_beta  float [T][W]      their log-scales
 *   site_begin/end     int   [T][W]      boundary SNPs (boundarySNP_begin / boundarySNP_end)
 * Any output pointer may be NULL to skip its copy. */
int rp_paint_targets(rp_chunk *c, int k_begin, int k_end, float *alpha, float *beta, float *ls_alpha,
                     float *ls_beta, int *site_begin, int *site_end, rp_stats *stats);
/* Same computation, results left in HBM (buffers owned by the chunk, valid until the next paint call on
 * it or rp_chunk_free).  Used by bench.py's device-resident leg and by consumers that stay on the GPU. */
int rp_paint_targets_device(rp_chunk *c, int k_begin, int k_end, const float **dev_alpha,
                            const float **dev_beta, const float **dev_ls_alpha, const float **dev_ls_beta,
                            const int **dev_site_begin, const int **dev_site_end, rp_stats *stats);

/* Paint targets [k_begin, k_end) and encode them on the device into the records the reference writes
 * (fast_painting.cpp:589-601 + CollapsedMatrix<float>::DumpToFile, collapsed_matrix.hpp:228-265): for each window
 * w the bytes that targets k_begin..k_end-1 contribute to chunk_<c>/paint/relate_<w>.bin, contiguous and in target
 * order.  The W images are laid out back to back in HBM; win_off (W+1 entries, host) receives their byte offsets,
 * win_off[W] = total size.  rp_records_copy then copies `bytes` bytes starting at byte `offset` to the host. */
int rp_paint_records(rp_chunk *c, int k_begin, int k_end, long long *win_off, rp_stats *stats);
int rp_records_copy(rp_chunk *c, long long offset, long long bytes, void *host_image, rp_stats *stats);

/* One call from a host-resident Data to host-resident stepping stones: chunk_create + paint_targets + free.
 * This is the call bench.py times for its end-to-end ("e2e") number: hap/r go host->device and the
 * stepping stones come device->host inside it.  tune may be NULL. */
int rp_paint_from_host(int device, int N, int L, const char *hap, const double *r, const int *wb, int n_wb,
                       double theta, unsigned flags, const rp_tune *tune, int k_begin, int k_end, float *alpha,
                       float *beta, float *ls_alpha, float *ls_beta, int *site_begin, int *site_end,
                       rp_stats *stats);

/* The whole stage: load chunk files (the genotype rows are bit-packed by the reader threads, so the devices
 * receive 1 bit per genotype), paint every target on the given devices, encode the records on the device and write
 * <out_dir>/chunk_<c>/paint/relate_<w>.bin in target order.  devices==NULL: device 0 only; a device index may
 * appear only once (RP_EINVAL otherwise).  Every device holds a replica of the chunk; targets are cut into
 * equal-count batches which the devices pull from one shared counter (dynamic balance, no collective); a batch's
 * records are written at absolute file offsets as soon as the sizes of all earlier batches are known.
 * Parity: with flags == 0 the state is fp32; stepping stones agree with the reference's to ~1e-6 relative (gate 1e-4),
 * so the paint files are NOT byte-identical to the reference's (the lossy codec may also start a run one element
 * earlier or later), but the reference's BuildTopology gives identical trees on the tested data.  With RP_FP64 the
 * files are byte-identical to the reference's on every fixture the tests hold. */
int rp_paint_chunk(const char *out_dir, int chunk_index, const char *painting, const int *devices,
                   int n_devices, unsigned flags, rp_stats *stats);

/* Statistics of device i (position in `devices`) of this process's last rp_paint_chunk call: its own kernel, prep,
 * encoder and copy times and the targets / visited sites it painted (rp_stats of the call itself holds the maximum
 * over devices of the times and the sums of the counts).  Host threads used for file I/O: RP_IO_THREADS in the
 * environment, else the core count divided by LOCAL_WORLD_SIZE (one process per GPU launchers set it) when present, else all cores. */
int rp_stage_device_stats(int i, rp_stats *out);

/* Several chunks of one data set: chunks first_chunk..last_chunk are distributed over the
 * devices as whole chunks, largest first, one host thread per device, no collective.  What a maintainer would call
 * instead of the per-chunk loop of RelateParallel.sh:221-225 when several GPUs are present. */
int rp_paint_chunks(const char *out_dir, int first_chunk, int last_chunk, const char *painting, const int *devices,
                    int n_devices, unsigned flags, rp_stats *stats);

/* ---- window repaint + distance matrices (consumer side, SURVEY.md 8 "next" row f1) ------------------
 *   rp_window_open      <- FastPainting::RePaintSection for every target     src/fast_painting.cpp:620-1092
 *                          (as DistanceMeasure::GetTopologyWithRepaint drives it, src/anc_builder.cpp:48-106)
 *   rp_window_distance  <- DistanceMeasure::GetMatrix(snp)                   src/anc_builder.cpp:108-207
 * The posterior rows of the whole window stay in HBM; each distance call returns the N x N float matrix d for
 * one SNP of the window (row n = "n painted against everyone", diagonal 0, row minimum subtracted). */
typedef struct rp_window rp_window;
/* alpha, beta: float [N][N] = the DECODED stepping stones of window w (what ReadFromFile yields) for targets
 * 0..N-1; ls_alpha, ls_beta: float [N]; rpos: double [L+1] (chunk_<c>.rpos). */
int rp_window_open(rp_chunk *c, int w, const float *alpha, const float *beta, const float *ls_alpha,
                   const float *ls_beta, const double *rpos, rp_window **out, rp_stats *stats);
/* Same, from the stepping stones still resident in HBM: the last paint call on the chunk must have covered targets
 * 0..N-1 (rp_paint_targets_device(c, 0, N, ...)).  The codec's lossy collapse (what a reader of the paint files
 * sees) is applied on the device, so the result is bit-identical to writing the files and opening them. */
int rp_window_open_resident(rp_chunk *c, int w, const double *rpos, rp_window **out, rp_stats *stats);
/* Same, reading <out_dir>/chunk_<c>/paint/relate_<w>.bin and chunk_<c>.rpos. */
int rp_window_open_files(rp_chunk *c, const char *out_dir, int chunk_index, int w, rp_window **out, rp_stats *stats);
/* Lifetime: a window refers to its chunk (bit matrices, stream); close it before freeing the chunk.  A window whose
 * chunk has been freed only accepts rp_window_close; rp_window_distance on it fails with RP_EINVAL. */
int rp_window_distance(rp_window *win, int snp, float *d);
long long rp_window_rows(const rp_window *win);
void rp_window_close(rp_window *win);

/* ---- tree builder (consumer side, SURVEY.md 8 "next" row f4) -----------------------------------------
 *   rp_minmatch_create      <- MinMatch::MinMatch(Data&)                       src/tree_builder.cpp:38-56
 *   rp_minmatch_quickbuild  <- MinMatch::QuickBuild(d, tree, sample_ages)      src/tree_builder.cpp:1060-1303
 *                              MinMatch::QuickBuild(d, tree, sample_ages, d_prior)            :2357-2646
 *                              (with Initialize :58-146 / :1646-1735, Coalesce :295-598 / :1843-2070, InitializeSym
 *                              :254-293, CoalesceSym :967-1058) in the form BuildTopology uses them without
 *                              --sample_ages: sample_ages empty, no template tree.
 * One handle = one reference MinMatch object: what survives from tree to tree inside it (min_values_CF, the lineage
 * names of reset candidates) survives in the handle, so a sequence of calls yields the sequence of trees the
 * reference builds — AncesTreeBuilder::BuildTopology creates one object per window (src/anc_builder.cpp:422).
 * The tree comes back as its merge list: merges[2t], merges[2t+1] = labels of child_left, child_right of node N+t
 * (t = 0..N-2), i.e. exactly what QuickBuild stores into tree.nodes (:1268-1274).  The lists are IDENTICAL to the
 * reference's for identical input matrices (same float arithmetic, same order of std::mt19937 draws).
 * d (and d_prior) are N x N row-major floats; neither is modified (the reference modifies d in place and its caller
 * does not read it again).  A handle is not re-entrant; distinct handles may be used from distinct threads, and each tree
 * occupies one SM, so trees of different windows (different handles) run side by side on one device: measured at N=1000,
 * 64 handles on 64 host threads, 2 900 trees/s against 109 for one handle (CUDA_DEVICE_MAX_CONNECTIONS=32 in the environment
 * before the first CUDA call: with the default 8 hardware queues kernels of streams that share a queue serialise, 850/s). */
typedef struct rp_minmatch rp_minmatch;
typedef struct rp_minmatch_stats {
    float ms_kernel;          /* device time of the tree kernel (CUDA events on the handle's stream) */
    long long draws;          /* random numbers drawn = feasible pairs met                          */
    int first_fallback_step;  /* first merge without a mutually minimal pair (-1: none)              */
    int fallback_steps;       /* merges taken while the symmetric fallback matrix was in use         */
    int general_steps;        /* merges whose pairs went through the any-size path (> 63 rows to rescan) */
    int medium_steps;         /* merges with more than 128 feasible pairs (ties): ballot-ranked path   */
    int launches;
} rp_minmatch_stats;
int rp_minmatch_create(int device, int N, double theta, rp_minmatch **out);
/* same from the two thresholds a MinMatch object holds (threshold, threshold_CF; tree_builder.cpp:43-44) */
int rp_minmatch_create_thresholds(int device, int N, float threshold, float threshold_cf, rp_minmatch **out);
void rp_minmatch_destroy(rp_minmatch *mm);
/* back to the state of a freshly constructed MinMatch object (what a new `MinMatch tb(data)` is in the reference): lets a
 * consumer that meets one object after the other (one per window) keep one handle and its N x N buffers */
int rp_minmatch_reset(rp_minmatch *mm);
/* d, d_prior on the host (d_prior NULL: the three-argument QuickBuild); merges: int [2*(N-1)] on the host */
int rp_minmatch_quickbuild(rp_minmatch *mm, const float *d, const float *d_prior, int *merges, rp_minmatch_stats *stats);
/* same with the matrices already in this device's memory (e.g. left there by the distance kernel) */
int rp_minmatch_quickbuild_device(rp_minmatch *mm, const float *dev_d, const float *dev_prior, int *merges,
                                  rp_minmatch_stats *stats);

/* rp_paint_chunk parks device buffers, pinned staging and streams per device between calls, and device memory released
 * by freed chunks and closed windows is kept in a per-device pool for the next allocation it fits (allocating a window's
 * posterior, GBs, costs more than repainting it); this returns all of it to the driver. */
void rp_release_cache(void);

/* ---- small pieces exposed for the parity tests -------------------------------------- */
/* Host restatement of the record codec's run rule (rp_paint_chunk / rp_paint_records use the device encoder,
 * rle_kernel; the tests hold the two byte-identical); returns the number of runs K (vals/lens sized n). */
int rp_rle_encode(const float *v, int n, float *vals, int *lens);
/* Evaluates the kernel's device fast_log on n floats (host in/out). */
int rp_fast_log_device(int device, const float *in, float *out, int n);
/* Measures the FP32 lane-operation rate of the device with the paint step's instruction mix (packed add, scalar
 * multiply, packed add; no memory): the measured denominator of the roofline bench.py reports.  `mix` uses
 * add.f32x2 as the kernel does, `scalar` the same arithmetic with scalar adds. */
int rp_peak_fp32(int device, double *lane_ops_per_s_mix, double *lane_ops_per_s_scalar);
/* Bit-packs on the device and returns the SNP-major and haplotype-major bit matrices (host out). */
int rp_debug_pack(int device, int N, int L, const char *hap, uint32_t *snp_major, int *words_per_snp,
                  uint32_t *hap_major, int *words_per_hap);

/* The stage driver's host-side packer (what its reader threads apply to the rows of chunk_<c>.hap they read): L rows of
 * N chars -> L rows of words_per_snp words, bit n&31 of word n>>5 = (hap[n] == '1'), padding zero.  Host-only. */
int rp_debug_pack_host(int N, int L, const char *hap, uint32_t *snp_major, int words_per_snp);

#ifdef __cplusplus
}
#endif
#endif /* RELATE_PAINT_H */
