// minmatch_gpu.cpp — the reference-side binding that lets `Relate --mode BuildTopology` build its trees on the GPU
// (SURVEY.md section 8 row f4).  PRODUCT artefact, linked into `Relate_gpu` next to
// distance_measure_gpu.cpp by relate_b200/integration/Makefile.
//
// These are the bodies a maintainer would give the two MinMatch::QuickBuild overloads BuildTopology calls
// (src/tree_builder.cpp:1060-1303 and :2357-2646; call sites src/anc_builder.cpp:447, 608, 612): hand the distance
// matrix (and the prior derived from the previous tree) to rp_minmatch_quickbuild, get the merge list back, and store
// it into tree.nodes exactly as QuickBuild does (:1268-1274).  The reference's sources are not edited and none of them
// is copied: the class is used through the reference's own header; at link time the two symbols are RENAMED in a copy of
// the reference's tree_builder.o (objcopy --redefine-sym), so that these definitions take their place and the
// originals stay reachable for what the GPU builder does not cover (--sample_ages, a template tree: the reference's own
// code runs, as it does for every stage this repo leaves to the reference) and for the built-in cross-check.
//
//   RELATE_GPU_MINMATCH=0          trees by the reference's CPU code (A/B timing)
//   RELATE_GPU_MINMATCH_MIN_N=<n>  smallest N built on the GPU (default 600: one CTA per tree pays ~10 us per merge whatever N
//                                  is, the CPU 0.6 ms per tree at N=200 and 16 ms at N=1000, 0.8 s at N=5000; measured on B200)
//   RELATE_GPU_MINMATCH_VERIFY=1   build every tree both ways and abort on the first differing merge
//   RELATE_GPU_MINMATCH_STATS=1    one line on stderr at exit: trees, seconds in QuickBuild, kernel seconds
//   RELATE_GPU_MINMATCH_POOL=<n>   GPU handles kept (default 2): a new MinMatch object takes the least recently used one, reset;
//                                  only objects that are in use at the same time need one each
//   RELATE_GPU_DEVICE=<i>          CUDA device (default 0)
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tree_builder.hpp"

#include "relate_paint.h" // this repo's include/

extern "C" { // the reference's own definitions under the names the Makefile gives them (member functions: `this` first)
void rp_ref_minmatch_quickbuild(MinMatch *self, CollapsedMatrix<float> &d, Tree &tree, std::vector<double> &sample_ages, Tree *tmpl_tree);
void rp_ref_minmatch_quickbuild_prior(MinMatch *self, CollapsedMatrix<float> &d, Tree &tree, std::vector<double> &sample_ages,
                                      const CollapsedMatrix<float> &d_prior);
}

namespace {

struct Builders {
    // BuildTopology makes one MinMatch object per window, one after the other, and never tells when it is done with one: the
    // handles (five N x N matrices each) are pooled and the least recently used one is reset for the next new object
    struct Slot { rp_minmatch *h; int owner, N; long last_use; float thr, thr_cf; };
    std::vector<Slot> pool;
    int next_owner = 0;
    long clock = 0;
    long trees = 0, trees_ref = 0, draws = 0, general_steps = 0, medium_steps = 0, fallback_steps = 0;
    double seconds = 0, kernel_seconds = 0;
    ~Builders()
    {
        for (Slot &sl : pool) rp_minmatch_destroy(sl.h);
        if (getenv("RELATE_GPU_MINMATCH_STATS"))
            fprintf(stderr, "Relate_gpu: QuickBuild: %ld trees on the GPU (%ld by the reference's code), %.3f s in the call, %.3f s in the kernel; "
                            "%ld draws, %ld merge steps with more than 128 pairs, %ld on the any-size pair path, %ld on the symmetric fallback\n",
                    trees, trees_ref, seconds, kernel_seconds, draws, medium_steps, general_steps, fallback_steps);
    }
};
Builders g;

const bool use_gpu = !(getenv("RELATE_GPU_MINMATCH") && atoi(getenv("RELATE_GPU_MINMATCH")) == 0);
const bool verify = getenv("RELATE_GPU_MINMATCH_VERIFY") != nullptr;
const int min_n = getenv("RELATE_GPU_MINMATCH_MIN_N") ? atoi(getenv("RELATE_GPU_MINMATCH_MIN_N")) : 600;

void store_tree(Tree &tree, const int *merges, int N)
{
    const int N_total = 2 * N - 1;
    tree.nodes.resize(N_total);
    tree.nodes[N_total - 1].label = N_total - 1;
    for (int k = 0; k < N; k++) tree.nodes[k].label = k;
    for (int t = 0; t < N - 1; t++) {
        Node &parent = tree.nodes[N + t], &left = tree.nodes[merges[2 * t]], &right = tree.nodes[merges[2 * t + 1]];
        left.parent = &parent;
        right.parent = &parent;
        left.num_events = 0.0;
        right.num_events = 0.0;
        parent.child_left = &left;
        parent.child_right = &right;
        parent.label = N + t;
    }
}

} // namespace

// One GPU handle per MinMatch object in use.  A freshly constructed object has candidates_to_check_size == 0 (a member the
// reference never touches after its in-class initialiser, tree_builder.hpp:56); the binding keeps the object's number
// there, so that a new object — BuildTopology makes one per window — starts from fresh state as the reference's does.
static void gpu_quickbuild(MinMatch *self, int &slot, int N, float threshold, float threshold_CF, CollapsedMatrix<float> &d,
                           const CollapsedMatrix<float> *prior, Tree &tree)
{
    const size_t pool_max = getenv("RELATE_GPU_MINMATCH_POOL") ? (size_t)std::max(1, atoi(getenv("RELATE_GPU_MINMATCH_POOL"))) : 2;
    Builders::Slot *mine = nullptr;
    if (slot != 0)
        for (Builders::Slot &sl : g.pool)
            if (sl.owner == slot) mine = &sl;
    if (slot != 0 && !mine) {
        fprintf(stderr, "Relate_gpu: more than %zu MinMatch objects in use at once; raise RELATE_GPU_MINMATCH_POOL\n", pool_max);
        exit(1);
    }
    if (slot == 0) { // a new object: a new handle while the pool has room, else the least recently used one, reset
        Builders::Slot *lru = nullptr;
        for (Builders::Slot &sl : g.pool)
            if (sl.N == N && sl.thr == threshold && sl.thr_cf == threshold_CF && (!lru || sl.last_use < lru->last_use)) lru = &sl;
        if (g.pool.size() < pool_max || !lru) {
            rp_minmatch *h = nullptr;
            const int device = getenv("RELATE_GPU_DEVICE") ? atoi(getenv("RELATE_GPU_DEVICE")) : 0;
            if (rp_minmatch_create_thresholds(device, N, threshold, threshold_CF, &h) != RP_OK) {
                fprintf(stderr, "Relate_gpu: rp_minmatch_create: %s\n", rp_last_error());
                exit(1);
            }
            g.pool.push_back({h, 0, N, 0, threshold, threshold_CF});
            mine = &g.pool.back();
        } else {
            if (rp_minmatch_reset(lru->h) != RP_OK) {
                fprintf(stderr, "Relate_gpu: rp_minmatch_reset: %s\n", rp_last_error());
                exit(1);
            }
            mine = lru;
        }
        mine->owner = slot = ++g.next_owner;
    }
    mine->last_use = ++g.clock;
    rp_minmatch *handle = mine->h;
    std::vector<int> merges(2 * (size_t)(N - 1));
    rp_minmatch_stats st;
    const CollapsedMatrix<float> *use_prior = (prior && prior->size() == d.size()) ? prior : nullptr;
    // d_CF = d_prior, or d itself when the prior has another size (:2365-2368)
    const float *prior_ptr = prior ? (use_prior ? &(*use_prior)[0][0] : &d[0][0]) : nullptr;
    if (rp_minmatch_quickbuild(handle, &d[0][0], prior_ptr, merges.data(), &st) != RP_OK) {
        fprintf(stderr, "Relate_gpu: rp_minmatch_quickbuild: %s\n", rp_last_error());
        exit(1);
    }
    g.kernel_seconds += 1e-3 * st.ms_kernel;
    g.draws += st.draws;
    g.general_steps += st.general_steps;
    g.medium_steps += st.medium_steps;
    g.fallback_steps += st.fallback_steps;
    store_tree(tree, merges.data(), N);
    if (verify) {
        Tree ref;
        std::vector<double> none;
        CollapsedMatrix<float> d_copy = d;
        if (prior) rp_ref_minmatch_quickbuild_prior(self, d_copy, ref, none, *prior);
        else rp_ref_minmatch_quickbuild(self, d_copy, ref, none, nullptr);
        for (int n = N; n < 2 * N - 1; n++)
            if (ref.nodes[n].child_left->label != merges[2 * (n - N)] || ref.nodes[n].child_right->label != merges[2 * (n - N) + 1]) {
                fprintf(stderr, "Relate_gpu: tree %ld differs from the reference's at node %d: (%d, %d) vs (%d, %d)\n", g.trees, n,
                        merges[2 * (n - N)], merges[2 * (n - N) + 1], ref.nodes[n].child_left->label, ref.nodes[n].child_right->label);
                exit(1);
            }
    }
    g.trees++;
}

void MinMatch::QuickBuild(CollapsedMatrix<float> &d, Tree &tree, std::vector<double> &sample_ages, Tree *tmpl_tree)
{
    const auto t0 = std::chrono::steady_clock::now();
    if (!use_gpu || N < min_n || tmpl_tree != NULL || (int)sample_ages.size() == N) {
        rp_ref_minmatch_quickbuild(this, d, tree, sample_ages, tmpl_tree);
        g.trees_ref++;
    } else
        gpu_quickbuild(this, candidates_to_check_size, N, threshold, threshold_CF, d, nullptr, tree);
    g.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void MinMatch::QuickBuild(CollapsedMatrix<float> &d, Tree &tree, std::vector<double> &sample_ages, const CollapsedMatrix<float> &d_prior)
{
    const auto t0 = std::chrono::steady_clock::now();
    if (!use_gpu || N < min_n || (int)sample_ages.size() == N) {
        rp_ref_minmatch_quickbuild_prior(this, d, tree, sample_ages, d_prior);
        g.trees_ref++;
    } else
        gpu_quickbuild(this, candidates_to_check_size, N, threshold, threshold_CF, d, &d_prior, tree);
    g.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
