// qblens.cpp — "tree lens": drives the UNMODIFIED reference's MinMatch::QuickBuild (src/tree_builder.cpp:1060-1303 and
// :2357-2646) over a file of distance matrices and dumps the merge lists, so that this repo's tree builder
// (rp_minmatch_*) and its oracle restatement (oracle/minmatch_oracle.c) can be compared with the reference tree by tree.
//
// TEST INFRASTRUCTURE ONLY (see oracle/Makefile).  One MinMatch object serves all matrices of a file, as the one
// AncesTreeBuilder::BuildTopology creates per window does (src/anc_builder.cpp:422): state that survives from tree to
// tree inside the object is part of what is compared.
//
// usage: qblens <in.bin> <out.bin> [repeat]
//   in : int N; double theta; int count; then per tree: int has_prior; float d[N*N]; float prior[N*N] if has_prior
//   out: per tree 2*(N-1) ints: labels of (child_left, child_right) of nodes N .. 2N-2
//   stderr: seconds spent inside QuickBuild (summed; `repeat` runs every tree that many times on copies for timing)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "anc.hpp"
#include "data.hpp"
#include "tree_builder.hpp"

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: qblens <in.bin> <out.bin> [repeat]\n");
        return 2;
    }
    const int repeat = argc > 3 ? atoi(argv[3]) : 1;
    FILE *in = fopen(argv[1], "rb");
    FILE *out = fopen(argv[2], "wb");
    if (!in || !out) return 1;
    int N, count;
    double theta;
    if (fread(&N, 4, 1, in) != 1 || fread(&theta, 8, 1, in) != 1 || fread(&count, 4, 1, in) != 1) return 1;
    Data data(N, 1);
    data.theta = theta;
    data.ntheta = 1.0 - theta;
    double seconds = 0;
    std::vector<double> no_ages;
    std::vector<float> buf((size_t)N * N);
    for (int rep = 0; rep < repeat; rep++) {
        if (rep) fseek(in, 16, SEEK_SET);
        MinMatch tb(data);
        for (int t = 0; t < count; t++) {
            int has_prior;
            if (fread(&has_prior, 4, 1, in) != 1) return 1;
            CollapsedMatrix<float> d, prior;
            d.resize(N, N);
            if (fread(buf.data(), 4, buf.size(), in) != buf.size()) return 1;
            for (int r = 0; r < N; r++)
                for (int c = 0; c < N; c++) d[r][c] = buf[(size_t)r * N + c];
            if (has_prior) {
                prior.resize(N, N);
                if (fread(buf.data(), 4, buf.size(), in) != buf.size()) return 1;
                for (int r = 0; r < N; r++)
                    for (int c = 0; c < N; c++) prior[r][c] = buf[(size_t)r * N + c];
            }
            Tree tree;
            auto t0 = std::chrono::steady_clock::now();
            if (has_prior) tb.QuickBuild(d, tree, no_ages, prior);
            else tb.QuickBuild(d, tree, no_ages);
            seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (rep == 0)
                for (int n = N; n < 2 * N - 1; n++) {
                    int pair[2] = {tree.nodes[n].child_left->label, tree.nodes[n].child_right->label};
                    fwrite(pair, 4, 2, out);
                }
        }
    }
    fclose(out);
    fprintf(stderr, "qblens: %d trees x %d, N=%d, QuickBuild %.6f s\n", count, repeat, N, seconds);
    return 0;
}
