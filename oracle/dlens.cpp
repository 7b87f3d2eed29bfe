// dlens.cpp — "d_ij lens": drives the UNMODIFIED reference's
// DistanceMeasure::GetMatrix (src/anc_builder.cpp:108-207) over one window of a
// painted chunk and dumps the distance matrices, so that two paint directories
// (reference Paint vs this repo's CUDA Paint) can be compared in d_ij space.
//
// TEST INFRASTRUCTURE ONLY (see oracle/Makefile).  The walk over SNPs mirrors what
// AncesTreeBuilder::BuildTopology does around GetMatrix (src/anc_builder.cpp:423-434,
// 487-495,543): v_snp_prev / v_rpos_prev advance for every carrier of every SNP.
//
// usage: dlens <outdir> <chunk> <section> <stride> <painting|-> <out.bin>
//   writes: int N; int count; then per selected snp: int snp; float d[N*N]
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "anc_builder.hpp"
#include "data.hpp"

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: dlens <outdir> <chunk> <section> <stride> <painting|-> <out.bin>\n");
        return 2;
    }
    const std::string dir = std::string(argv[1]) + "/";
    const int chunk = atoi(argv[2]);
    const int section = atoi(argv[3]);
    const int stride = atoi(argv[4]);
    const std::string painting = argv[5];
    const std::string base = dir + "chunk_" + std::to_string(chunk);

    int N, L, nb;
    FILE *fp = fopen((dir + "parameters_c" + std::to_string(chunk) + ".bin").c_str(), "rb");
    if (!fp) { fprintf(stderr, "dlens: no parameters file\n"); return 1; }
    if (fread(&N, 4, 1, fp) != 1 || fread(&L, 4, 1, fp) != 1 || fread(&nb, 4, 1, fp) != 1) return 1;
    std::vector<int> wb(nb);
    if (fread(wb.data(), 4, nb, fp) != (size_t)nb) return 1;
    fclose(fp);
    const int W = nb - 1;
    if (section < 0 || section >= W) return 1;

    Data data((base + ".hap").c_str(), (base + ".bp").c_str(), (base + ".dist").c_str(),
              (base + ".r").c_str(), (base + ".rpos").c_str(), (base + ".state").c_str());
    data.name = base + "/paint/relate";
    if (painting != "-") { // pipeline/BuildTopology.cpp:48-70
        size_t c = painting.find(',');
        data.theta = std::stof(painting.substr(0, c));
        data.ntheta = 1.0 - data.theta;
        double rho = std::stof(painting.substr(c + 1));
        for (auto &x : data.r) x *= rho;
    }
    const int start = wb[section];
    const int end = (section < W - 1) ? wb[section + 1] - 1 : data.L - 1;

    DistanceMeasure d(data, section);
    FILE *out = fopen(argv[6], "wb");
    if (!out) return 1;
    int count = 0;
    fwrite(&N, 4, 1, out);
    fwrite(&count, 4, 1, out);
    for (int snp = start; snp <= end; snp++) {
        if (snp > start) {
            for (int n = 0; n < data.N; n++) {
                if (data.sequence[snp][n] == '1') {
                    d.v_snp_prev[n]++;
                    d.v_rpos_prev[n] = data.rpos[snp];
                }
            }
        }
        if (snp == start || (snp - start) % stride == 0 || snp == end) {
            d.GetMatrix(snp);
            fwrite(&snp, 4, 1, out);
            for (int n = 0; n < data.N; n++) fwrite(&d.matrix[n][0], 4, data.N, out);
            count++;
        }
    }
    fseek(out, 4, SEEK_SET);
    fwrite(&count, 4, 1, out);
    fclose(out);
    return 0;
}
