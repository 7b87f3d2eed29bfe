// distance_measure_gpu.cpp — the reference-side binding that lets `Relate --mode BuildTopology` take its distance
// matrices from the GPU.  PRODUCT artefact: relate_b200/integration/Makefile links it with a
// reference checkout into `Relate_gpu`; this repo's CLI runs that binary for `--mode BuildTopology` / `--mode All`
// when it is present, and with `--resident` Paint -> d_ij never touches the disk.
//
// It is the body a maintainer would give DistanceMeasure::GetMatrix(int) (src/anc_builder.cpp:108-207, which calls
// GetTopologyWithRepaint, :48-106): rp_window_open_files / rp_window_open_resident = RePaintSection for every target of
// the window (src/fast_painting.cpp:620-1092), rp_window_distance = GetMatrix.  The reference's sources are not edited
// and none of them is copied: the class is used through the reference's own header, and at link time the one symbol
// DistanceMeasure::GetMatrix(int) is weakened in a copy of the reference's anc_builder.o (objcopy --weaken-symbol) so
// that this strong definition wins.  Everything else of BuildTopology (MinMatch tree building, mutation mapping, the
// .anc/.mut writers) stays the reference's own code.
//
//   RELATE_GPU_RESIDENT=1   no paint files: paint every target once inside this process, keep the stepping stones in
//                           HBM and open each window from there (bit-identical matrices to the file round trip)
//   RELATE_GPU_DEVICE=<i>   CUDA device (default 0)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "anc_builder.hpp"
#include "data.hpp"

#include "relate_paint.h" // this repo's include/

namespace {

struct GpuState {
    Data *data = nullptr;
    rp_chunk *chunk = nullptr;
    rp_window *win = nullptr;
    bool painted = false; // RELATE_GPU_RESIDENT: the stepping stones of all targets are in HBM
    std::string out_dir;
    int chunk_index = 0;
    std::vector<int> wb;
    ~GpuState()
    {
        if (win) rp_window_close(win);
        if (chunk) rp_chunk_free(chunk);
    }
};
GpuState g;

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "Relate_gpu: %s: %s\n", what, rp_last_error());
    exit(1); // the reference's own error behaviour on this path is assert / exit(1)
}

// data.name is "<out>/chunk_<c>/paint/relate" (pipeline/BuildTopology.cpp:36)
void open_chunk(Data &data)
{
    std::string name = data.name;
    const std::string tail = "/paint/relate";
    if (name.size() < tail.size() || name.compare(name.size() - tail.size(), tail.size(), tail) != 0) {
        fprintf(stderr, "Relate_gpu: unexpected data.name %s\n", name.c_str());
        exit(1);
    }
    name.erase(name.size() - tail.size());
    const size_t us = name.rfind("chunk_");
    g.out_dir = name.substr(0, us);
    if (g.out_dir.empty()) g.out_dir = "./";
    g.chunk_index = atoi(name.c_str() + us + 6);
    FILE *fp = fopen((g.out_dir + "parameters_c" + std::to_string(g.chunk_index) + ".bin").c_str(), "rb");
    int N, L, nb;
    if (!fp || fread(&N, 4, 1, fp) != 1 || fread(&L, 4, 1, fp) != 1 || fread(&nb, 4, 1, fp) != 1) {
        fprintf(stderr, "Relate_gpu: cannot read the chunk's parameters file\n");
        exit(1);
    }
    g.wb.resize(nb);
    if (fread(g.wb.data(), 4, nb, fp) != (size_t)nb) exit(1);
    fclose(fp);
    // the chunk as BuildTopology holds it: theta from --painting, r already multiplied by rho
    const int device = getenv("RELATE_GPU_DEVICE") ? atoi(getenv("RELATE_GPU_DEVICE")) : 0;
    if (rp_chunk_create(device, data.N, data.L, &data.sequence[0][0], data.r.data(), g.wb.data(), nb, data.theta, 0,
                        &g.chunk) != RP_OK)
        die("rp_chunk_create");
    g.data = &data;
}

} // namespace

// Replaces DistanceMeasure::GetMatrix (src/anc_builder.cpp:108-207) and, through it, GetTopologyWithRepaint (:48-106).
void DistanceMeasure::GetMatrix(const int snp)
{
    if (g.data != data) { // a new DistanceMeasure on another Data object: (re)load the chunk
        if (g.win) rp_window_close(g.win), g.win = nullptr;
        if (g.chunk) rp_chunk_free(g.chunk), g.chunk = nullptr;
        g.painted = false;
        open_chunk(*data);
    }
    if (snp > section_endpos || g.win == nullptr) {
        if (g.win) rp_window_close(g.win), g.win = nullptr;
        if (getenv("RELATE_GPU_RESIDENT")) {
            // No paint files at all: paint every target once, keep the stepping stones in HBM (8*N*N*W bytes) and open
            // each window from there (the codec's lossy collapse is applied on the device, so the matrices are
            // bit-identical to the file round trip).  `--mode Paint` becomes unnecessary for this consumer.
            if (!g.painted) {
                if (rp_paint_targets_device(g.chunk, 0, data->N, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr) != RP_OK)
                    die("rp_paint_targets_device");
                g.painted = true;
            }
            if (rp_window_open_resident(g.chunk, section, data->rpos.data(), &g.win, nullptr) != RP_OK)
                die("rp_window_open_resident");
        } else if (rp_window_open_files(g.chunk, g.out_dir.c_str(), g.chunk_index, section, &g.win, nullptr) != RP_OK)
            die("rp_window_open_files");
        // what the reference reads from the head of every record (fast_painting.cpp:589-590)
        section_startpos = g.wb[section];
        section_endpos = g.wb[section + 1] - 1;
        section++;
    }
    if (rp_window_distance(g.win, snp, &matrix[0][0]) != RP_OK) die("rp_window_distance");
}
