// bt_gpu_shim.cpp — the reference-side binding of INTEGRATION.md section 4, made real for the tests.
//
// TEST INFRASTRUCTURE ONLY (see oracle/Makefile, target _ref/Relate_gpu).  The UNMODIFIED reference is compiled
// from /root/reference as for _ref/Relate; in a copy of its anc_builder.o the one symbol
// DistanceMeasure::GetMatrix(int) is weakened (objcopy --weaken-symbol), and this file supplies the strong
// definition: the body a maintainer would write to let BuildTopology take its distance matrices from the GPU
// (rp_window_open_files = RePaintSection for every target of the window, src/fast_painting.cpp:620-1092 as driven by
// DistanceMeasure::GetTopologyWithRepaint, src/anc_builder.cpp:48-106; rp_window_distance = GetMatrix,
// src/anc_builder.cpp:108-207).  Everything else of `Relate --mode BuildTopology` (MinMatch tree building, mutation
// mapping, the .anc/.mut writers) is the reference's own code, so "identical downstream topologies from GPU d_ij"
// is checked by the reference itself.
//
// No reference source is copied: the class is used through the reference's own header.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "anc_builder.hpp"
#include "data.hpp"

#include "../include/relate_paint.h"

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
    if (rp_chunk_create(0, data.N, data.L, &data.sequence[0][0], data.r.data(), g.wb.data(), nb, data.theta, 0,
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
