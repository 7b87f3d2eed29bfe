// host_io.hpp — host-side file formats at the Paint boundary.
//
//  * chunk loader: what Data::Data(6 files) + Paint() read
//    (/root/reference/include/src/data.cpp:86-97,531-540; data.hpp:91-101;
//     collapsed_matrix.hpp:215-225; pipeline/Paint.cpp:21-61)
//  * stepping-stone record encoder: CollapsedMatrix<float>::DumpToFile
//    (/root/reference/include/src/collapsed_matrix.hpp:228-265)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

namespace rp {

struct HostChunk {
    int N = 0, L = 0;
    std::vector<int> wb;      // W+1 window boundaries
    char *hap = nullptr;      // L*N chars, SNP-major; points into hap_own or into caller-provided (pinned) memory
    std::vector<char> hap_own;
    std::vector<double> r;    // L, multiplied by rho
    double theta = 0.001;     // data.cpp:95
};

inline bool file_exists(const std::string &p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

// Everything of a chunk except the genotype bytes: parameters, .r, presence of the four files the reference loader
// also opens, --painting handling.  Opens chunk_<c>.hap, checks its header and returns the descriptor in *hap_fd
// (the L*N genotype chars start at byte 16).  Returns "" on success, else the error text.
inline std::string load_chunk_small(const std::string &dir, int chunk, const char *painting, HostChunk &hc, int *hap_fd)
{
    const std::string base = dir + "/chunk_" + std::to_string(chunk);
    *hap_fd = -1;
    {
        const std::string p = dir + "/parameters_c" + std::to_string(chunk) + ".bin";
        FILE *fp = fopen(p.c_str(), "rb");
        if (!fp) return "cannot open " + p;
        int nb = 0;
        bool ok = fread(&hc.N, 4, 1, fp) == 1 && fread(&hc.L, 4, 1, fp) == 1 && fread(&nb, 4, 1, fp) == 1 && nb >= 2;
        if (ok) {
            hc.wb.resize(nb);
            ok = fread(hc.wb.data(), 4, nb, fp) == (size_t)nb;
        }
        fclose(fp);
        if (!ok) return "short read in " + p;
    }
    for (const char *ext : {".bp", ".dist", ".rpos", ".state"}) // the reference loader opens all six
        if (!file_exists(base + ext)) return "missing " + base + ext;
    {
        const std::string p = base + ".r";
        FILE *fp = fopen(p.c_str(), "rb");
        if (!fp) return "cannot open " + p;
        unsigned n = 0;
        bool ok = fread(&n, 4, 1, fp) == 1 && (int)n == hc.L;
        if (ok) {
            hc.r.resize(n);
            ok = fread(hc.r.data(), 8, n, fp) == n;
        }
        fclose(fp);
        if (!ok) return "short read in " + p;
    }
    hc.theta = 0.001;
    if (painting) { // both numbers are parsed as float (std::stof), Paint.cpp:47,56
        char *end = nullptr;
        hc.theta = (double)strtof(painting, &end);
        double rho = 1.0;
        if (end && *end == ',') rho = (double)strtof(end + 1, nullptr);
        for (auto &x : hc.r) x *= rho;
    }
    if (hc.wb.back() != hc.L || hc.wb.front() != 0) return "window boundaries do not span the chunk";
    {
        const std::string p = base + ".hap";
        const int fd = open(p.c_str(), O_RDONLY);
        if (fd < 0) return "cannot open " + p;
        uint64_t hdr[2] = {0, 0};
        if (pread(fd, hdr, 16, 0) != 16) {
            close(fd);
            return "short read in " + p;
        }
        if ((int)hdr[0] != hc.L || (int)hdr[1] != hc.N) {
            close(fd);
            return p + ": dimensions disagree with parameters file";
        }
        struct stat sb;
        if (fstat(fd, &sb) != 0 || (uint64_t)sb.st_size < 16 + hdr[0] * hdr[1]) {
            close(fd);
            return "short read in " + p;
        }
        *hap_fd = fd;
    }
    return "";
}

// Reads bytes [off, off+n) of the genotype matrix (file offset 16+off) into dst; false on a short read.
inline bool read_hap_range(int fd, size_t off, size_t n, char *dst)
{
    while (n > 0) {
        const ssize_t got = pread(fd, dst, n, (off_t)(16 + off));
        if (got <= 0) return false;
        dst += got;
        off += (size_t)got;
        n -= (size_t)got;
    }
    return true;
}

// alloc_hap(bytes) may hand out pinned memory for the L*N genotype chars (nullptr -> std::vector)
template <typename AllocHap>
inline std::string load_chunk_files(const std::string &dir, int chunk, const char *painting, HostChunk &hc,
                                    AllocHap alloc_hap)
{
    int fd = -1;
    std::string err = load_chunk_small(dir, chunk, painting, hc, &fd);
    if (!err.empty()) return err;
    const size_t nbytes = (size_t)hc.L * hc.N;
    hc.hap = alloc_hap(nbytes);
    if (!hc.hap) {
        hc.hap_own.resize(nbytes);
        hc.hap = hc.hap_own.data();
    }
    const bool ok = read_hap_range(fd, 0, nbytes, hc.hap);
    close(fd);
    if (!ok) return "short read in " + dir + "/chunk_" + std::to_string(chunk) + ".hap";
    return "";
}

inline std::string load_chunk_files(const std::string &dir, int chunk, const char *painting, HostChunk &hc)
{
    return load_chunk_files(dir, chunk, painting, hc, [](size_t) -> char * { return nullptr; });
}

// Run-length rule of the stepping-stone codec: v joins the run when
// fabs(head - v) < 1e-3 * min(head, v)   (float difference, double comparison).
inline int rle_encode(const float *v, int n, float *vals, int *lens)
{
    float head = v[0];
    int k = 0;
    vals[0] = head;
    lens[0] = 1;
    for (int j = 1; j < n; j++) {
        const float x = v[j];
        const float mn = std::min(head, x);
        if ((double)std::fabs(head - x) < 1e-3 * (double)mn) {
            lens[k]++;
        } else {
            head = x;
            vals[++k] = x;
            lens[k] = 1;
        }
    }
    return k + 1;
}

// appends one record to `out`:  size_t 1; size_t N; int site; float logscale; int K; float val[K]; int len[K]
inline void append_record(std::vector<char> &out, const float *v, int n, int site, float logscale,
                          std::vector<float> &vals, std::vector<int> &lens)
{
    vals.resize(n);
    lens.resize(n);
    const int k = rle_encode(v, n, vals.data(), lens.data());
    const uint64_t one = 1, sub = (uint64_t)n;
    const size_t at = out.size();
    out.resize(at + 28 + (size_t)8 * k);
    char *p = out.data() + at;
    memcpy(p, &one, 8);
    memcpy(p + 8, &sub, 8);
    memcpy(p + 16, &site, 4);
    memcpy(p + 20, &logscale, 4);
    memcpy(p + 24, &k, 4);
    memcpy(p + 28, vals.data(), (size_t)4 * k);
    memcpy(p + 28 + (size_t)4 * k, lens.data(), (size_t)4 * k);
}

// Reads chunk_<c>/paint/relate_<w>.bin into dense [N][N] alpha/beta (run heads expanded, as
// CollapsedMatrix<float>::ReadFromFile does, collapsed_matrix.hpp:268-296) and the per-target log-scales.
inline std::string decode_paint_file(const std::string &path, int N, int want_start, int want_end, float *alpha,
                                     float *beta, float *ls_alpha, float *ls_beta)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return "cannot open " + path;
    std::vector<float> vals;
    std::vector<int> lens;
    auto rec = [&](float *dst, float *ls) -> bool {
        uint64_t one = 0, sub = 0;
        int site = 0, k = 0;
        if (fread(&one, 8, 1, fp) != 1 || fread(&sub, 8, 1, fp) != 1 || one != 1 || (int)sub != N) return false;
        if (fread(&site, 4, 1, fp) != 1 || fread(ls, 4, 1, fp) != 1 || fread(&k, 4, 1, fp) != 1 || k < 1 || k > N) return false;
        vals.resize(k);
        lens.resize(k);
        if (fread(vals.data(), 4, k, fp) != (size_t)k || fread(lens.data(), 4, k, fp) != (size_t)k) return false;
        int i = 0;
        for (int j = 0; j < k; j++)
            for (int t = 0; t < lens[j]; t++) {
                if (i >= N) return false;
                dst[i++] = vals[j];
            }
        return i == N;
    };
    for (int n = 0; n < N; n++) {
        int a = 0, b = 0;
        bool ok = fread(&a, 4, 1, fp) == 1 && fread(&b, 4, 1, fp) == 1 && a == want_start && b == want_end &&
                  rec(alpha + (size_t)n * N, ls_alpha + n) && rec(beta + (size_t)n * N, ls_beta + n);
        if (!ok) {
            fclose(fp);
            return "malformed record for target " + std::to_string(n) + " in " + path;
        }
    }
    fclose(fp);
    return "";
}

} // namespace rp
