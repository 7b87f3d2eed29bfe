// make_chunks.hpp — the loader in front of the Paint path: SHAPEIT haps/sample + genetic map -> the chunk files
// `Relate --mode Paint` and every later stage read.  Restates, file for file and byte for byte,
//   Data::MakeChunks            /root/reference/include/src/data.cpp:117-518
//   haps::haps / haps::ReadSNP  /root/reference/include/src/data.hpp:128-162, data.cpp:544-573
//   map::map                    /root/reference/include/src/data.cpp:591-625
//   gzip::open                  /root/reference/include/src/data.cpp:7-60  (gzip input through `gunzip -c`)
// Outputs in <out>/: parameters.bin, props.bin, and per chunk c: parameters_c<c>.bin, chunk_<c>.{hap,state,bp,dist,rpos,r}.
// Host-only (text parsing and file layout); the painter's bit-packing of the genotype bytes happens on the GPU
// when a chunk is loaded (rp_chunk_load / rp_paint_chunk).
//
// Differences from the reference are confined to failure behaviour: where it assert()s or exit()s, this returns an
// error string (the C ABI never terminates the process), and the whole haps file is streamed once per pass with a
// large buffer instead of fscanf/fgets (same tokenisation: whitespace-separated fields, then the '0'/'1' characters
// of the rest of the line).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace rp {

struct MakeChunksInfo {
    int N = 0, L = 0, num_chunks = 0, max_windows = 0;
    double actual_min_memory_gb = 0;
    std::string warnings; // the lines the reference prints to stderr
};

namespace mc {

// gzip::open: a file is gzip if it starts 1f 8b 08, and is then read through `gunzip -c '<name>'`
struct InFile {
    FILE *fp = nullptr;
    bool piped = false;
    std::string open(const std::string &name)
    {
        FILE *chk = fopen(name.c_str(), "rb");
        if (!chk) return "Failed to open file " + name;
        unsigned char b[3] = {0, 0, 0};
        const size_t got = fread(b, 1, 3, chk);
        fclose(chk);
        const bool gz = got == 3 && b[0] == 0x1f && b[1] == 0x8b && b[2] == 0x08;
        if (gz) {
            const std::string cmd = "gunzip -c '" + name + "'";
            fp = popen(cmd.c_str(), "r");
            piped = true;
        } else {
            fp = fopen(name.c_str(), "r");
        }
        if (!fp) return "Failed to open file " + name;
        return "";
    }
    void close()
    {
        if (!fp) return;
        if (piped) pclose(fp);
        else fclose(fp);
        fp = nullptr;
    }
    ~InFile() { close(); }
};

inline bool is_space(int c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

// buffered reader with the two primitives the reference's parsers are made of
struct Reader {
    FILE *fp;
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    explicit Reader(FILE *f) : fp(f), buf(1 << 22) {}
    int peek()
    {
        if (pos == len) {
            len = fread(buf.data(), 1, buf.size(), fp);
            pos = 0;
            if (len == 0) return EOF;
        }
        return (unsigned char)buf[pos];
    }
    int get()
    {
        const int c = peek();
        if (c != EOF) pos++;
        return c;
    }
    // fscanf("%s"): skip white space, then the run of non-space characters; false at end of input
    bool token(std::string &out)
    {
        out.clear();
        int c;
        while ((c = peek()) != EOF && is_space(c)) pos++;
        if (c == EOF) return false;
        while ((c = peek()) != EOF && !is_space(c)) {
            out.push_back((char)c);
            pos++;
        }
        return true;
    }
    // fgets: the rest of the line including the '\n' (or up to end of input)
    void rest_of_line(std::string &out)
    {
        out.clear();
        for (;;) {
            if (pos == len && peek() == EOF) return;
            const char *b = buf.data() + pos;
            const char *nl = (const char *)memchr(b, '\n', len - pos);
            if (nl) {
                out.append(b, nl - b + 1);
                pos += nl - b + 1;
                return;
            }
            out.append(b, len - pos);
            pos = len;
        }
    }
};

inline bool parse_int(const std::string &s, int &v) // fscanf("%d")
{
    char *end = nullptr;
    const long x = strtol(s.c_str(), &end, 10);
    if (end == s.c_str()) return false;
    v = (int)x;
    return true;
}

inline std::string count_newlines(const std::string &name, long long &lines)
{
    InFile f;
    std::string e = f.open(name);
    if (!e.empty()) return e;
    std::vector<char> buf(1 << 22);
    lines = 0;
    size_t n;
    while ((n = fread(buf.data(), 1, buf.size(), f.fp)) > 0) {
        const char *p = buf.data(), *end = p + n;
        while ((p = (const char *)memchr(p, '\n', end - p)) != nullptr) {
            lines++;
            p++;
        }
    }
    return "";
}

template <typename T> inline bool put(FILE *fp, const T *p, size_t n) { return fwrite(p, sizeof(T), n, fp) == n; }

inline bool is_transition(const std::string &a, const std::string &b) // data.cpp:301-302, 333-334
{
    return (a == "C" && b == "T") || (a == "T" && b == "C") || (a == "G" && b == "A") || (a == "A" && b == "G");
}

} // namespace mc

// returns "" on success, else the error text
inline std::string make_chunks(const std::string &f_haps, const std::string &f_sample, const std::string &f_map,
                               const std::string &f_dist /* "unspecified" if none */, const std::string &out,
                               bool use_transitions, float min_memory, MakeChunksInfo *info)
{
    using namespace mc;
    std::ostringstream warn;
    // ---- haps::haps: N from the sample file, L from the number of lines of the haps file ----
    int N = 0, L = 0;
    {
        InFile f;
        std::string e = f.open(f_sample);
        if (!e.empty()) return e;
        Reader rd(f.fp);
        std::string a, b, c;
        for (int h = 0; h < 2; h++)
            if (!(rd.token(a) && rd.token(b) && rd.token(c))) return "sample file " + f_sample + ": missing header";
        while (rd.token(a) && rd.token(b) && rd.token(c)) N += (a == b) ? 2 : 1;
    }
    {
        long long lines = 0;
        std::string e = count_newlines(f_haps, lines);
        if (!e.empty()) return e;
        L = (int)lines;
    }
    if (N < 1 || L < 1) return "empty haps/sample input";
    const std::vector<char>::size_type uN = (std::vector<char>::size_type)N;

    std::vector<int> bp_pos((size_t)L + 1);
    std::vector<std::string> ancestral(L), alternative(L), rsid(L);

    double min_memory_size = (min_memory)*1e9 / 4.0 - (2 * N * N + 3 * N), actual_min_memory_size = 0.0;
    if (min_memory_size <= 0) return "Error: Need larger memory allowance.";
    const int windows_per_section = 500;
    int max_windows_per_section = 0;
    const int overlap = 20000;
    int max_chunk_size = std::min(L + 1, (int)(min_memory_size / N));
    if (min_memory >= 100) max_chunk_size = 2500000;

    // the reference keeps max_chunk_size rows of N chars; rows are allocated here as they are filled
    std::vector<std::vector<char>> p_seq, p_overlap;
    std::vector<int> window_boundaries(windows_per_section + 1), window_boundaries_overlap(windows_per_section + 1);
    std::vector<int> section_boundary_start, section_boundary_end;
    section_boundary_start.push_back(0);
    int state_val = 1;
    int min_snps_in_window = max_chunk_size;
    float mean_snps_in_window = 0.0;
    int num_windows = 0, num_windows_overlap = 0;
    int overlap_in_section = 0;
    int chunk_size = 0;
    int chunk_index = 0;
    double window_memory_size = 0.0;

    InFile hf;
    {
        std::string e = hf.open(f_haps);
        if (!e.empty()) return e;
    }
    Reader hr(hf.fp);
    std::string tok, line, chr;

    auto state_of = [&](int s) -> int {
        if (use_transitions) return state_val;
        state_val = is_transition(ancestral[s], alternative[s]) ? 0 : 1;
        return state_val;
    };

    int snp = 0;
    while (snp < L) {
        const std::string base = out + "/chunk_" + std::to_string(chunk_index);
        FILE *fp_haps_chunk = fopen((base + ".hap").c_str(), "wb");
        FILE *fp_state = fopen((base + ".state").c_str(), "wb");
        if (!fp_haps_chunk || !fp_state) {
            if (fp_haps_chunk) fclose(fp_haps_chunk);
            if (fp_state) fclose(fp_state);
            return "cannot create " + base + ".hap/.state";
        }
        auto bail = [&](const std::string &msg) {
            fclose(fp_haps_chunk);
            fclose(fp_state);
            return msg;
        };

        if (snp > 0) { // data.cpp:166-194: the last `overlap` SNPs of the previous chunk open this one
            if (snp - section_boundary_start.back() < overlap) return bail("chunk shorter than the 20000-SNP overlap (increase --memory)");
            overlap_in_section = overlap;
            if (overlap_in_section > chunk_size) return bail("overlap exceeds the chunk size");
            const int snp_section_begin = snp - overlap_in_section;
            section_boundary_start.push_back(snp_section_begin);
            p_overlap.assign(p_seq.begin() + (chunk_size - overlap_in_section), p_seq.begin() + chunk_size);
            int *wo = window_boundaries_overlap.data();
            wo[0] = snp_section_begin;
            num_windows_overlap = 1;
            for (int i = 0; i < num_windows; i++)
                if (window_boundaries[i] > snp_section_begin) wo[num_windows_overlap++] = window_boundaries[i];
            if (!(num_windows_overlap < windows_per_section - 1)) return bail("too many windows in the overlap");
        }

        const int snp_begin = snp;
        window_memory_size = 0.0;
        chunk_size = 0;
        window_boundaries[0] = snp_begin;
        num_windows = 1;
        int snps_in_window = 0;
        while (num_windows + num_windows_overlap < windows_per_section && chunk_size < max_chunk_size && snp < L) {
            // haps::ReadSNP: "%s %s %d %s %s", then the '0'/'1' characters of the rest of the line
            std::string s_bp;
            if (!(hr.token(chr) && hr.token(rsid[snp]) && hr.token(s_bp) && hr.token(ancestral[snp]) && hr.token(alternative[snp])) ||
                !parse_int(s_bp, bp_pos[snp]))
                return bail("haps file: malformed line " + std::to_string(snp + 1));
            hr.rest_of_line(line);
            if ((int)p_seq.size() <= chunk_size) p_seq.emplace_back(N);
            std::vector<char> &row = p_seq[chunk_size];
            row.resize(N);
            int n = 0, num_derived = 0;
            for (size_t i = 0; i < line.size() && n < N; i++) {
                const char d = line[i];
                if (d == '0') row[n++] = '0';
                else if (d == '1') {
                    row[n++] = '1';
                    num_derived++;
                }
            }
            if (n != N) return bail("haps file: " + chr + " " + rsid[snp] + " " + std::to_string(bp_pos[snp]) + ": fewer than N alleles");

            window_memory_size += num_derived * (N + 1);
            if (window_memory_size >= min_memory_size && snps_in_window > 10) {
                if (actual_min_memory_size < window_memory_size) actual_min_memory_size = window_memory_size;
                if (min_snps_in_window > snps_in_window) min_snps_in_window = snps_in_window;
                snps_in_window = 0;
                window_memory_size = 0.0;
                window_boundaries[num_windows] = snp;
                num_windows++;
            }
            snp++;
            snps_in_window++;
            chunk_size++;
        }
        if (actual_min_memory_size < window_memory_size) actual_min_memory_size = window_memory_size;
        if (min_snps_in_window > snps_in_window) min_snps_in_window = snps_in_window;
        mean_snps_in_window = chunk_size / num_windows;
        window_boundaries[num_windows] = snp;
        if (num_windows > max_windows_per_section) max_windows_per_section = num_windows;
        if (mean_snps_in_window < 100) {
            warn << "Memory allowance should be set " << 100 / mean_snps_in_window << " times larger than\n";
            warn << "the current setting using --memory (Default 5GB).\n";
        }
        section_boundary_end.push_back(snp);

        int snp_tmp = section_boundary_start.back();
        bool ok = true;
        {
            const std::string pp = out + "/parameters_c" + std::to_string(chunk_index) + ".bin";
            FILE *fp = fopen(pp.c_str(), "w");
            if (!fp) return bail("cannot create " + pp);
            if (snp_begin == 0) {
                const std::vector<char>::size_type uL_chunk = (std::vector<char>::size_type)chunk_size;
                ok = ok && put(fp_haps_chunk, &uL_chunk, 1) && put(fp_haps_chunk, &uN, 1);
                const int num_windows_in_section = num_windows + 1;
                ok = ok && put(fp, &N, 1) && put(fp, &chunk_size, 1) && put(fp, &num_windows_in_section, 1) &&
                     put(fp, window_boundaries.data(), (size_t)num_windows_in_section);
                ok = ok && put(fp_state, &chunk_size, 1);
            } else {
                const int L_chunk = chunk_size + overlap_in_section;
                const std::vector<char>::size_type uL_chunk = (std::vector<char>::size_type)L_chunk;
                ok = ok && put(fp_haps_chunk, &uL_chunk, 1) && put(fp_haps_chunk, &uN, 1);
                const int window_start = window_boundaries_overlap[0];
                std::vector<int> wo(window_boundaries_overlap.begin(), window_boundaries_overlap.begin() + num_windows_overlap);
                std::vector<int> wn(window_boundaries.begin(), window_boundaries.begin() + num_windows + 1);
                for (int &x : wo) x -= window_start;
                for (int &x : wn) x -= window_start;
                const int num_windows_in_section = num_windows + num_windows_overlap + 1;
                ok = ok && put(fp, &N, 1) && put(fp, &L_chunk, 1) && put(fp, &num_windows_in_section, 1) &&
                     put(fp, wo.data(), wo.size()) && put(fp, wn.data(), wn.size());
                ok = ok && put(fp_state, &L_chunk, 1);
                for (int i = 0; i < overlap_in_section && ok; i++) {
                    const int sv = state_of(snp_tmp);
                    snp_tmp++;
                    ok = put(fp_state, &sv, 1) && put(fp_haps_chunk, p_overlap[i].data(), (size_t)N);
                }
            }
            fclose(fp);
        }
        for (int i = 0; i < chunk_size && ok; i++) {
            const int sv = state_of(snp_tmp);
            snp_tmp++;
            ok = put(fp_state, &sv, 1) && put(fp_haps_chunk, p_seq[i].data(), (size_t)N);
        }
        fclose(fp_haps_chunk);
        fclose(fp_state);
        if (!ok) return "short write to " + base + ".hap/.state";
        chunk_index++;
    }
    bp_pos[L] = bp_pos[L - 1] + 1;
    hf.close();
    p_seq.clear();
    p_overlap.clear();

    const int num_chunks = (int)section_boundary_start.size();
    {
        std::ostringstream w;
        w << std::setprecision(2) << "Warning: Will use min " << 2.0 * (4.0 * N * N * (max_windows_per_section + 2.0)) / 1e9
          << "GB of hard disc.\n";
        warn << w.str();
    }
    {
        FILE *fp = fopen((out + "/parameters.bin").c_str(), "w");
        if (!fp) return "cannot create " + out + "/parameters.bin";
        actual_min_memory_size += (2 * N * N + 3 * N);
        actual_min_memory_size *= 4.0 / 1e9;
        const bool ok = put(fp, &N, 1) && put(fp, &L, 1) && put(fp, &num_chunks, 1) && put(fp, &actual_min_memory_size, 1) &&
                        put(fp, section_boundary_start.data(), (size_t)num_chunks) &&
                        put(fp, section_boundary_end.data(), (size_t)num_chunks);
        fclose(fp);
        if (!ok) return "short write to parameters.bin";
    }

    // ---- dist (data.cpp:385-425) ----
    std::vector<int> dist(L);
    if (f_dist == "unspecified") {
        for (int s = 0; s + 1 < L; s++) {
            dist[s] = bp_pos[s + 1] - bp_pos[s];
            if (dist[s] <= 0)
                return "Failed at BP " + std::to_string(bp_pos[s]) + "\nSNPs are not sorted by bp or more than one SNP at same position.";
        }
        dist[L - 1] = 1;
    } else {
        InFile f;
        std::string e = f.open(f_dist);
        if (!e.empty()) return e;
        Reader rd(f.fp);
        std::string a, b;
        rd.token(a);
        rd.token(b);
        int s = 0;
        while (rd.token(a) && rd.token(b)) {
            int mbp = 0, mdist = 0;
            if (!parse_int(a, mbp) || !parse_int(b, mdist)) break;
            if (s >= L) return "dist file has more lines than the haps file";
            if (bp_pos[s] != mbp) return "dist file: position " + std::to_string(mbp) + " does not match the haps file";
            dist[s++] = mdist;
        }
    }

    // ---- props.bin (data.cpp:427-449) ----
    {
        FILE *fp = fopen((out + "/props.bin").c_str(), "wb");
        if (!fp) return "cannot create " + out + "/props.bin";
        std::vector<char> rec(12 + 3 * 1024);
        bool ok = true;
        for (int s = 0; s < L && ok; s++) {
            std::fill(rec.begin(), rec.end(), 0);
            memcpy(rec.data(), &s, 4);
            memcpy(rec.data() + 4, &bp_pos[s], 4);
            memcpy(rec.data() + 8, &dist[s], 4);
            memcpy(rec.data() + 12, rsid[s].c_str(), std::min<size_t>(rsid[s].size(), 1023));
            memcpy(rec.data() + 12 + 1024, ancestral[s].c_str(), std::min<size_t>(ancestral[s].size(), 1023));
            memcpy(rec.data() + 12 + 2048, alternative[s].c_str(), std::min<size_t>(alternative[s].size(), 1023));
            ok = put(fp, rec.data(), rec.size());
        }
        fclose(fp);
        if (!ok) return "short write to props.bin";
    }

    // ---- genetic map -> rpos, r (data.cpp:451-481, map::map 591-625) ----
    std::vector<int> mbp;
    std::vector<double> mgen;
    {
        long long lines = 0;
        std::string e = count_newlines(f_map, lines);
        if (!e.empty()) return e;
        lines--; // header
        if (lines < 2) return "genetic map " + f_map + " needs at least two rows";
        InFile f;
        e = f.open(f_map);
        if (!e.empty()) return e;
        Reader rd(f.fp);
        std::string a, b, c;
        rd.token(a);
        rd.token(b);
        rd.token(c);
        mbp.resize(lines);
        mgen.resize(lines);
        for (long long s = 0; s < lines; s++) { // "%lf %f %lf"
            if (!(rd.token(a) && rd.token(b) && rd.token(c))) return "genetic map " + f_map + ": short row " + std::to_string(s + 1);
            mbp[s] = (int)strtod(a.c_str(), nullptr);
            mgen[s] = strtod(c.c_str(), nullptr);
        }
    }
    std::vector<double> r(L), rpos((size_t)L + 1);
    {
        size_t ir = 0, ib = 0;
        size_t map_pos = 0;
        if (mbp[map_pos] > bp_pos[ib]) {
            rpos[ir++] = mgen[map_pos] * 1e-2;
            ib++;
        }
        for (; ir < rpos.size();) {
            while (mbp[map_pos + 1] <= bp_pos[ib] && map_pos < mbp.size() - 2) map_pos++;
            if (mbp[map_pos + 1] - mbp[map_pos] < 0) return "genetic map is not sorted by position";
            if (mbp[map_pos + 1] - mbp[map_pos] == 0 || mbp[map_pos] > bp_pos[ib]) {
                rpos[ir] = mgen[map_pos] * 1e-2;
            } else {
                rpos[ir] = ((bp_pos[ib] - mbp[map_pos]) / ((double)(mbp[map_pos + 1] - mbp[map_pos])) * (mgen[map_pos + 1] - mgen[map_pos]) +
                            mgen[map_pos]) *
                           1e-2;
            }
            ir++;
            ib++;
        }
        const double lower_bound = 1e-10; // data.cpp:4
        for (int s = 0; s < L; s++) {
            r[s] = rpos[s + 1] - rpos[s];
            if (r[s] < lower_bound) r[s] = lower_bound;
            r[s] *= 2500;
        }
    }

    // ---- per-chunk position files (data.cpp:485-516) ----
    for (int c = 0; c < num_chunks; c++) {
        const std::string base = out + "/chunk_" + std::to_string(c);
        const int s0 = section_boundary_start[c];
        const unsigned int L_chunk = (unsigned)(section_boundary_end[c] - s0), L1 = L_chunk + 1;
        FILE *fp_pos = fopen((base + ".bp").c_str(), "wb"), *fp_dist = fopen((base + ".dist").c_str(), "wb");
        FILE *fp_rpos = fopen((base + ".rpos").c_str(), "wb"), *fp_r = fopen((base + ".r").c_str(), "wb");
        bool ok = fp_pos && fp_dist && fp_rpos && fp_r;
        ok = ok && put(fp_pos, &L_chunk, 1) && put(fp_dist, &L_chunk, 1) && put(fp_rpos, &L1, 1) && put(fp_r, &L_chunk, 1);
        ok = ok && put(fp_pos, &bp_pos[s0], L_chunk) && put(fp_dist, &dist[s0], L_chunk) && put(fp_rpos, &rpos[s0], (size_t)L_chunk + 1) &&
             put(fp_r, &r[s0], L_chunk);
        for (FILE *f : {fp_pos, fp_dist, fp_rpos, fp_r})
            if (f) fclose(f);
        if (!ok) return "cannot write the position files of chunk " + std::to_string(c);
    }

    if (info) {
        info->N = N;
        info->L = L;
        info->num_chunks = num_chunks;
        info->max_windows = max_windows_per_section;
        info->actual_min_memory_gb = actual_min_memory_size;
        info->warnings = warn.str();
    }
    return "";
}

} // namespace rp
