// make_chunks.hpp — the loader in front of the Paint path: SHAPEIT haps/sample + genetic map -> the chunk files that
// `Relate --mode Paint` and every later stage read.
//
// What it must reproduce, byte for byte, is what the reference's `Relate --mode MakeChunks` leaves in <out>/
// (Data::MakeChunks, /root/reference/include/src/data.cpp:117-518; haps / map readers data.hpp:128-162,
// data.cpp:544-625): parameters.bin, props.bin and per chunk parameters_c<c>.bin, chunk_<c>.{hap,state,bp,dist,rpos,r}.
// How it gets there is this repo's own three passes:
//
//   1. scan   one streaming pass over the haps text: per SNP its position / ids / alleles and the genotype row,
//             bit-packed straight into the painter's HBM layout (bit n&31 of word n>>5, rows padded to 16 bytes);
//             the whole data set stays resident that way (N*L/8 bytes: 125 MB at N = 10 000 x L = 100 000);
//   2. plan   a pure function of the per-SNP derived-allele counts: chunk extents (with the 20 000-SNP overlap) and
//             window boundaries under the --memory budget (the rules of data.cpp:129-231, see plan_chunks);
//   3. emit   per chunk, the files; rows are expanded back to the reference's one-char-per-allele format, and —
//             optionally — also written as they are to `chunk_<c>.hapbits`, which rp_paint_chunk reads instead of the
//             8x larger .hap (and deletes once it has painted the chunk: the reference's Finalize refuses to remove a
//             directory that still holds a file it does not know).
//
// Host-only.  Failures come back as an error string (the C ABI never terminates the process).
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

// header of chunk_<c>.hapbits (little-endian): rows follow, `wps` 32-bit words each
struct HapBitsHeader {
    char magic[8]; // "RPHBITS1"
    int N, L, wps, reserved;
};
inline const char *hapbits_magic() { return "RPHBITS1"; }

namespace mc {

// ---- input: plain or gzip (through `gunzip -c`, as the reference does it), read in large blocks, handed out by line ----
class TextFile {
  public:
    ~TextFile() { close(); }
    std::string open(const std::string &name)
    {
        close();
        FILE *probe = fopen(name.c_str(), "rb");
        if (!probe) return "Failed to open file " + name;
        unsigned char sig[3] = {0, 0, 0};
        const size_t got = fread(sig, 1, 3, probe);
        fclose(probe);
        piped_ = got == 3 && sig[0] == 0x1f && sig[1] == 0x8b && sig[2] == 0x08;
        fp_ = piped_ ? popen(("gunzip -c '" + name + "'").c_str(), "r") : fopen(name.c_str(), "rb");
        if (!fp_) return "Failed to open file " + name;
        block_.resize(8u << 20);
        have_ = used_ = 0;
        return "";
    }
    void close()
    {
        if (fp_) (piped_ ? pclose : fclose)(fp_);
        fp_ = nullptr;
    }
    // next line without its '\n' (a final line without one is delivered too); *terminated tells which it was
    bool line(const char *&begin, const char *&end, bool *terminated = nullptr)
    {
        for (;;) {
            const char *b = block_.data() + used_;
            const char *nl = static_cast<const char *>(memchr(b, '\n', have_ - used_));
            if (nl) {
                begin = b;
                end = nl;
                used_ = (size_t)(nl - block_.data()) + 1;
                if (terminated) *terminated = true;
                return true;
            }
            if (eof_) {
                if (used_ == have_) return false;
                begin = b;
                end = block_.data() + have_;
                used_ = have_;
                if (terminated) *terminated = false;
                return true;
            }
            // keep the partial line, refill behind it
            const size_t keep = have_ - used_;
            memmove(block_.data(), block_.data() + used_, keep);
            used_ = 0;
            have_ = keep;
            if (have_ == block_.size()) block_.resize(block_.size() * 2);
            const size_t got = fread(block_.data() + have_, 1, block_.size() - have_, fp_);
            have_ += got;
            if (got == 0) eof_ = true;
        }
    }

  private:
    FILE *fp_ = nullptr;
    bool piped_ = false, eof_ = false;
    std::vector<char> block_;
    size_t have_ = 0, used_ = 0;
};

inline bool blank(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f' || c == '\r' || c == '\n'; }

// next whitespace-separated field of [p, end); false if there is none
inline bool field(const char *&p, const char *end, const char *&fb, const char *&fe)
{
    while (p < end && blank(*p)) p++;
    if (p == end) return false;
    fb = p;
    while (p < end && !blank(*p)) p++;
    fe = p;
    return true;
}

// all fields of a whole file, for the small inputs (sample, dist, map)
inline std::string all_fields(const std::string &name, std::vector<std::string> &out, long long *newlines = nullptr)
{
    TextFile f;
    std::string e = f.open(name);
    if (!e.empty()) return e;
    const char *b, *en, *fb, *fe;
    bool term = false;
    long long nl = 0;
    while (f.line(b, en, &term)) {
        nl += term;
        while (field(b, en, fb, fe)) out.emplace_back(fb, fe);
    }
    if (newlines) *newlines = nl;
    return "";
}

// ---- pass 1 result: per-SNP columns + the bit-packed genotype rows ----
struct SnpTable {
    int N = 0, L = 0, wps = 0;
    std::vector<int> bp;                      // L+1 entries (the last one is bp[L-1]+1, data.cpp:351)
    std::vector<int> derived;                 // derived alleles per SNP
    std::vector<std::string> rsid, anc, alt;
    std::vector<uint32_t> bits;               // L rows of wps words
    const uint32_t *row(int s) const { return bits.data() + (size_t)s * wps; }
};

inline int count_samples(const std::vector<std::string> &f) // haps::haps: two header rows, then ID_1 ID_2 missing
{
    int n = 0;
    for (size_t i = 6; i + 3 <= f.size(); i += 3) n += (f[i] == f[i + 1]) ? 2 : 1;
    return n;
}

inline std::string scan_haps(const std::string &f_haps, int N, SnpTable &t)
{
    TextFile f;
    std::string e = f.open(f_haps);
    if (!e.empty()) return e;
    t.N = N;
    t.wps = ((N + 31) / 32 + 3) / 4 * 4;
    const char *b, *en, *fb, *fe;
    bool term = false;
    long long newlines = 0;
    std::vector<uint32_t> rowbits((size_t)t.wps);
    while (f.line(b, en, &term)) {
        newlines += term;
        const char *p = b;
        if (!field(p, en, fb, fe)) continue; // blank line
        const int s = (int)t.bp.size();
        std::string id, a0, a1;
        int pos = 0;
        bool ok = field(p, en, fb, fe);
        if (ok) id.assign(fb, fe);
        ok = ok && field(p, en, fb, fe);
        if (ok) {
            char *stop = nullptr;
            const std::string num(fb, fe);
            pos = (int)strtol(num.c_str(), &stop, 10);
            ok = stop != num.c_str();
        }
        ok = ok && field(p, en, fb, fe);
        if (ok) a0.assign(fb, fe);
        ok = ok && field(p, en, fb, fe);
        if (ok) a1.assign(fb, fe);
        if (!ok) return "haps file: malformed line " + std::to_string(s + 1);
        // the alleles: every '0' / '1' character of the rest of the line, anything else is a separator
        std::fill(rowbits.begin(), rowbits.end(), 0u);
        int n = 0, ones = 0;
        for (; p < en && n < N; p++) {
            const unsigned d = (unsigned)(*p - '0');
            if (d <= 1u) {
                rowbits[n >> 5] |= d << (n & 31);
                ones += (int)d;
                n++;
            }
        }
        if (n != N) return "haps file: " + std::string(b, std::min(en, b + 60)) + "...: fewer than N alleles";
        t.bp.push_back(pos);
        t.derived.push_back(ones);
        t.rsid.push_back(std::move(id));
        t.anc.push_back(std::move(a0));
        t.alt.push_back(std::move(a1));
        t.bits.insert(t.bits.end(), rowbits.begin(), rowbits.end());
    }
    // the reference takes L from the number of '\n' characters; a data line without one would be read but not counted
    t.L = (int)std::min<long long>(newlines, (long long)t.bp.size());
    if (t.L < 1) return "empty haps/sample input";
    t.bp.resize((size_t)t.L);
    t.bp.push_back(t.bp[(size_t)t.L - 1] + 1);
    return "";
}

// ---- pass 2: the plan ----
struct ChunkPlan {
    int file_begin = 0; // first SNP stored in the chunk's files (= new_begin - overlap for every chunk but the first)
    int new_begin = 0;  // first SNP that no earlier chunk holds
    int end = 0;        // one past the last SNP
    std::vector<int> bounds; // window boundaries, absolute SNP indices: file_begin, ..., end
    int new_windows = 0;     // windows opened by this chunk's own SNPs
};

struct PlanTotals {
    double peak_window_floats = 0; // largest  sum over a window of derived*(N+1)
    int max_new_windows = 0;
    std::string warnings;
};

// Rules (data.cpp:129-231), stated on the derived-allele counts alone:
//  * budget = memory*1e9/4 - (2N^2+3N) floats.  A window closes AT the SNP whose derived*(N+1) brings the running sum to
//    the budget, provided the window already holds more than 10 SNPs; that SNP opens the next window (its own
//    contribution is not carried over).
//  * a chunk takes new SNPs while it has fewer than 500 boundaries (the ones inherited through the overlap included),
//    fewer than `cap` new SNPs (cap = min(L+1, budget/N); 2 500 000 when memory >= 100) and SNPs remain.
//  * every chunk but the first starts 20 000 SNPs early; it inherits the previous chunk's boundaries that lie beyond
//    that point.  The previous chunk must be able to supply the overlap.
inline std::string plan_chunks(const std::vector<int> &derived, int N, float memory_gb, std::vector<ChunkPlan> &plan, PlanTotals &tot)
{
    constexpr int kBoundaryCap = 500, kOverlap = 20000;
    const int L = (int)derived.size();
    const double fixed = (double)(2LL * N * N + 3LL * N);
    const double budget = (double)memory_gb * 1e9 / 4.0 - fixed;
    if (budget <= 0) return "Error: Need larger memory allowance.";
    int cap = std::min(L + 1, (int)(budget / N));
    if (memory_gb >= 100) cap = 2500000;
    std::ostringstream warn;
    int s = 0;
    std::vector<int> prev_opened; // boundaries the previous chunk opened itself (its end not included)
    while (s < L) {
        ChunkPlan c;
        c.new_begin = c.file_begin = s;
        int inherited = 0;
        if (!plan.empty()) {
            const ChunkPlan &p = plan.back();
            if (s - p.file_begin < kOverlap) return "chunk shorter than the 20000-SNP overlap (increase --memory)";
            if (s - p.new_begin < kOverlap) return "overlap exceeds the chunk size";
            c.file_begin = s - kOverlap;
            c.bounds.push_back(c.file_begin);
            for (int b : prev_opened)
                if (b > c.file_begin) c.bounds.push_back(b);
            inherited = (int)c.bounds.size();
            if (!(inherited < kBoundaryCap - 1)) return "too many windows in the overlap";
        }
        std::vector<int> opened{s};
        double load = 0.0;
        int in_window = 0, taken = 0;
        while ((int)opened.size() + inherited < kBoundaryCap && taken < cap && s < L) {
            load += (double)(derived[s] * (N + 1));
            if (load >= budget && in_window > 10) {
                tot.peak_window_floats = std::max(tot.peak_window_floats, load);
                opened.push_back(s);
                load = 0.0;
                in_window = 0;
            }
            s++;
            in_window++;
            taken++;
        }
        tot.peak_window_floats = std::max(tot.peak_window_floats, load);
        c.end = s;
        c.new_windows = (int)opened.size();
        c.bounds.insert(c.bounds.end(), opened.begin(), opened.end());
        c.bounds.push_back(s);
        tot.max_new_windows = std::max(tot.max_new_windows, c.new_windows);
        const float mean = (float)(taken / c.new_windows); // (integer division, as the reference reports it)
        if (mean < 100) {
            warn << "Memory allowance should be set " << 100 / mean << " times larger than\n";
            warn << "the current setting using --memory (Default 5GB).\n";
        }
        prev_opened = std::move(opened);
        plan.push_back(std::move(c));
    }
    tot.warnings = warn.str();
    return "";
}

// ---- pass 3: files ----
class OutFile {
  public:
    OutFile(const std::string &path, const char *mode = "wb") : path_(path), fp_(fopen(path.c_str(), mode))
    {
        if (fp_) setvbuf(fp_, nullptr, _IOFBF, 1 << 20);
    }
    ~OutFile()
    {
        if (fp_) fclose(fp_);
    }
    template <typename T> OutFile &put(const T &v) { return raw(&v, sizeof(T)); }
    template <typename T> OutFile &put(const T *p, size_t n) { return raw(p, n * sizeof(T)); }
    OutFile &raw(const void *p, size_t bytes)
    {
        if (fp_ && bytes && fwrite(p, 1, bytes, fp_) != bytes) bad_ = true;
        return *this;
    }
    std::string finish() // "" or what went wrong
    {
        if (!fp_) return "cannot create " + path_;
        const bool bad = bad_ || fclose(fp_) != 0;
        fp_ = nullptr;
        return bad ? "short write to " + path_ : "";
    }

  private:
    std::string path_;
    FILE *fp_;
    bool bad_ = false;
};

inline bool transition(const std::string &a, const std::string &b) // data.cpp:301-302, 333-334
{
    return (a == "C" && b == "T") || (a == "T" && b == "C") || (a == "G" && b == "A") || (a == "A" && b == "G");
}

// bits of one row -> N chars '0' / '1'
inline void expand_row(const uint32_t *w, int N, char *out)
{
    for (int n = 0; n < N; n += 32) {
        uint32_t x = w[n >> 5];
        const int m = std::min(32, N - n);
        for (int j = 0; j < m; j++, x >>= 1) out[n + j] = (char)('0' + (x & 1u));
    }
}

inline std::string emit_chunk(const std::string &out, int index, const ChunkPlan &c, const SnpTable &t, bool all_states_one,
                              bool write_hapbits)
{
    const std::string base = out + "/chunk_" + std::to_string(index);
    const int N = t.N, Lc = c.end - c.file_begin;
    {
        OutFile par(out + "/parameters_c" + std::to_string(index) + ".bin");
        std::vector<int> rel(c.bounds);
        for (int &b : rel) b -= c.file_begin;
        par.put(N).put(Lc).put((int)rel.size()).put(rel.data(), rel.size());
        const std::string e = par.finish();
        if (!e.empty()) return e;
    }
    {
        OutFile st(base + ".state");
        st.put(Lc);
        for (int s = c.file_begin; s < c.end; s++) st.put<int>(all_states_one ? 1 : (transition(t.anc[s], t.alt[s]) ? 0 : 1));
        const std::string e = st.finish();
        if (!e.empty()) return e;
    }
    {
        OutFile hap(base + ".hap");
        hap.put((size_t)Lc).put((size_t)N);
        std::vector<char> row((size_t)N);
        for (int s = c.file_begin; s < c.end; s++) {
            expand_row(t.row(s), N, row.data());
            hap.raw(row.data(), row.size());
        }
        const std::string e = hap.finish();
        if (!e.empty()) return e;
    }
    if (write_hapbits) {
        OutFile hb(base + ".hapbits");
        HapBitsHeader h{};
        memcpy(h.magic, hapbits_magic(), 8);
        h.N = N;
        h.L = Lc;
        h.wps = t.wps;
        hb.put(h).put(t.row(c.file_begin), (size_t)Lc * t.wps);
        const std::string e = hb.finish();
        if (!e.empty()) return e;
    }
    return "";
}

// genetic map (data.cpp:591-625) with the cursor the interpolation of data.cpp:451-470 walks along it
class GeneticMap {
  public:
    std::string load(const std::string &name)
    {
        std::vector<std::string> f;
        long long newlines = 0;
        std::string e = all_fields(name, f, &newlines);
        if (!e.empty()) return e;
        const long long rows = newlines - 1; // the header line is not data
        if (rows < 2) return "genetic map " + name + " needs at least two rows";
        if ((long long)f.size() < 3 + 3 * rows) return "genetic map " + name + ": short row " + std::to_string((f.size() - 3) / 3 + 1);
        bp_.resize((size_t)rows);
        cm_.resize((size_t)rows);
        for (long long i = 0; i < rows; i++) { // position (read as a double, kept as an int), rate (unused), cM
            bp_[(size_t)i] = (int)strtod(f[(size_t)(3 + 3 * i)].c_str(), nullptr);
            cm_[(size_t)i] = strtod(f[(size_t)(5 + 3 * i)].c_str(), nullptr);
        }
        for (size_t i = 0; i + 1 < bp_.size(); i++)
            if (bp_[i + 1] < bp_[i]) return "genetic map is not sorted by position";
        return "";
    }
    // Morgans at base pair `pos`; calls must come with non-decreasing pos
    double morgans(int pos)
    {
        while (bp_[at_ + 1] <= pos && at_ < bp_.size() - 2) at_++;
        const int span = bp_[at_ + 1] - bp_[at_];
        if (span == 0 || bp_[at_] > pos) return cm_[at_] * 1e-2;
        return ((pos - bp_[at_]) / ((double)span) * (cm_[at_ + 1] - cm_[at_]) + cm_[at_]) * 1e-2;
    }

  private:
    std::vector<int> bp_;
    std::vector<double> cm_;
    size_t at_ = 0;
};

} // namespace mc

// returns "" on success, else the error text
inline std::string make_chunks(const std::string &f_haps, const std::string &f_sample, const std::string &f_map,
                               const std::string &f_dist /* "unspecified" if none */, const std::string &out,
                               bool use_transitions, float min_memory, MakeChunksInfo *info, bool write_hapbits = false)
{
    using namespace mc;
    // ---- scan ----
    int N = 0;
    {
        std::vector<std::string> f;
        std::string e = all_fields(f_sample, f);
        if (!e.empty()) return e;
        if (f.size() < 6) return "sample file " + f_sample + ": missing header";
        N = count_samples(f);
    }
    if (N < 1) return "empty haps/sample input";
    SnpTable t;
    {
        std::string e = scan_haps(f_haps, N, t);
        if (!e.empty()) return e;
    }
    const int L = t.L;

    // ---- plan ----
    std::vector<ChunkPlan> plan;
    PlanTotals tot;
    {
        std::vector<int> derived(t.derived.begin(), t.derived.begin() + L);
        std::string e = plan_chunks(derived, N, min_memory, plan, tot);
        if (!e.empty()) return e;
    }
    const int num_chunks = (int)plan.size();

    // ---- emit: genotype / state / window files per chunk ----
    for (int c = 0; c < num_chunks; c++) {
        std::string e = emit_chunk(out, c, plan[c], t, use_transitions, write_hapbits);
        if (!e.empty()) return e;
    }
    std::ostringstream warn;
    warn << tot.warnings;
    warn << std::setprecision(2) << "Warning: Will use min " << 2.0 * (4.0 * N * N * (tot.max_new_windows + 2.0)) / 1e9 << "GB of hard disc.\n";
    const double peak_gb = (tot.peak_window_floats + (double)(2LL * N * N + 3LL * N)) * (4.0 / 1e9);
    {
        OutFile par(out + "/parameters.bin", "w");
        par.put(N).put(L).put(num_chunks).put(peak_gb);
        for (const ChunkPlan &c : plan) par.put(c.file_begin);
        for (const ChunkPlan &c : plan) par.put(c.end);
        std::string e = par.finish();
        if (!e.empty()) return e;
    }

    // ---- distances between SNPs (data.cpp:385-425) ----
    std::vector<int> dist((size_t)L);
    if (f_dist == "unspecified") {
        for (int s = 0; s + 1 < L; s++) {
            dist[s] = t.bp[s + 1] - t.bp[s];
            if (dist[s] <= 0)
                return "Failed at BP " + std::to_string(t.bp[s]) + "\nSNPs are not sorted by bp or more than one SNP at same position.";
        }
        dist[L - 1] = 1;
    } else {
        std::vector<std::string> f;
        std::string e = all_fields(f_dist, f);
        if (!e.empty()) return e;
        int s = 0;
        for (size_t i = 2; i + 1 < f.size(); i += 2, s++) { // after the two header fields: position, distance
            char *stop = nullptr;
            const int pos = (int)strtol(f[i].c_str(), &stop, 10);
            if (stop == f[i].c_str()) break;
            if (s >= L) return "dist file has more lines than the haps file";
            if (t.bp[s] != pos) return "dist file: position " + std::to_string(pos) + " does not match the haps file";
            dist[s] = (int)strtol(f[i + 1].c_str(), nullptr, 10);
        }
    }

    // ---- props.bin (data.cpp:427-449): int snp, bp, dist; three 1024-byte strings ----
    {
        OutFile props(out + "/props.bin");
        std::vector<char> rec(12 + 3 * 1024);
        for (int s = 0; s < L; s++) {
            std::fill(rec.begin(), rec.end(), 0);
            const int head[3] = {s, t.bp[s], dist[s]};
            memcpy(rec.data(), head, 12);
            const std::string *str[3] = {&t.rsid[s], &t.anc[s], &t.alt[s]};
            for (int k = 0; k < 3; k++) memcpy(rec.data() + 12 + 1024 * k, str[k]->data(), std::min<size_t>(str[k]->size(), 1023));
            props.raw(rec.data(), rec.size());
        }
        std::string e = props.finish();
        if (!e.empty()) return e;
    }

    // ---- genetic positions and recombination distances (data.cpp:451-481) ----
    std::vector<double> rpos((size_t)L + 1), r((size_t)L);
    {
        GeneticMap gm;
        std::string e = gm.load(f_map);
        if (!e.empty()) return e;
        for (int s = 0; s <= L; s++) rpos[s] = gm.morgans(t.bp[s]);
        for (int s = 0; s < L; s++) r[s] = std::max(rpos[s + 1] - rpos[s], 1e-10) * 2500;
    }

    // ---- per-chunk position files (data.cpp:485-516) ----
    for (int c = 0; c < num_chunks; c++) {
        const std::string base = out + "/chunk_" + std::to_string(c);
        const int s0 = plan[c].file_begin;
        const unsigned Lc = (unsigned)(plan[c].end - s0);
        OutFile fbp(base + ".bp"), fdist(base + ".dist"), frpos(base + ".rpos"), fr(base + ".r");
        fbp.put(Lc).put(&t.bp[s0], Lc);
        fdist.put(Lc).put(&dist[s0], Lc);
        frpos.put(Lc + 1).put(&rpos[s0], (size_t)Lc + 1);
        fr.put(Lc).put(&r[s0], Lc);
        for (OutFile *f : {&fbp, &fdist, &frpos, &fr}) {
            std::string e = f->finish();
            if (!e.empty()) return e;
        }
    }

    if (info) {
        info->N = N;
        info->L = L;
        info->num_chunks = num_chunks;
        info->max_windows = tot.max_new_windows;
        info->actual_min_memory_gb = peak_gb;
        info->warnings = warn.str();
    }
    return "";
}

} // namespace rp
