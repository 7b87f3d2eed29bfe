// relate_cli.cpp — `relate`: drop-in for the reference's `Relate` binary on the Paint path.
//
// Mirrors the CLI surface of /root/reference/include/pipeline/Relate.cpp:15-328:
//   * the same flag set (Relate.cpp:19-45) is accepted; unknown flags are an error, as with cxxopts;
//   * `--mode Paint` (Relate.cpp:64-79 -> pipeline/Paint.cpp:17-108) runs natively on the GPUs through
//     the C ABI in include/relate_paint.h;
//   * `--mode MakeChunks` (Relate.cpp:60-63 -> pipeline/MakeChunks.cpp:14-117) runs natively (host code, rp_make_chunks);
//   * `--mode All` (Relate.cpp:190-296) keeps the reference's stage order, with MakeChunks and Paint native and
//     every other stage delegated to the reference binary;
//   * every other mode is handed to the reference binary unchanged (exec).
// The reference binary is found via $RELATE_REFERENCE_BIN, else "<dir of this exe>/Relate.ref".
// Extra flags of this build (stripped before delegating): --gpus a,b,c   --fp64   --chunks a-b   --resident   --gpu_topology
// `--resident` (modes BuildTopology and All) runs BuildTopology in `Relate_gpu` — the reference linked with this repo's
// binding of DistanceMeasure::GetMatrix (relate_b200/integration/; found via $RELATE_GPU_BIN, else next to this
// executable) — with the stepping stones kept in HBM: no paint files are written or read.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <sys/resource.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

#include "../../include/relate_paint.h"

namespace {

struct Flag {
    const char *name;
    char shortname; // 0 if none
    bool has_value;
    const char *help;
};

// Relate.cpp:19-45, in the reference's order
const Flag kFlags[] = {
    {"help", 0, false, "Print help."},
    {"mode", 0, true, "Choose which part of the algorithm to run."},
    {"haps", 0, true, "Filename of haps file (Output file format of Shapeit)."},
    {"sample", 0, true, "Filename of sample file (Output file format of Shapeit)."},
    {"map", 0, true, "Genetic map."},
    {"mutation_rate", 'm', true, "Mutation rate."},
    {"effectiveN", 'N', true, "Effective population size."},
    {"output", 'o', true, "Filename of output without file extension."},
    {"dist", 0, true, "Optional but recommended. Distance in BP between SNPs."},
    {"annot", 0, true, "Optional. Filename of file containing additional annotation of snps."},
    {"memory", 0, true, "Optional. Approximate memory allowance in GB for storing distance matrices. Default is 5GB."},
    {"sample_ages", 0, true, "Optional. Filename of file containing sample ages (one per line)."},
    {"chunk_index", 0, true, "Optional. Index of chunk."},
    {"first_section", 0, true, "Optional. Index of first section to infer."},
    {"last_section", 0, true, "Optional. Index of last section to infer."},
    {"coal", 0, true, "Optional. Filename of file containing coalescent rates."},
    {"fb", 0, true, "Optional. Force build a new tree every x bases."},
    {"no_consistency", 0, false, "Optional. Disable consistency option."},
    {"transversion", 0, false, "Only use transversion for bl estimation."},
    {"postprocess", 0, false, "(beta option) Postprocess topology."},
    {"randomise", 0, false, "(beta option) Randomise topology in post processing step."},
    {"input", 'i', true, "Filename of input."},
    {"painting", 0, true, "Optional. Copying and transition parameters in chromosome painting algorithm. Format: theta,rho. Default: 0.001,1."},
    {"seed", 0, true, "Optional. Seed for MCMC in branch lengths estimation."},
    // this build only
    {"gpus", 0, true, "(relate_b200) Comma-separated CUDA device indices for --mode Paint. Default: all visible."},
    {"fp64", 0, false, "(relate_b200) fp64 state in the painting kernel (verification mode)."},
    {"chunks", 0, true, "(relate_b200) --mode Paint: paint chunks a-b (instead of --chunk_index), whole chunks per GPU."},
    {"resident", 0, false, "(relate_b200) BuildTopology / All: distance matrices from stepping stones resident in HBM (needs Relate_gpu); no paint files."},
    {"gpu_topology", 0, false, "(relate_b200) BuildTopology / All: distance matrices and trees on the GPU from the paint files (needs Relate_gpu)."},
};

void print_help()
{
    std::cout << "Usage:\n  Relate [OPTION...]\n\n";
    for (const Flag &f : kFlags) {
        std::string lhs = "  ";
        if (f.shortname) lhs += std::string("-") + f.shortname + ", ";
        else lhs += "    ";
        lhs += std::string("--") + f.name;
        if (f.has_value) lhs += " arg";
        std::cout << std::left << std::setw(30) << lhs << f.help << "\n";
    }
    std::cout << std::endl;
}

struct Args {
    std::map<std::string, std::string> val;
    std::set<std::string> present;
    std::vector<std::string> passthrough; // argv for the reference binary (our own flags removed)
    bool count(const char *k) const { return present.count(k) != 0; }
    const std::string &get(const char *k) const { return val.at(k); }
};

const Flag *find_long(const std::string &n)
{
    for (const Flag &f : kFlags)
        if (n == f.name) return &f;
    return nullptr;
}
const Flag *find_short(char c)
{
    for (const Flag &f : kFlags)
        if (f.shortname && f.shortname == c) return &f;
    return nullptr;
}

bool parse(int argc, char **argv, Args &a, std::string &err)
{
    for (int i = 1; i < argc; i++) {
        std::string tok = argv[i];
        const Flag *f = nullptr;
        std::string value;
        bool inline_value = false;
        if (tok.rfind("--", 0) == 0) {
            std::string name = tok.substr(2);
            size_t eq = name.find('=');
            if (eq != std::string::npos) {
                value = name.substr(eq + 1);
                name = name.substr(0, eq);
                inline_value = true;
            }
            f = find_long(name);
            if (!f) { err = "Option '" + name + "' does not exist"; return false; }
        } else if (tok.size() == 2 && tok[0] == '-') {
            f = find_short(tok[1]);
            if (!f) { err = std::string("Option '") + tok[1] + "' does not exist"; return false; }
        } else {
            err = "unexpected argument '" + tok + "'";
            return false;
        }
        if (f->has_value && !inline_value) {
            if (i + 1 >= argc) { err = std::string("Option '") + f->name + "' is missing an argument"; return false; }
            value = argv[++i];
        }
        a.present.insert(f->name);
        a.val[f->name] = value;
        const bool ours = !strcmp(f->name, "gpus") || !strcmp(f->name, "fp64") || !strcmp(f->name, "chunks") || !strcmp(f->name, "resident") ||
                          !strcmp(f->name, "gpu_topology");
        if (!ours) {
            a.passthrough.push_back(std::string("--") + f->name);
            if (f->has_value) a.passthrough.push_back(value);
        }
    }
    return true;
}

std::string reference_binary(const char *argv0)
{
    if (const char *e = getenv("RELATE_REFERENCE_BIN")) return e;
    char buf[4096];
    ssize_t n = readlink("/proc/self/exe", buf, sizeof buf - 1);
    std::string self = n > 0 ? std::string(buf, (size_t)n) : std::string(argv0);
    size_t slash = self.rfind('/');
    std::string dir = slash == std::string::npos ? "." : self.substr(0, slash);
    return dir + "/Relate.ref";
}

// the reference linked with this repo's GetMatrix binding (relate_b200/integration/Makefile); "" if there is none
std::string gpu_consumer_binary(const char *argv0)
{
    std::string p;
    if (const char *e = getenv("RELATE_GPU_BIN")) p = e;
    else {
        const std::string ref = reference_binary(argv0);
        p = ref.substr(0, ref.rfind('/') + 1) + "Relate_gpu";
    }
    return access(p.c_str(), X_OK) == 0 ? p : std::string();
}

// run the reference binary (or Relate_gpu) with `args`; returns its exit status (or 127).  resident: RELATE_GPU_RESIDENT=1
int run_reference(const std::string &bin, const std::vector<std::string> &args, bool resident = false)
{
    if (access(bin.c_str(), X_OK) != 0) {
        std::cerr << "relate: this build implements --mode Paint only; set RELATE_REFERENCE_BIN (or place the reference binary at "
                  << bin << ") to run the other stages." << std::endl;
        return 127;
    }
    std::vector<char *> av;
    av.push_back(const_cast<char *>(bin.c_str()));
    for (const std::string &s : args) av.push_back(const_cast<char *>(s.c_str()));
    av.push_back(nullptr);
    pid_t pid = fork();
    if (pid == 0) {
        if (resident) setenv("RELATE_GPU_RESIDENT", "1", 1);
        execv(bin.c_str(), av.data());
        _exit(127);
    }
    int status = 0;
    waitpid(pid, &status, 0);
    if (WIFEXITED(status)) return WEXITSTATUS(status);
    return 128 + (WIFSIGNALED(status) ? WTERMSIG(status) : 0);
}

std::vector<std::string> with_mode(const Args &a, const std::string &mode, const std::vector<std::string> &extra = {})
{
    std::vector<std::string> out;
    const std::set<std::string> replaced = {"--mode", "--chunk_index", "--first_section", "--last_section"};
    for (size_t i = 0; i < a.passthrough.size(); i++) {
        const std::string &t = a.passthrough[i];
        if (replaced.count(t)) { i++; continue; }
        out.push_back(t);
    }
    out.push_back("--mode");
    out.push_back(mode);
    for (const std::string &e : extra) out.push_back(e);
    return out;
}

// pipeline/Paint.cpp:17-108 behind the C ABI
int paint(const Args &a, int chunk_index, int last_chunk = -1, std::ostream &log = std::cerr)
{
    std::vector<int> devs;
    if (a.count("gpus")) {
        std::stringstream ss(a.get("gpus"));
        std::string t;
        while (std::getline(ss, t, ','))
            if (!t.empty()) devs.push_back(atoi(t.c_str()));
        for (size_t i = 0; i < devs.size(); i++)
            for (size_t j = 0; j < i; j++)
                if (devs[i] == devs[j]) {
                    log << "relate: --gpus lists device " << devs[i] << " twice." << std::endl;
                    return 1;
                }
    } else {
        int n = rp_device_count();
        for (int i = 0; i < n; i++) devs.push_back(i);
    }
    if (devs.empty()) {
        log << "relate: no CUDA device visible; --mode Paint has no CPU path in this build." << std::endl;
        return 1;
    }
    log << "---------------------------------------------------------" << std::endl;
    log << "Painting sequences..." << std::endl;
    rp_stats st;
    memset(&st, 0, sizeof st);
    const char *painting = a.count("painting") ? a.get("painting").c_str() : nullptr;
    int rc = last_chunk < 0 ? rp_paint_chunk(a.get("output").c_str(), chunk_index, painting, devs.data(), (int)devs.size(),
                                             a.count("fp64") ? RP_FP64 : 0u, &st)
                            : rp_paint_chunks(a.get("output").c_str(), chunk_index, last_chunk, painting, devs.data(),
                                              (int)devs.size(), a.count("fp64") ? RP_FP64 : 0u, &st);
    if (rc != RP_OK) {
        log << "relate: Paint failed: " << rp_last_error() << std::endl;
        return 1;
    }
    rusage usage;
    getrusage(RUSAGE_SELF, &usage);
    log << "GPU Paint: " << devs.size() << " device(s), " << st.n_targets << " targets, kernel " << std::fixed
              << std::setprecision(3) << st.ms_paint << " ms, prep " << st.ms_prep << " ms, record encoder " << st.ms_rle << " ms, file writes " << st.ms_write
              << " ms, total " << st.ms_total << " ms." << std::endl;
    log << "CPU Time spent: " << usage.ru_utime.tv_sec << "." << std::setfill('0') << std::setw(6) << usage.ru_utime.tv_usec
              << "s; Max Memory usage: " << std::setprecision(6) << usage.ru_maxrss / 1000.0 << "Mb." << std::endl;
    log << "---------------------------------------------------------" << std::endl << std::endl;
    return 0;
}

// pipeline/MakeChunks.cpp:14-117 behind the C ABI
int make_chunks(const Args &a, bool paint_follows = false)
{
    bool help = false;
    if (!a.count("haps") || !a.count("sample") || !a.count("map") || !a.count("output")) {
        std::cout << "Not enough arguments supplied." << std::endl;
        std::cout << "Needed: haps, sample, map, output. Optional: memory, dist, transversion." << std::endl;
        help = true;
    }
    if (a.count("help") || help) {
        print_help();
        std::cout << "Use to make smaller chunks from the data." << std::endl;
        return 0;
    }
    std::cerr << "---------------------------------------------------------" << std::endl;
    std::cerr << "Parsing data.." << std::endl;
    char warnings[4096];
    warnings[0] = 0;
    const float mem = a.count("memory") ? strtof(a.get("memory").c_str(), nullptr) : 5.0f;
    // under --mode All this build's Paint is certain to follow: leave it the bit-packed genotype rows (chunk_<c>.hapbits)
    const char *hb = getenv("RELATE_HAPBITS");
    const unsigned mcf = (paint_follows || (hb && atoi(hb) != 0)) ? RP_MC_HAPBITS : 0u;
    const int rc = rp_make_chunks_ex(a.get("haps").c_str(), a.get("sample").c_str(), a.get("map").c_str(),
                                     a.count("dist") ? a.get("dist").c_str() : nullptr, a.get("output").c_str(),
                                     a.count("transversion") ? 1 : 0, mem, mcf, nullptr, warnings, sizeof warnings);
    if (rc != RP_OK) {
        std::cerr << rp_last_error() << std::endl;
        return 1;
    }
    std::cerr << warnings;
    rusage usage;
    getrusage(RUSAGE_SELF, &usage);
    std::cerr << "CPU Time spent: " << usage.ru_utime.tv_sec << "." << std::setfill('0') << std::setw(6) << usage.ru_utime.tv_usec
              << "s; Max Memory usage: " << usage.ru_maxrss / 1000.0 << "Mb." << std::endl;
    std::cerr << "---------------------------------------------------------" << std::endl << std::endl;
    return 0;
}

bool read_ints(const std::string &path, int *dst, int n)
{
    FILE *fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    bool ok = fread(dst, 4, n, fp) == (size_t)n;
    fclose(fp);
    return ok;
}

} // namespace

int main(int argc, char **argv)
{
    Args a;
    std::string err;
    if (!parse(argc, argv, a, err)) {
        std::cerr << "relate: " << err << std::endl;
        return 1;
    }
    const std::string mode = a.count("mode") ? a.get("mode") : "";
    if (a.count("output")) { // Relate.cpp:50-58
        if (a.get("output").find('/') != std::string::npos) {
            std::cerr << "Output needs to be in working directory." << std::endl;
            return 1;
        }
    }
    const std::string ref = reference_binary(argv[0]);

    if (mode == "MakeChunks") return make_chunks(a);

    if (mode == "Paint") {
        bool help = false;
        if (a.count("chunks") && a.count("output") && !a.count("help")) {
            int c0 = 0, c1 = -1;
            if (sscanf(a.get("chunks").c_str(), "%d-%d", &c0, &c1) != 2 || c0 < 0 || c1 < c0) {
                std::cerr << "relate: --chunks expects a-b" << std::endl;
                return 1;
            }
            return paint(a, c0, c1);
        }
        if (!a.count("chunk_index") || !a.count("output")) { // Relate.cpp:66-76
            std::cout << "Not enough arguments supplied." << std::endl;
            std::cout << "Needed: chunk_index, output." << std::endl;
            help = true;
        }
        if (a.count("help") || help) {
            print_help();
            std::cout << "Use after MakeChunks to paint all haps against all." << std::endl;
            return 0;
        }
        return paint(a, atoi(a.get("chunk_index").c_str()));
    }

    const bool resident = a.count("resident") != 0;
    // --gpu_topology: BuildTopology in Relate_gpu as well (window repaint, distance matrices and trees on the GPU), but from the
    // paint files a Paint stage wrote
    const bool gpu_topology = a.count("gpu_topology") != 0;
    std::string gpu_bin;
    if (resident || gpu_topology) {
        gpu_bin = gpu_consumer_binary(argv[0]);
        if (gpu_bin.empty()) {
            std::cerr << "relate: --resident / --gpu_topology need Relate_gpu (make -C relate_b200/integration REF=<reference checkout>; or set RELATE_GPU_BIN)." << std::endl;
            return 1;
        }
    }
    if (mode == "BuildTopology" && resident) { // Relate.cpp:81-115 with the distance matrices from the GPU
        if (a.count("output") && !a.count("help")) {
            const std::string cdir = a.get("output") + "/chunk_" + (a.count("chunk_index") ? a.get("chunk_index") : std::string("0"));
            mkdir(cdir.c_str(), 0700); // (the directories Paint would have created: BuildTopology writes its .anc/.mut there,
            mkdir((cdir + "/paint").c_str(), 0700); //  Finalize / Clean remove both and exit(1) if one is missing)
        }
        return run_reference(gpu_bin, a.passthrough, true);
    }
    if (mode == "BuildTopology" && gpu_topology) return run_reference(gpu_bin, a.passthrough, false);

    if (mode == "All") { // Relate.cpp:190-296 with Paint native
        if (a.count("help") || !a.count("output") ||
            (!a.count("chunk_index") && (!a.count("haps") || !a.count("sample") || !a.count("map"))))
            return run_reference(ref, a.passthrough); // let the reference print its own usage text
        const std::string out = a.get("output");
        int start_chunk = 0, end_chunk = 0;
        if (a.count("chunk_index")) {
            start_chunk = end_chunk = atoi(a.get("chunk_index").c_str());
        } else {
            int rc = make_chunks(a, /*paint_follows=*/!resident); // (with --resident no Paint stage runs that would consume the sidecar)
            if (rc != 0) return rc;
            int hdr[3];
            if (!read_ints(out + "/parameters.bin", hdr, 3)) {
                std::cerr << "relate: cannot read " << out << "/parameters.bin" << std::endl;
                return 1;
            }
            end_chunk = hdr[2] - 1;
        }
        // Paint runs one chunk ahead of the reference's CPU stages: while BuildTopology ...
        // CombineSections work on chunk c, chunk c+1 is painted in the background; its banner is printed when its turn
        // comes.  InferBranchLengths deletes a chunk's paint files (InferBranchLengths.cpp:60-75), so at most two
        // chunks' paint files are on disk at any time, as with the cluster scripts' bounded number of paintings.
        struct Ahead { std::future<int> rc; std::ostringstream log; };
        std::unique_ptr<Ahead> ahead;
        auto start_paint = [&](int c) {
            std::unique_ptr<Ahead> p(new Ahead());
            Ahead *raw = p.get();
            p->rc = std::async(std::launch::async, [&a, c, raw]() { return paint(a, c, -1, raw->log); });
            return p;
        };
        for (int c = start_chunk; c <= end_chunk; c++) {
            std::cerr << "---------------------------------------------------------" << std::endl;
            std::cerr << "Starting chunk " << c << " of " << end_chunk << "." << std::endl;
            std::cerr << "---------------------------------------------------------" << std::endl << std::endl;
            int hdr[3];
            if (!read_ints(out + "/parameters_c" + std::to_string(c) + ".bin", hdr, 3)) {
                std::cerr << "relate: cannot read parameters_c" << c << ".bin" << std::endl;
                return 1;
            }
            const int num_sections = hdr[2] - 1;
            const std::string cs = std::to_string(c), ls = std::to_string(num_sections - 1);
            int rc = 0;
            if (resident) {
                // no Paint stage and no paint files: BuildTopology (Relate_gpu) paints the chunk inside its own process
                // and opens every window from the stepping stones it keeps in HBM
                mkdir((out + "/chunk_" + cs).c_str(), 0700);
                mkdir((out + "/chunk_" + cs + "/paint").c_str(), 0700); // stays empty; Finalize insists on removing it
                if ((rc = run_reference(gpu_bin, with_mode(a, "BuildTopology", {"--chunk_index", cs, "--first_section", "0", "--last_section", ls}), true))) return rc;
            } else {
                if (!ahead) ahead = start_paint(c);
                rc = ahead->rc.get();
                std::cerr << ahead->log.str();
                ahead.reset();
                if (rc) return rc;
                if (c < end_chunk) ahead = start_paint(c + 1);
                // (an error return below waits for the background painter: a std::async future joins in its destructor)
                if ((rc = run_reference(gpu_topology ? gpu_bin : ref, with_mode(a, "BuildTopology", {"--chunk_index", cs, "--first_section", "0", "--last_section", ls})))) return rc;
            }
            if ((rc = run_reference(ref, with_mode(a, "FindEquivalentBranches", {"--chunk_index", cs})))) return rc;
            if (a.count("postprocess")) {
                if ((rc = run_reference(ref, with_mode(a, "PostProcess", {"--chunk_index", cs})))) return rc;
            }
            if ((rc = run_reference(ref, with_mode(a, "InferBranchLengths", {"--chunk_index", cs, "--first_section", "0", "--last_section", ls})))) return rc;
            if ((rc = run_reference(ref, with_mode(a, "CombineSections", {"--chunk_index", cs})))) return rc;
        }
        if (!a.count("chunk_index")) {
            int rc = run_reference(ref, with_mode(a, "Finalize"));
            if (rc) return rc;
        }
        std::cerr << "---------------------------------------------------------" << std::endl;
        std::cerr << "Done." << std::endl;
        std::cerr << "---------------------------------------------------------" << std::endl;
        return 0;
    }

    // everything else belongs to the reference binary
    return run_reference(ref, a.passthrough);
}
