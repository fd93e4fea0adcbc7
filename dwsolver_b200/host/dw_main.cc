// dw_main.cc — `dwsolver -g <guidefile> [options]`, the reference's command line
// (/root/reference/src/dw_main.c:58-90 usage text, :213-259 guidefile format) on top of this build's
// dw_solve().  Guidefile: n, n subproblem LP files, the master LP file, a monolithic file name (read
// and ignored), an optional objective constant.  File names are relative to the current directory.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/dwsolver.h"

static void usage(const char *argv0)
{
    std::printf("Usage: %s -g <guide_file_name> [options]\n\n", argv0);
    std::printf(" -v, --version           Print version information and exit.\n");
    std::printf(" -g, --guidefile <file>  REQUIRED. Use <file> as the guidefile.\n");
    std::printf(" -h, -help, --help       Print this usage and return.\n");
    std::printf(" -i, --integerize        (not supported by the GPU build; ignored)\n");
    std::printf(" -e, --sub-int-enforce   (not supported by the GPU build; ignored)\n");
    std::printf(" -r, --round             (not supported by the GPU build; ignored)\n");
    std::printf(" -p, --perturb           Perturb the linking right-hand sides against degeneracy during the solve.\n");
    std::printf(" --verbose | --output-all | -n, --output-normal | -q, --quiet | --silent\n");
    std::printf(" --skip-monolithic-read  Guide file has no monolithic entry.\n");
    std::printf(" --write-final-master | --no-write-final-master   Write the final master to done.cpxlp (default: no).\n");
    std::printf(" --print-timing | --no-timing\n");
    std::printf(" --phase1_max <n>        Max phase I iterations (default 100).\n");
    std::printf(" --phase2_max <n>        Max phase II iterations (default 3000).\n");
    std::printf(" --write-container <f>   With -g: store the parsed instance in the binary container <f> and exit.\n");
    std::printf(" --container <f>         Solve the instance stored in container <f> (no guidefile, no LP text).\n");
    std::printf(" --devices <list|all>    Shard the subproblems over these CUDA devices, e.g. 0,1,2,3 (default: device 0).\n\n");
    std::printf("The guide file format:\n  n\n  sub_file_1\n  ...\n  sub_file_n\n  master_file\n  monolithic_file\n  objective_constant\n\n");
}

static std::string chomp(std::string s)
{
    while (!s.empty() && (s.back() == '\n' || s.back() == '\r' || s.back() == ' ' || s.back() == '\t')) s.pop_back();
    size_t b = 0;
    while (b < s.size() && (s[b] == ' ' || s[b] == '\t')) ++b;
    return s.substr(b);
}

int main(int argc, char **argv)
{
    dw_options_t opt;
    dw_options_init(&opt);
    std::string guide, container_in, container_out;
    bool skip_mono = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-v" || a == "--version") { std::printf("%s\n", dw_version()); return 0; }
        else if (a == "-h" || a == "-help" || a == "--help") { usage(argv[0]); return 0; }
        else if ((a == "-g" || a == "--guidefile") && i + 1 < argc) guide = argv[++i];
        else if (a == "-i" || a == "--integerize") opt.integerize_flag = 1;
        else if (a == "-e" || a == "--sub-int-enforce") opt.enforce_sub_integrality = 1;
        else if (a == "-r" || a == "--round") opt.rounding_flag = 1;
        else if (a == "-p" || a == "--perturb") opt.perturb = 1;
        else if ((a == "--mip-gap" || a == "--mip_gap") && i + 1 < argc) opt.mip_gap = std::atof(argv[++i]);
        else if (a == "--verbose") opt.verbosity = DW_OUTPUT_VERBOSE;
        else if (a == "--output-all") opt.verbosity = DW_OUTPUT_ALL;
        else if (a == "-n" || a == "--output-normal") opt.verbosity = DW_OUTPUT_NORMAL;
        else if (a == "-q" || a == "--quiet") opt.verbosity = DW_OUTPUT_QUIET;
        else if (a == "--silent") opt.verbosity = DW_OUTPUT_SILENT;
        else if (a == "--skip-monolithic-read") skip_mono = true;
        else if (a == "--write-bases" || a == "--write-int-probs")        // src/dw_main.c:171-174
            std::fprintf(stderr, "Warning: %s is not supported by this build; option ignored.\n", a.c_str());
        else if (a == "--write-final-master") opt.print_final_master = 1;
        else if (a == "--no-write-final-master") opt.print_final_master = 0;
        else if (a == "--print-timing") opt.print_timing_data = 1;
        else if (a == "--no-timing") opt.print_timing_data = 0;
        else if (a == "--phase1_max" && i + 1 < argc) opt.max_phase1_iterations = std::atoi(argv[++i]);
        else if (a == "--phase2_max" && i + 1 < argc) opt.max_phase2_iterations = std::atoi(argv[++i]);
        else if (a == "--write-container" && i + 1 < argc) container_out = argv[++i];
        else if (a == "--container" && i + 1 < argc) container_in = argv[++i];
        else if (a == "--devices" && i + 1 < argc) setenv("DWSOLVER_DEVICES", argv[++i], 1);
        else std::printf("Command line argument %s not recognized. Trying to continue.\n", argv[i]);
    }
    if (!container_in.empty()) {                                  // binary container: nothing else to read
        dw_result_t res;
        const int rc = dw_solve_container(container_in.c_str(), &opt, &res);
        dw_result_free(&res);
        return rc == DW_STATUS_OK ? 0 : rc;
    }
    if (guide.empty()) { usage(argv[0]); return 1; }
    std::ifstream gf(guide);
    if (!gf) { std::fprintf(stderr, "Unable to open guide file %s\n", guide.c_str()); return 1; }
    std::string line;
    if (!std::getline(gf, line)) { std::fprintf(stderr, "Guide file is empty.\n"); return 1; }
    const int n = std::atoi(line.c_str());
    if (n < 1 || n > 26000) { std::fprintf(stderr, "Invalid subproblem count: %d\n", n); return 1; }
    std::vector<std::string> subs(n);
    for (int i = 0; i < n; ++i) {
        if (!std::getline(gf, line)) { std::fprintf(stderr, "Guide file truncated reading subproblem %d\n", i); return 1; }
        subs[i] = chomp(line);
    }
    std::string master;
    if (!std::getline(gf, master)) { std::fprintf(stderr, "Guide file truncated reading master file name\n"); return 1; }
    master = chomp(master);
    if (!skip_mono) std::getline(gf, line);                       // monolithic file name: read and ignored
    if (std::getline(gf, line)) {
        const double s = std::strtod(line.c_str(), nullptr);
        if (s != 0.0) opt.shift = s;
    }
    std::vector<const char *> ptrs(n);
    for (int i = 0; i < n; ++i) ptrs[i] = subs[i].c_str();
    if (!container_out.empty()) {
        const int rc = dw_write_container(master.c_str(), ptrs.data(), n, opt.shift, container_out.c_str());
        if (rc == DW_STATUS_OK && opt.verbosity >= DW_OUTPUT_NORMAL) std::printf("Wrote %s\n", container_out.c_str());
        return rc == DW_STATUS_OK ? 0 : rc;
    }
    dw_result_t res;
    const int rc = dw_solve(master.c_str(), ptrs.data(), n, &opt, &res);
    dw_result_free(&res);
    return rc == DW_STATUS_OK ? 0 : rc;
}
