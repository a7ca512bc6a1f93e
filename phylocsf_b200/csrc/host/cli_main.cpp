// cli_main.cpp — `PhyloCSF parameter_set [file1 file2 ...] [options]`: the reference's command line
// (src/PhyloCSF.ml:17-49,469-491) over the B200 compute library. Same options, same output lines.
#include <unistd.h>

#include <atomic>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "driver.hpp"

using namespace pcsf::host;

static const char* kUsage =
    "usage: PhyloCSF parameter_set [file1 file2 ...]\n"
    "input will be read from stdin if no filenames are given.\n\n"
    "options:\n"
    "  --strategy=mle|fixed|omega   evaluation strategy (default mle)\n"
    "  --debug                      print extra information about parameters and errors\n"
    " input interpretation:\n"
    "  --files                      input list(s) of alignment filenames instead of individual alignment(s)\n"
    "  --removeRefGaps              automatically remove any alignment columns that are gapped in the reference sequence\n"
    "  --species=Species1,Species2,...  hint that only this subset of species will be used in any of the alignments\n"
    " searching multiple reading frames and ORFs:\n"
    "  -f 1|3|6, --frames=1|3|6     how many reading frames to search (default 1)\n"
    "  --orf=AsIs|ATGStop|StopStop|StopStop3|ToFirstStop|FromLastStop|ToOrFromStop  search for ORFs (default AsIs)\n"
    "  --minCodons=INT              minimum ORF length for searching over ORFs (default 25 codons)\n"
    "  --allScores                  report scores of all regions evaluated, not just the max\n"
    "  -p INT                       host threads that read and prepare alignments (default: all cores); scoring is batched on the GPU\n"
    " output control:\n"
    "  --bls                        include alignment branch length score (BLS) for the reported region in output\n"
    "  --ancComp                    include ancestral sequence composition score in output\n"
    "  --dna                        include DNA sequence in output\n"
    "  --aa                         include amino acid translation in output\n";

static std::string lower(std::string s) {
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}
[[noreturn]] static void usage_exit(const std::string& msg) {
    if (!msg.empty()) std::cerr << "PhyloCSF: " << msg << "\n";
    std::cerr << kUsage;
    std::exit(255);  // exit (-1)
}

static std::vector<std::string> read_lines(std::istream& in) {
    std::vector<std::string> lines;
    std::string ln;
    while (std::getline(in, ln)) lines.push_back(ln);
    return lines;
}

int main(int argc, char** argv) {
    Options opt;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto value = [&](const std::string& name) -> std::string {  // --name=value | --name value
            const size_t eq = a.find('=');
            if (eq != std::string::npos) return a.substr(eq + 1);
            if (i + 1 >= argc) usage_exit("option " + name + " requires an argument");
            return argv[++i];
        };
        const std::string key = a.substr(0, a.find('='));
        if (a.size() < 2 || a[0] != '-') pos.push_back(a);
        else if (key == "--strategy") {
            const std::string v = lower(value(key));
            if (v == "mle") opt.strategy = STRAT_MLE;
            else if (v == "fixed") opt.strategy = STRAT_FIXED;
            else if (v == "omega") opt.strategy = STRAT_OMEGA;
            else if (v == "nop") opt.strategy = STRAT_NOP;
            else usage_exit("invalid strategy " + v);
        } else if (key == "--files") opt.filenames = true;
        else if (key == "--removeRefGaps") opt.remove_ref_gaps = true;
        else if (key == "--allowRefGaps") opt.allow_ref_gaps = true;
        else if (key == "--species") opt.species = value(key);
        else if (key == "--frames" || key == "-f" || (a.size() > 2 && a[0] == '-' && a[1] == 'f' && a[2] != '-' && key[1] != '-')) {
            std::string v = (key == "--frames" || a == "-f") ? value(key) : a.substr(2);
            if (v == "1") opt.frames = 1; else if (v == "3") opt.frames = 3; else if (v == "6") opt.frames = 6;
            else usage_exit("invalid reading frame " + v);
        } else if (key == "--orf") {
            const std::string v = lower(value(key));
            if (v == "asis") opt.orf = AsIs; else if (v == "atgstop") opt.orf = ATGStop; else if (v == "stopstop") opt.orf = StopStop;
            else if (v == "stopstop3") opt.orf = StopStop3; else if (v == "tofirststop") opt.orf = ToFirstStop;
            else if (v == "fromlaststop") opt.orf = FromLastStop; else if (v == "toorfromstop") opt.orf = ToOrFromStop;
            else usage_exit("invalid ORF search mode " + v);
        } else if (key == "--minCodons") opt.min_codons = std::atoi(value(key).c_str());
        else if (key == "--allScores") opt.all_scores = true;
        else if (key == "-p" || (a.size() > 2 && a[0] == '-' && a[1] == 'p' && key[1] != '-')) {
            opt.procs = std::atoi((a == "-p" ? value(key) : a.substr(2)).c_str());
        } else if (key == "--bls") opt.bls = true;
        else if (key == "--ancComp") opt.anc_comp = true;
        else if (key == "--dna") opt.dna = true;
        else if (key == "--aa") opt.aa = true;
        else if (key == "--debug") opt.debug = true;
        else if (key == "--help" || key == "-h") { std::cout << kUsage; return 0; }
        else usage_exit("no such option: " + a);
    }
    if (pos.empty()) usage_exit("");
    const std::string paramset = pos[0];
    std::vector<std::string> fns_input(pos.begin() + 1, pos.end());
    if (opt.orf != AsIs && opt.allow_ref_gaps) {
        std::cerr << "--allowRefGaps should not be used with --orf\n";
        usage_exit("");
    }
    if (opt.orf != AsIs && opt.frames == 1)
        std::cerr << "Warning: --orf with --frames=1; are you sure you don't want to search for ORFs in three or six frames?\n";
    if (const char* d = std::getenv("PCSF_DEVICE")) opt.device = std::atoi(d);
    if (const char* b = std::getenv("PCSF_BATCH_COLS")) opt.batch_cols = std::atoll(b);

    try {
        // parameter file resolution, src/PhyloCSF.ml:407-421
        std::string prefix = paramset;
        if (paramset.find('/') == std::string::npos) {
            const char* base = std::getenv("PHYLOCSF_BASE");
            if (!base || access(base, R_OK) != 0)
                throw failure("PHYLOCSF_BASE environment variable must be set to the root directory of the source or executable distribution");
            prefix = std::string(base) + "/PhyloCSF_Parameters/" + paramset;
        }
        if (opt.strategy == STRAT_OMEGA) {  // optional <paramset>_omega key<TAB>value file, :453-463
            std::ifstream cfg(prefix + "_omega");
            if (cfg) {
                std::map<std::string, std::string> kv;
                std::string ln;
                while (std::getline(cfg, ln)) {
                    ln = ocaml_trim(ln);
                    if (ln.empty() || ln[0] == '#') continue;
                    const size_t tab = ln.find('\t');
                    kv[tab == std::string::npos ? ln : ln.substr(0, tab)] = tab == std::string::npos ? "" : ln.substr(tab + 1);
                }
                try {
                    opt.omega_H1 = float_of_string(kv.at("omega_H1"));
                    opt.sigma_H1 = float_of_string(kv.at("sigma_H1"));
                } catch (...) {
                    throw failure("configuration file " + prefix + "_omega exists but does not specify valid omega_H1 and sigma_H1 values");
                }
            }
        }
        std::vector<int> devices;  // PCSF_DEVICES=0,1,.. or "all": batches go round-robin over these GPUs
        if (const char* ds = std::getenv("PCSF_DEVICES")) {
            if (std::string(ds) == "all") for (int d = 0; d < pcsf_device_count(); d++) devices.push_back(d);
            else {
                std::stringstream ss(ds);
                std::string tok;
                while (std::getline(ss, tok, ',')) if (!tok.empty()) devices.push_back(std::atoi(tok.c_str()));
            }
        }
        Driver drv(opt, prefix, devices);

        std::vector<std::string> fns;
        bool from_stdin = false;
        if (opt.filenames) {
            if (fns_input.empty()) fns = read_lines(std::cin);
            else
                for (auto& f : fns_input) {
                    std::ifstream in(f);
                    if (!in) throw HostError("Sys_error(\"" + f + ": No such file or directory\")");
                    for (auto& l : read_lines(in)) fns.push_back(l);
                }
        } else if (fns_input.empty()) {
            from_stdin = true;
            fns.push_back("");
        } else fns = fns_input;

        // Files are read and prepared (parse, checks, regions, leaf codes) by a pool of host threads, a
        // block at a time, and appended to the GPU batch in input order. The reference's -p N forked
        // per-region workers; here N only caps the host threads (default: all cores).
        struct Slot {
            Driver::Prepared prep;
            bool missing = false;
        };
        unsigned nthreads = std::max(1u, std::thread::hardware_concurrency());
        if (opt.procs > 1) nthreads = (unsigned)opt.procs;
        if (const char* t = std::getenv("PCSF_HOST_THREADS")) nthreads = std::max(1, std::atoi(t));
        const size_t block = 64 * (size_t)nthreads;
        for (size_t b0 = 0; b0 < fns.size(); b0 += block) {
            const size_t b1 = std::min(fns.size(), b0 + block);
            std::vector<Slot> slots(b1 - b0);
            std::atomic<size_t> next{b0};
            auto work = [&]() {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= b1) return;
                    const std::string& fn = fns[i];
                    const std::string name = fn.empty() ? "(STDIN)" : fn;
                    std::vector<std::string> lines;
                    if (fn.empty() && from_stdin) lines = read_lines(std::cin);
                    else {
                        std::ifstream in(fn);
                        if (!in) {
                            slots[i - b0].missing = true;
                            slots[i - b0].prep.job.name = name;
                            continue;
                        }
                        lines = read_lines(in);
                    }
                    slots[i - b0].prep = drv.prepare(name, lines);
                }
            };
            const unsigned nt = (unsigned)std::min<size_t>(nthreads, b1 - b0);
            if (nt <= 1) work();
            else {
                std::vector<std::thread> pool;
                for (unsigned t = 0; t < nt; t++) pool.emplace_back(work);
                for (auto& t : pool) t.join();
            }
            for (size_t i = b0; i < b1; i++) {
                Slot& sl = slots[i - b0];
                if (sl.missing) {
                    drv.finish(std::cout);
                    std::cout << sl.prep.job.name << "\tabort\tSys_error(\"" << fns[i] << ": No such file or directory\")\n";
                    std::cout.flush();
                    return 255;
                }
                if (!drv.append(std::move(sl.prep), std::cout)) return 255;
            }
        }
        drv.finish(std::cout);
    } catch (const std::exception& e) {
        std::cerr << "Fatal error: exception " << e.what() << "\n";
        return 2;
    }
    return 0;
}
