// cli_main.cpp — `PhyloCSF parameter_set [file1 file2 ...] [options]`: the reference's command line
// (src/PhyloCSF.ml:17-49,469-491) over the B200 compute library. Same options, same output lines.
#include <fcntl.h>
#include <malloc.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <mutex>

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

// Whole file into `buf` (kept at its high-water size by the caller, so nothing is allocated or cleared per
// file); `len` = bytes read. False if the file cannot be opened.
static bool read_file(const std::string& fn, std::vector<char>& buf, size_t& len) {
    const int fd = open(fn.c_str(), O_RDONLY);
    if (fd < 0) return false;
    if (buf.size() < (1u << 16)) buf.resize(1u << 16);
    len = 0;
    for (;;) {
        if (len == buf.size()) buf.resize(buf.size() * 2);
        const ssize_t r = read(fd, buf.data() + len, buf.size() - len);
        if (r <= 0) break;
        len += (size_t)r;
    }
    close(fd);
    return true;
}

static std::vector<std::string> read_lines(std::istream& in) {
    std::vector<std::string> lines;
    std::string ln;
    while (std::getline(in, ln)) lines.push_back(ln);
    return lines;
}

int main(int argc, char** argv) {
    // The reader threads allocate and the appender frees tens of KB per alignment: keep freed memory in the
    // process instead of trimming and re-faulting it (both serialise the threads on the address-space lock).
    const double t_main = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
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
        // Two scoring contexts per GPU by default: while one batch is being scored, the next one's nucleotides
        // are already copied and framed on the other context's stream (fixed, 58mammals: 62 k -> 82 k alignments/s).
        if (devices.empty()) devices.push_back(opt.device);
        int per_device = 2;
        if (const char* e = std::getenv("PCSF_CONTEXTS_PER_DEVICE")) per_device = std::max(1, std::atoi(e));
        {
            const std::vector<int> once = devices;
            for (int k = 1; k < per_device; k++) devices.insert(devices.end(), once.begin(), once.end());
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

        // Files are read and prepared (parse, checks, regions, nucleotide rows or leaf codes) by a pool of host
        // threads and appended to the GPU batch in input order by this thread: an ordered pipeline over chunks
        // of files, with a bounded window so that the readers run ahead of the appender but not away from it.
        // The reference's -p N forked per-region workers; here N only caps the host threads (default: all cores).
        struct Slot {
            Driver::Prepared prep;
            bool missing = false;
        };
        struct Chunk {
            std::vector<Slot> slots;
            std::shared_ptr<std::vector<uint8_t>> nt;  // nucleotide rows of the chunk's alignments (fast form): handed to the batch, not copied
        };
        unsigned nthreads = std::max(1u, std::thread::hardware_concurrency());
        if (opt.procs > 1) nthreads = (unsigned)opt.procs;
        if (const char* t = std::getenv("PCSF_HOST_THREADS")) nthreads = std::max(1, std::atoi(t));
        const size_t n_files = fns.size();
        const size_t chunk = 32, n_chunks = (n_files + chunk - 1) / chunk;
        const size_t window = std::max<size_t>(4, 4 * (size_t)nthreads);  // chunks in flight
        std::vector<Chunk> ring(window);
        BufferPool nt_pool;
        std::vector<char> ready(window, 0);
        std::mutex mu;
        std::condition_variable cv_ready, cv_space;
        size_t next_chunk = 0, consumed = 0;  // under mu
        bool stop = false;
        auto prepare_one = [&](size_t i, Slot& sl, std::vector<uint8_t>& nt_buf) {
            const std::string& fn = fns[i];
            const std::string name = fn.empty() ? "(STDIN)" : fn;
            static thread_local std::vector<char> text;  // reused: no allocation per file
            size_t len = 0;
            if (fn.empty() && from_stdin) {
                text.assign(std::istreambuf_iterator<char>(std::cin), std::istreambuf_iterator<char>());
                len = text.size();
            } else if (!read_file(fn, text, len)) {
                sl.missing = true;
                sl.prep.job.name = name;
                return;
            }
            sl.prep = drv.prepare_text(name, text.data(), len, nt_buf);
        };
        auto work = [&]() {
            Chunk ch;
            size_t nt_hint = 0;  // bytes the previous chunks needed: reserve once instead of growing
            for (;;) {
                size_t c;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_space.wait(lk, [&] { return stop || next_chunk >= n_chunks || next_chunk < consumed + window; });
                    if (stop || next_chunk >= n_chunks) return;
                    c = next_chunk++;
                }
                ch.slots.clear();
                ch.slots.resize(std::min(chunk, n_files - c * chunk));
                ch.nt = nt_pool.get();
                if (ch.nt->capacity() < nt_hint) ch.nt->reserve(nt_hint);
                for (size_t k = 0; k < ch.slots.size(); k++) prepare_one(c * chunk + k, ch.slots[k], *ch.nt);
                nt_hint = std::max(nt_hint, ch.nt->size() + ch.nt->size() / 8);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    ring[c % window] = std::move(ch);
                    ready[c % window] = 1;
                }
                ch = Chunk();
                cv_ready.notify_all();
            }
        };
        std::vector<std::thread> pool;
        const unsigned nt = (unsigned)std::min<size_t>(nthreads, std::max<size_t>(1, n_chunks));
        for (unsigned t = 0; t < nt; t++) pool.emplace_back(work);
        auto shut = [&]() {
            {
                std::lock_guard<std::mutex> lk(mu);
                stop = true;
            }
            cv_space.notify_all();
            for (auto& t : pool)
                if (t.joinable()) t.join();
        };
        struct PoolGuard {  // an exception on the way (a scoring context that failed, a CUDA error) must not leave joinable threads behind
            std::function<void()> f;
            ~PoolGuard() { f(); }
        } pool_guard{shut};
        const bool host_profile = std::getenv("PCSF_HOST_PROFILE") != nullptr;  // where the appender's time goes, on stderr
        const double t_ready = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        double t_wait = 0, t_append = 0;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        for (size_t c = 0; c < n_chunks; c++) {
            Chunk ch;
            const double t0 = now();
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_ready.wait(lk, [&] { return ready[c % window] != 0; });
                ch = std::move(ring[c % window]);
                ring[c % window] = Chunk();
                ready[c % window] = 0;
                consumed = c + 1;
            }
            std::vector<Slot>& slots = ch.slots;
            cv_space.notify_all();
            const double t1 = now();
            t_wait += t1 - t0;
            for (size_t k = 0; k < slots.size(); k++) {
                Slot& sl = slots[k];
                if (sl.missing) {
                    shut();
                    drv.finish(std::cout);
                    std::cout << sl.prep.job.name << "\tabort\tSys_error(\"" << fns[c * chunk + k] << ": No such file or directory\")\n";
                    std::cout.flush();
                    return 255;
                }
                if (!drv.append(std::move(sl.prep), std::cout, ch.nt)) {
                    shut();
                    return 255;
                }
            }
            t_append += now() - t1;
        }
        shut();
        if (host_profile) {
            const double t2 = now();
            drv.finish(std::cout);
            std::cerr << "host profile: start-up (parameters, scoring contexts, file list) " << t_ready - t_main << " s, waited for readers "
                      << t_wait << " s, appended (incl. batch hand-off) " << t_append << " s, final drain " << now() - t2 << " s, " << nt
                      << " reader threads; fast reader took " << drv.n_fast << " alignments, general reader " << drv.n_general << "\n";
        }
        drv.finish(std::cout);
        // Everything is printed. Leave without tearing down gigabytes of staging buffers, device allocations and the
        // CUDA context piece by piece (0.5-2 s on the test box): the operating system and the driver reclaim them.
        std::cout.flush();
        std::cerr.flush();
        fflush(nullptr);
        if (!std::getenv("PCSF_FULL_TEARDOWN")) _exit(0);
    } catch (const std::exception& e) {
        std::cerr << "Fatal error: exception " << e.what() << "\n";
        return 2;
    }
    return 0;
}
