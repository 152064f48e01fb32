// main.cpp -- `raxtax` command-line binary on top of the B200 library.
//
// Mirrors the reference's CLI surface and output files (src/io.rs:112-154, src/main.rs:14-173):
//   raxtax -d <db.fasta[.gz]> -i <queries.fasta[.gz]> [-o PREFIX] [--skip-exact-matches] [--raw-confidence] [--tsv]
//          [--redo] [--skip-db] [-c] [-t N] [--pin] [-v|-q] [--gpu N]
// writes <PREFIX>/raxtax.out, raxtax.log, raxtax.ckp and (with --tsv) raxtax.tsv in the reference's formats
// (lineage.rs:17-48).  Not carried over (DESIGN.md "out of scope"): the bincode .bin database (--only-db is refused,
// --skip-db / --clean are accepted and have nothing to do), checkpoint resume (raxtax.json), thread pinning.
// Exit codes follow main.rs: 73 CANTCREAT, 66 NOINPUT, 74 IOERR, 75 TEMPFAIL, 0 OK.

#include <sys/stat.h>
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "raxtax_host.h"

namespace {

enum { EX_OK_ = 0, EX_USAGE_ = 64, EX_NOINPUT_ = 66, EX_CANTCREAT_ = 73, EX_IOERR_ = 74, EX_TEMPFAIL_ = 75 };

struct Args {
    std::string database_path, query_file, prefix = "raxtax";
    bool skip_exact_matches = false, tsv = false, only_db = false, skip_db = false, clean = false, raw_confidence = false, redo = false,
         pin = false;
    int threads = 0, verbosity = 3 /* Info */, gpu = 0;
    size_t batch = 0;
};

void usage() {
    fprintf(stderr,
            "Usage: raxtax [OPTIONS] --database-path <DATABASE_PATH>\n\n"
            "Options:\n"
            "  -d, --database-path <DATABASE_PATH>  Path to the database fasta file (.gz/.gzip accepted)\n"
            "  -i, --query-file <QUERY_FILE>        Path to the query file\n"
            "      --skip-exact-matches             If used for mislabling analysis, you want to skip exact sequence matches\n"
            "      --tsv                            Output primary result file in tsv format\n"
            "      --only-db                        (not supported by the B200 build: no binary database)\n"
            "      --skip-db                        Don't create the binary database (always the case here)\n"
            "  -c, --clean                          Remove checkpoint files after a successful run\n"
            "      --raw-confidence                 Don't adjust confidence values for 1 exact match\n"
            "  -t, --threads <THREADS>              Accepted for compatibility (the hot path runs on the GPU)\n"
            "  -o, --prefix <PREFIX>                Output prefix [default: raxtax]\n"
            "      --redo                           Force override of existing output files\n"
            "      --pin                            Accepted for compatibility\n"
            "      --gpu <ORDINAL>                  CUDA device to use [default: 0]\n"
            "      --batch <N>                      Queries per device batch [default: all]\n"
            "  -v / -q                              More / less output\n");
}

bool read_file(const std::string& path, std::string* out) {  // utils::get_reader (utils.rs:42-60): gz by extension; gzopen also reads plain files
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    int n;
    while ((n = gzread(f, buf, sizeof buf)) > 0) out->append(buf, (size_t)n);
    gzclose(f);
    return n == 0;
}

struct Writers {
    FILE *primary = nullptr, *tsv = nullptr, *log = nullptr, *progress = nullptr;
    int verbosity = 3;
};

int send_cb(void* user, const char* label, const char* primary, const char* tsv) {  // writer thread body (main.rs:128-134)
    Writers* w = (Writers*)user;
    if (w->tsv && tsv) fprintf(w->tsv, "%s\n", tsv);
    if (fprintf(w->primary, "%s\n", primary) < 0) return 1;
    fprintf(w->progress, "%s\n", label);
    return 0;
}

void log_cb(void* user, int level, const char* msg) {  // env_logger without timestamps/targets (main.rs:34-39)
    Writers* w = (Writers*)user;
    if (level > w->verbosity) return;
    fprintf(w->log, "[%s] %s\n", level == 2 ? "WARN " : "INFO ", msg);
}

bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

}  // namespace

int main(int argc, char** argv) {
    Args a;
    for (int i = 1; i < argc; ++i) {
        std::string s = argv[i];
        auto val = [&](const char* name) -> const char* {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: a value is required for '%s'\n", name);
                exit(EX_USAGE_);
            }
            return argv[++i];
        };
        if (s == "-d" || s == "--database-path") a.database_path = val("--database-path");
        else if (s == "-i" || s == "--query-file") a.query_file = val("--query-file");
        else if (s == "--skip-exact-matches") a.skip_exact_matches = true;
        else if (s == "--tsv") a.tsv = true;
        else if (s == "--only-db") a.only_db = true;
        else if (s == "--skip-db") a.skip_db = true;
        else if (s == "-c" || s == "--clean") a.clean = true;
        else if (s == "--raw-confidence") a.raw_confidence = true;
        else if (s == "-t" || s == "--threads") a.threads = atoi(val("--threads"));
        else if (s == "-o" || s == "--prefix") a.prefix = val("--prefix");
        else if (s == "--redo") a.redo = true;
        else if (s == "--pin") a.pin = true;
        else if (s == "--gpu") a.gpu = atoi(val("--gpu"));
        else if (s == "--batch") a.batch = (size_t)atoll(val("--batch"));
        else if (s == "-v") a.verbosity = 4;
        else if (s == "-q") a.verbosity = 2;
        else if (s == "-h" || s == "--help") {
            usage();
            return 0;
        } else {
            fprintf(stderr, "error: unexpected argument '%s'\n", s.c_str());
            usage();
            return EX_USAGE_;
        }
    }
    if (a.database_path.empty() || (a.query_file.empty() && !a.only_db)) {
        usage();
        return EX_USAGE_;
    }
    if (a.only_db && a.skip_db) {
        fprintf(stderr, "error: the argument '--only-db' cannot be used with '--skip-db'\n");
        return EX_USAGE_;
    }
    if (a.only_db) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m --only-db: the B200 build keeps no binary database\n");
        return EX_CANTCREAT_;
    }
    // output folder (io.rs:202-263)
    const std::string out_path = a.prefix + "/raxtax.out";
    if (exists(a.prefix) && !a.redo) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m Output folder %s already exists! Please specify another folder with -o <PATH> or run with --redo to force overriding existing files!\n",
                a.prefix.c_str());
        return EX_CANTCREAT_;
    }
    if (!exists(a.prefix) && mkdir(a.prefix.c_str(), 0777) != 0) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m cannot create %s\n", a.prefix.c_str());
        return EX_CANTCREAT_;
    }
    Writers w;
    w.verbosity = a.verbosity;
    w.primary = fopen(out_path.c_str(), "w");
    w.log = fopen((a.prefix + "/raxtax.log").c_str(), "w");
    w.progress = fopen((a.prefix + "/raxtax.ckp").c_str(), "w");
    if (a.tsv) w.tsv = fopen((a.prefix + "/raxtax.tsv").c_str(), "w");
    if (!w.primary || !w.log || !w.progress || (a.tsv && !w.tsv)) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m cannot create output files under %s\n", a.prefix.c_str());
        return EX_CANTCREAT_;
    }
    {  // io::write_build_info (io.rs:92-110)
        std::string cmd;
        for (int i = 0; i < argc; ++i) cmd += std::string(i ? " " : "") + argv[i];
        fprintf(w.log, "raxtax-b200 1.5.0 (B200 sm_100a build of the raxtax hot path)\nBuild flags: \nCommand: %s\n"
                       "------------------------------------------------------------\n", cmd.c_str());
        fflush(w.log);
    }
    auto t_total = std::chrono::steady_clock::now();

    std::string db_text, q_text;
    if (!read_file(a.database_path, &db_text)) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m Failed to parse %s: cannot read file\n", a.database_path.c_str());
        return EX_NOINPUT_;
    }
    rxh_tree* tree = rxh_tree_from_fasta(db_text.data(), db_text.size());
    if (!tree) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m Failed to parse %s: %s\n", a.database_path.c_str(), rxh_last_error());
        return EX_NOINPUT_;
    }
    std::string().swap(db_text);
    if (!read_file(a.query_file, &q_text)) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m Failed to parse %s: cannot read file\n", a.query_file.c_str());
        return EX_NOINPUT_;
    }
    rxh_queries* queries = rxh_queries_from_fasta(q_text.data(), q_text.size());
    if (!queries) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m Failed to parse %s: %s\n", a.query_file.c_str(), rxh_last_error());
        return EX_NOINPUT_;
    }
    std::string().swap(q_text);

    rtx_ctx* ctx = nullptr;
    if (rtx_ctx_create(a.gpu, &ctx) != 0) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m %s\n", rtx_last_error(nullptr));
        return EX_TEMPFAIL_;
    }
    if (rxh_tree_upload(tree, ctx, 0, 0) != 0) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m index upload: %s\n", rxh_last_error());
        return EX_TEMPFAIL_;
    }
    int warnings = 0;
    auto t0 = std::chrono::steady_clock::now();
    int rc = rxh_raxtax(ctx, queries, tree, a.skip_exact_matches, a.raw_confidence, a.batch, send_cb, &w, a.tsv, log_cb, &w, &warnings);
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc != 0) {
        fprintf(stderr, "\x1b[31m[ERROR]\x1b[0m Error while sending results to IO-thread!: %s\n", rxh_last_error());
        return EX_TEMPFAIL_;
    }
    if (warnings && a.verbosity >= 2)
        fprintf(stderr, "\x1b[33m[WARN ]\x1b[0m Exact matches for some queries differ above the species level! Check the log file for more information!\n");
    if (a.verbosity >= 3) {
        fprintf(w.log, "[INFO ] raxtax(), Elapsed=%.6fs\n", secs);
        fprintf(w.log, "[INFO ] Total Runtime, Elapsed=%.6fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_total).count());
    }
    for (FILE* f : {w.primary, w.tsv, w.log, w.progress})
        if (f) fclose(f);
    if (a.clean) remove((a.prefix + "/raxtax.ckp").c_str());  // Checkpoint::cleanup: nothing else was written
    rxh_queries_free(queries);
    rxh_tree_free(tree);
    rtx_ctx_destroy(ctx);
    return EX_OK_;
}
