// granges_cli.cpp -- `granges` command-line mirror for the hot path: map / filter / windows.
//
// Same sub-commands and flags as the reference's clap definitions (src/main/mod.rs:96-117 filter,
// :159-184 map, :194-214 windows), same TSV output (src/granges.rs:225-283, src/data/mod.rs:49-65,
// src/io/tsv.rs:5-12), same row order (seqnames in genome-file order, then input order --
// src/iterators.rs:43-76), errors as `Error: <message>` on stderr with exit status 1
// (src/main/mod.rs:347-355).  All range arithmetic runs on the GPU through the C ABI
// (include/granges_b200.h); this file only parses and prints.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "format.hpp"
#include "granges_b200.h"

namespace {

[[noreturn]] void die(const std::string &msg) {
    fprintf(stderr, "Error: %s\n", msg.c_str());
    exit(1);
}

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    explicit Mapped(const std::string &path) {
        int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) die(path + ": " + strerror(errno));
        struct stat st;
        if (fstat(fd, &st) != 0) die(path + ": " + strerror(errno));
        n = (size_t)st.st_size;
        if (n) {
            void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) die(path + ": mmap failed");
            p = (const char *)m;
            if (n >= 2 && (unsigned char)p[0] == 0x1f && (unsigned char)p[1] == 0x8b)
                die(path + ": gzip-compressed input is not supported by this mirror (decompress it first)");
        }
        close(fd);
    }
};

struct Genome {
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    std::unordered_map<std::string_view, int> id;  // views into `names` (stable: filled once)
};

// read_seqlens -- src/io/file.rs:20-41
Genome read_genome(const std::string &path) {
    Mapped f(path);
    Genome g;
    const char *p = f.p, *end = f.p + f.n;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', end - p);
        const char *le = nl ? nl : end;
        if (le > p) {
            const char *tab = (const char *)memchr(p, '\t', le - p);
            if (!tab) die("invalid genome file: expected <seqname>\\t<length>");
            std::string name(p, tab);
            char *e;
            unsigned long long len = strtoull(tab + 1, &e, 10);
            for (auto &x : g.names)
                if (x == name) die("Invalid genome file: sequence '" + name + "' is duplicated");
            g.names.push_back(name);
            g.lens.push_back((uint32_t)len);
        }
        p = le + 1;
    }
    g.names.shrink_to_fit();
    for (size_t i = 0; i < g.names.size(); i++) g.id.emplace(std::string_view(g.names[i]), (int)i);
    return g;
}

struct SeqRanges {
    std::vector<uint32_t> starts, ends;
    std::vector<double> score;       // BED5 only
    std::vector<uint8_t> valid;      // BED5 only: 0 = "." (None)
    std::vector<std::string_view> rest;  // bedlike: the columns after the third, verbatim
};

enum class Cols { Bed3, Bed5, Bedlike };

// BED parsing as the reference's typed iterators do it (src/io/parsers/tsv.rs:27-121): tab-separated,
// '#' lines skipped; BED5 = name + score with "." -> None (src/io/parsers/bed/mod.rs:37-73).
size_t parse_bed(const Mapped &f, const Genome &g, Cols cols, bool skip_missing, std::vector<SeqRanges> &out,
                 bool *any_rest = nullptr) {
    out.assign(g.names.size(), SeqRanges());
    const char *p = f.p, *end = f.p + f.n;
    size_t rows = 0, line_no = 0;
    int last_id = -1;
    std::string_view last_name;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', end - p);
        const char *le = nl ? nl : end;
        const char *next = le + 1;
        if (le > p && le[-1] == '\r') le--;
        line_no++;
        if (le == p || *p == '#') {
            p = next;
            continue;
        }
        const char *t1 = (const char *)memchr(p, '\t', le - p);
        if (!t1) die("line " + std::to_string(line_no) + ": expected at least 3 tab-separated columns");
        const char *t2 = (const char *)memchr(t1 + 1, '\t', le - t1 - 1);
        if (!t2) die("line " + std::to_string(line_no) + ": expected at least 3 tab-separated columns");
        const char *t3 = (const char *)memchr(t2 + 1, '\t', le - t2 - 1);
        const char *e3 = t3 ? t3 : le;
        std::string_view name(p, t1 - p);
        int id;
        if (last_id >= 0 && name == last_name) {
            id = last_id;
        } else {
            auto it = g.id.find(name);
            if (it == g.id.end()) {
                if (skip_missing) {
                    p = next;
                    continue;
                }
                die("The sequence name '" + std::string(name) +
                    "' is not found within the provided ranges container. Check the sequence names for typos or missing entries.");
            }
            id = it->second;
            last_id = id;
            last_name = name;
        }
        uint64_t s = 0, e = 0;
        for (const char *q = t1 + 1; q < t2; q++) {
            if (*q < '0' || *q > '9') die("line " + std::to_string(line_no) + ": invalid start coordinate");
            s = s * 10 + (uint64_t)(*q - '0');
        }
        for (const char *q = t2 + 1; q < e3; q++) {
            if (*q < '0' || *q > '9') die("line " + std::to_string(line_no) + ": invalid end coordinate");
            e = e * 10 + (uint64_t)(*q - '0');
        }
        if (e <= s || e > 0x7fffffffull)
            die("line " + std::to_string(line_no) + ": invalid range (the reference panics: end must be > start and fit i32)");
        SeqRanges &sr = out[id];
        sr.starts.push_back((uint32_t)s);
        sr.ends.push_back((uint32_t)e);
        if (cols == Cols::Bed5) {
            if (!t3) die("line " + std::to_string(line_no) + ": BED5 needs a name and a score column");
            const char *t4 = (const char *)memchr(t3 + 1, '\t', le - t3 - 1);
            if (!t4) die("line " + std::to_string(line_no) + ": BED5 needs a name and a score column");
            const char *t5 = (const char *)memchr(t4 + 1, '\t', le - t4 - 1);
            const char *e5 = t5 ? t5 : le;
            if (e5 - (t4 + 1) == 1 && t4[1] == '.') {
                sr.score.push_back(0.0);
                sr.valid.push_back(0);
            } else {
                std::string tmp(t4 + 1, e5);
                char *ep;
                double v = strtod(tmp.c_str(), &ep);
                if (ep == tmp.c_str() || *ep) die("line " + std::to_string(line_no) + ": parsing error: invalid float literal");
                sr.score.push_back(v);
                sr.valid.push_back(1);
            }
        } else if (cols == Cols::Bedlike) {
            sr.rest.push_back(t3 ? std::string_view(t3 + 1, le - t3 - 1) : std::string_view());
            if (t3 && any_rest) *any_rest = true;
        }
        rows++;
        p = next;
    }
    return rows;
}

struct Out {
    FILE *f;
    std::vector<char> buf;
    size_t n = 0;
    explicit Out(const char *path) : buf(1 << 22) {
        f = path ? fopen(path, "wb") : stdout;
        if (!f) die(std::string(path) + ": " + strerror(errno));
    }
    char *room(size_t need) {
        if (n + need > buf.size()) flush();
        return buf.data() + n;
    }
    void flush() {
        if (n) fwrite(buf.data(), 1, n, f);
        n = 0;
    }
    ~Out() {
        flush();
        if (f != stdout) fclose(f);
    }
};

void check(grb_ctx *ctx, int rc) {
    if (rc != GRB_OK) die(ctx ? grb_last_error(ctx) : "cannot create a CUDA context (granges_b200 has no CPU fallback)");
}

int parse_op(const std::string &s) {
    // clap ValueEnum names of FloatOperation (src/data/operations.rs:35-51), kebab-case
    if (s == "sum") return GRB_OP_SUM;
    if (s == "sum-not-empty") return GRB_OP_SUM_NOT_EMPTY;
    if (s == "min") return GRB_OP_MIN;
    if (s == "max") return GRB_OP_MAX;
    if (s == "mean") return GRB_OP_MEAN;
    if (s == "median") return GRB_OP_MEDIAN;
    if (s == "collapse") die("collapse is not offered: its output depends on the reference's unspecified visit order");
    die("invalid value '" + s + "' for '--func <FUNC>'");
}

struct Args {
    std::string genome, left, right, output;
    bool has_output = false, skip_missing = false, chop = false;
    std::vector<int> ops;
    uint32_t width = 0, step = 0;
    bool has_width = false;
};

int cmd_map(const Args &a) {
    if (a.ops.empty()) die("No operation was specified. See granges map --help.");
    Genome g = read_genome(a.genome);
    Mapped lf(a.left), rf(a.right);
    std::vector<SeqRanges> L, R;
    size_t nl = parse_bed(lf, g, Cols::Bed3, a.skip_missing, L);
    size_t nr = parse_bed(rf, g, Cols::Bed5, a.skip_missing, R);
    if (nl == 0 || nr == 0) die("The input file had no rows.");
    grb_ctx *ctx = nullptr;
    check(nullptr, grb_ctx_create(0, &ctx));
    const int nseq = (int)g.names.size();
    std::vector<grb_index *> idx(nseq, nullptr);
    std::vector<uint64_t> off(nseq + 1, 0);
    std::vector<uint32_t> ls, le;
    ls.reserve(nl);
    le.reserve(nl);
    for (int s = 0; s < nseq; s++) {
        off[s] = ls.size();
        ls.insert(ls.end(), L[s].starts.begin(), L[s].starts.end());
        le.insert(le.end(), L[s].ends.begin(), L[s].ends.end());
        if (!R[s].starts.empty()) {
            check(ctx, grb_index_build(ctx, R[s].starts.data(), R[s].ends.data(), nullptr, R[s].starts.size(), &idx[s]));
            check(ctx, grb_index_set_scores(idx[s], R[s].score.data(), R[s].valid.data(), R[s].score.size()));
        }
    }
    off[nseq] = ls.size();
    const int k = (int)a.ops.size();
    std::vector<double> out(nl * (size_t)k);
    std::vector<uint8_t> ov(nl * (size_t)k);
    check(ctx, grb_map_joins_f64_batch(ctx, idx.data(), off.data(), nseq, ls.data(), le.data(), a.ops.data(), k, out.data(), ov.data()));
    Out w(a.has_output ? a.output.c_str() : nullptr);
    for (int s = 0; s < nseq; s++) {
        const std::string &name = g.names[s];
        for (uint64_t i = off[s]; i < off[s + 1]; i++) {
            char *p = w.room(name.size() + 64 + (size_t)k * 420);
            char *q = p;
            memcpy(q, name.data(), name.size());
            q += name.size();
            *q++ = '\t';
            q = grh::format_u32(q, ls[i]);
            *q++ = '\t';
            q = grh::format_u32(q, le[i]);
            for (int c = 0; c < k; c++) {
                *q++ = '\t';
                if (ov[i * k + c])
                    q = grh::format_f64(q, out[i * k + c]);
                else
                    *q++ = '.';
            }
            *q++ = '\n';
            w.n += q - p;
        }
    }
    for (auto ix : idx) grb_index_destroy(ix);
    grb_ctx_destroy(ctx);
    return 0;
}

int cmd_filter(const Args &a) {
    Genome g = read_genome(a.genome);
    Mapped lf(a.left), rf(a.right);
    std::vector<SeqRanges> L, R;
    size_t nl = parse_bed(lf, g, Cols::Bedlike, a.skip_missing, L);
    parse_bed(rf, g, Cols::Bed3, a.skip_missing, R);
    grb_ctx *ctx = nullptr;
    check(nullptr, grb_ctx_create(0, &ctx));
    const int nseq = (int)g.names.size();
    std::vector<grb_index *> idx(nseq, nullptr);
    std::vector<uint64_t> off(nseq + 1, 0);
    std::vector<uint32_t> ls, le;
    ls.reserve(nl);
    le.reserve(nl);
    for (int s = 0; s < nseq; s++) {
        off[s] = ls.size();
        ls.insert(ls.end(), L[s].starts.begin(), L[s].starts.end());
        le.insert(le.end(), L[s].ends.begin(), L[s].ends.end());
        // a seqname the right file never mentions has no container: the semi-join drops its lefts
        // (src/granges.rs:1433-1437) -- idx stays NULL
        if (!R[s].starts.empty()) check(ctx, grb_index_build(ctx, R[s].starts.data(), R[s].ends.data(), nullptr, R[s].starts.size(), &idx[s]));
    }
    off[nseq] = ls.size();
    std::vector<uint8_t> keep(nl);
    uint64_t nk = 0;
    check(ctx, grb_filter_overlaps_batch(ctx, idx.data(), off.data(), nseq, ls.data(), le.data(), 0, keep.data(), &nk));
    Out w(a.has_output ? a.output.c_str() : nullptr);
    for (int s = 0; s < nseq; s++) {
        const std::string &name = g.names[s];
        for (uint64_t i = off[s]; i < off[s + 1]; i++) {
            if (!keep[i]) continue;
            std::string_view rest = L[s].rest[i - off[s]];
            char *p = w.room(name.size() + 64 + rest.size());
            char *q = p;
            memcpy(q, name.data(), name.size());
            q += name.size();
            *q++ = '\t';
            q = grh::format_u32(q, ls[i]);
            *q++ = '\t';
            q = grh::format_u32(q, le[i]);
            if (!rest.empty()) {
                *q++ = '\t';
                memcpy(q, rest.data(), rest.size());
                q += rest.size();
            }
            *q++ = '\n';
            w.n += q - p;
        }
    }
    for (auto ix : idx) grb_index_destroy(ix);
    grb_ctx_destroy(ctx);
    return 0;
}

int cmd_windows(const Args &a) {
    if (!a.has_width) die("the following required arguments were not provided: --width <WIDTH>");
    Genome g = read_genome(a.genome);
    grb_ctx *ctx = nullptr;
    check(nullptr, grb_ctx_create(0, &ctx));
    grb_ranges *r = nullptr;
    check(ctx, grb_windows(ctx, g.lens.data(), (int)g.lens.size(), a.width, a.step, a.chop ? 1 : 0, &r));
    const size_t n = grb_ranges_len(r);
    std::vector<uint32_t> seq(n), st(n), en(n);
    check(ctx, grb_ranges_copy(r, seq.data(), st.data(), en.data()));
    Out w(a.has_output ? a.output.c_str() : nullptr);
    for (size_t i = 0; i < n; i++) {
        const std::string &name = g.names[seq[i]];
        char *p = w.room(name.size() + 40);
        char *q = p;
        memcpy(q, name.data(), name.size());
        q += name.size();
        *q++ = '\t';
        q = grh::format_u32(q, st[i]);
        *q++ = '\t';
        q = grh::format_u32(q, en[i]);
        *q++ = '\n';
        w.n += q - p;
    }
    grb_ranges_destroy(r);
    grb_ctx_destroy(ctx);
    return 0;
}

void usage() {
    fprintf(stderr,
            "granges (B200 mirror of the overlap-join commands)\n"
            "  granges map     -g <genome> -l <left.bed> -r <right.bed5> -f <op[,op..]> [-o <out>] [-s]\n"
            "  granges filter  -g <genome> -l <left.bed> -r <right.bed> [-o <out>] [-s]\n"
            "  granges windows -g <genome> -w <width> [-s <step>] [-c] [-o <out>]\n"
            "ops: sum, sum-not-empty, min, max, mean, median\n");
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) {
        usage();
        return 2;
    }
    const std::string cmd = argv[1];
    const bool windows = cmd == "windows";
    Args a;
    for (int i = 2; i < argc; i++) {
        std::string f = argv[i], val;
        bool has_val = false;
        size_t eq = f.find('=');
        if (f.rfind("--", 0) == 0 && eq != std::string::npos) {
            val = f.substr(eq + 1);
            f = f.substr(0, eq);
            has_val = true;
        }
        auto need = [&]() -> std::string {
            if (has_val) return val;
            if (i + 1 >= argc) die("a value is required for '" + f + "' but none was supplied");
            return argv[++i];
        };
        if (f == "-g" || f == "--genome")
            a.genome = need();
        else if (f == "-l" || f == "--left")
            a.left = need();
        else if (f == "-r" || f == "--right")
            a.right = need();
        else if (f == "-o" || f == "--output") {
            a.output = need();
            a.has_output = true;
        } else if (f == "-f" || f == "--func") {
            std::string v = need();
            size_t b = 0;
            while (b <= v.size()) {
                size_t c = v.find(',', b);
                if (c == std::string::npos) c = v.size();
                if (c > b) a.ops.push_back(parse_op(v.substr(b, c - b)));
                b = c + 1;
            }
        } else if (f == "-w" || f == "--width") {
            a.width = (uint32_t)strtoul(need().c_str(), nullptr, 10);
            a.has_width = true;
        } else if (windows && (f == "-s" || f == "--step"))
            a.step = (uint32_t)strtoul(need().c_str(), nullptr, 10);
        else if (f == "-c" || f == "--chop")
            a.chop = true;
        else if (!windows && (f == "-s" || f == "--skip-missing"))
            a.skip_missing = true;
        else if (f == "-h" || f == "--help") {
            usage();
            return 0;
        } else
            die("unexpected argument '" + f + "' found");
    }
    if (a.genome.empty()) die("the following required arguments were not provided: --genome <GENOME>");
    if (cmd == "map") {
        if (a.left.empty() || a.right.empty()) die("the following required arguments were not provided: --left <LEFT> --right <RIGHT>");
        return cmd_map(a);
    }
    if (cmd == "filter") {
        if (a.left.empty() || a.right.empty()) die("the following required arguments were not provided: --left <LEFT> --right <RIGHT>");
        return cmd_filter(a);
    }
    if (windows) return cmd_windows(a);
    die("unrecognized subcommand '" + cmd + "' (this mirror implements map, filter and windows)");
}
