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

#include <time.h>
#include <zlib.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>
#include <map>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "format.hpp"
#include "granges_b200.h"

namespace {

// GRANGES_TIMING=1 prints the wall time of every phase on stderr
struct Phase {
    const char *name;
    timespec t0;
    explicit Phase(const char *n) : name(n) { clock_gettime(CLOCK_MONOTONIC, &t0); }
    ~Phase() {
        static const bool on = getenv("GRANGES_TIMING") != nullptr;
        if (!on) return;
        timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[timing] %-28s %8.1f ms\n", name, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
    }
};

[[noreturn]] void die(const std::string &msg) {
    fprintf(stderr, "Error: %s\n", msg.c_str());
    exit(1);
}

// gzip input is detected by its magic bytes, like the reference's reader (src/io/parsers/tsv.rs:27-45), and inflated
// into memory (all members of a multi-member / bgzip file).
std::vector<char> *inflate_all(const unsigned char *src, size_t n, const std::string &path);

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    std::vector<char> *inflated = nullptr;
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
            if (n >= 2 && (unsigned char)p[0] == 0x1f && (unsigned char)p[1] == 0x8b) {
                inflated = inflate_all((const unsigned char *)p, n, path);
                munmap(m, n);
                p = inflated->data();
                n = inflated->size();
            }
        }
        close(fd);
    }
    Mapped(const Mapped &) = delete;
    Mapped &operator=(const Mapped &) = delete;
    ~Mapped() { delete inflated; }  // (a plain mapping lives until the process exits)
};

std::vector<char> *inflate_all(const unsigned char *src, size_t n, const std::string &path) {
    auto *out = new std::vector<char>();
    out->reserve(n * 4);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) die(path + ": cannot initialise zlib");
    std::vector<unsigned char> buf(1 << 20);
    size_t pos = 0;
    int rc = Z_OK;
    while (pos < n) {
        const size_t chunk = std::min(n - pos, (size_t)1 << 30);
        zs.next_in = const_cast<unsigned char *>(src + pos);
        zs.avail_in = (uInt)chunk;
        while (zs.avail_in > 0) {
            zs.next_out = buf.data();
            zs.avail_out = (uInt)buf.size();
            rc = inflate(&zs, Z_NO_FLUSH);
            if (rc != Z_OK && rc != Z_STREAM_END) {
                inflateEnd(&zs);
                die(path + ": corrupt gzip stream");
            }
            out->insert(out->end(), (const char *)buf.data(), (const char *)buf.data() + (buf.size() - zs.avail_out));
            if (rc == Z_STREAM_END) {
                if (zs.avail_in == 0) break;
                if (inflateReset(&zs) != Z_OK) die(path + ": corrupt gzip stream");  // the next member
            }
        }
        pos += chunk - zs.avail_in;
        if (rc == Z_STREAM_END && pos < n && inflateReset(&zs) != Z_OK) die(path + ": corrupt gzip stream");  // a member ended on the chunk boundary
    }
    inflateEnd(&zs);
    if (rc != Z_STREAM_END) die(path + ": truncated gzip stream");
    return out;
}

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

int n_threads() {
    const char *e = getenv("GRANGES_THREADS");
    int t = (e && *e) ? atoi(e) : (int)std::thread::hardware_concurrency();
    return t < 1 ? 1 : (t > 64 ? 64 : t);
}

// run f(0..n-1) on up to n_threads() host threads
void parallel_for(int n, const std::function<void(int)> &f) {
    const int t = std::min(n, n_threads());
    if (t <= 1) {
        for (int i = 0; i < n; i++) f(i);
        return;
    }
    std::vector<std::thread> th;
    for (int w = 0; w < t; w++)
        th.emplace_back([&, w] {
            for (int i = w; i < n; i += t) f(i);
        });
    for (auto &x : th) x.join();
}

// BED parsing as the reference's typed iterators do it (src/io/parsers/tsv.rs:27-121): tab-separated,
// '#' lines skipped; BED5 = name + score with "." -> None (src/io/parsers/bed/mod.rs:37-73).
// The text is cut into chunks at line boundaries, the chunks are parsed on all host threads and
// stitched back in file order, so the per-seqname insertion order is the file's.
struct ParseError {
    bool set = false;
    std::string msg;
};

size_t parse_chunk(const char *p, const char *end, const Genome &g, Cols cols, bool skip_missing, std::vector<SeqRanges> &out, bool *any_rest,
                   ParseError &err) {
    out.assign(g.names.size(), SeqRanges());
    size_t rows = 0;
    int last_id = -1;
    std::string_view last_name;
    auto fail = [&](const char *line, const std::string &m) {
        if (!err.set) {
            err.set = true;
            err.msg = m + " (at byte offset of line start " + std::to_string((size_t)(line - p)) + " of its chunk)";
        }
    };
    while (p < end) {
        const char *line = p;
        const char *nl = (const char *)memchr(p, '\n', end - p);
        const char *le = nl ? nl : end;
        const char *next = le + 1;
        if (le > p && le[-1] == '\r') le--;
        p = next;
        if (le == line || *line == '#') continue;
        const char *t1 = (const char *)memchr(line, '\t', le - line);
        const char *t2 = t1 ? (const char *)memchr(t1 + 1, '\t', le - t1 - 1) : nullptr;
        if (!t1 || !t2) {
            fail(line, "expected at least 3 tab-separated columns");
            return rows;
        }
        const char *t3 = (const char *)memchr(t2 + 1, '\t', le - t2 - 1);
        const char *e3 = t3 ? t3 : le;
        std::string_view name(line, t1 - line);
        int id;
        if (last_id >= 0 && name == last_name) {
            id = last_id;
        } else {
            auto it = g.id.find(name);
            if (it == g.id.end()) {
                if (skip_missing) continue;
                fail(line, "The sequence name '" + std::string(name) +
                               "' is not found within the provided ranges container. Check the sequence names for typos or missing entries.");
                return rows;
            }
            id = it->second;
            last_id = id;
            last_name = name;
        }
        uint64_t s = 0, e = 0;
        bool bad = t2 == t1 + 1 || e3 == t2 + 1;
        for (const char *q = t1 + 1; q < t2; q++) {
            bad |= *q < '0' || *q > '9';
            s = s * 10 + (uint64_t)(*q - '0');
        }
        for (const char *q = t2 + 1; q < e3; q++) {
            bad |= *q < '0' || *q > '9';
            e = e * 10 + (uint64_t)(*q - '0');
        }
        if (bad) {
            fail(line, "invalid start / end coordinate");
            return rows;
        }
        if (e <= s || e > 0x7fffffffull) {
            fail(line, "invalid range (the reference panics: end must be > start and fit i32)");
            return rows;
        }
        SeqRanges &sr = out[id];
        sr.starts.push_back((uint32_t)s);
        sr.ends.push_back((uint32_t)e);
        if (cols == Cols::Bed5) {
            const char *t4 = t3 ? (const char *)memchr(t3 + 1, '\t', le - t3 - 1) : nullptr;
            if (!t4) {
                fail(line, "BED5 needs a name and a score column");
                return rows;
            }
            const char *t5 = (const char *)memchr(t4 + 1, '\t', le - t4 - 1);
            const char *e5 = t5 ? t5 : le;
            if (e5 - (t4 + 1) == 1 && t4[1] == '.') {
                sr.score.push_back(0.0);
                sr.valid.push_back(0);
            } else {
                char tmp[64];
                size_t len = (size_t)(e5 - (t4 + 1));
                double v = 0.0;
                bool okf = len > 0 && len < sizeof(tmp);
                if (okf) {
                    memcpy(tmp, t4 + 1, len);
                    tmp[len] = 0;
                    char *ep;
                    v = strtod(tmp, &ep);
                    okf = ep != tmp && *ep == 0;
                }
                if (!okf) {
                    fail(line, "parsing error: invalid float literal");
                    return rows;
                }
                sr.score.push_back(v);
                sr.valid.push_back(1);
            }
        } else if (cols == Cols::Bedlike) {
            sr.rest.push_back(t3 ? std::string_view(t3 + 1, le - t3 - 1) : std::string_view());
            if (t3 && any_rest) *any_rest = true;
        }
        rows++;
    }
    return rows;
}

size_t parse_bed(const Mapped &f, const Genome &g, Cols cols, bool skip_missing, std::vector<SeqRanges> &out,
                 bool *any_rest = nullptr) {
    const int nchunks = f.n < (1u << 20) ? 1 : n_threads() * 4;
    std::vector<const char *> cut((size_t)nchunks + 1);
    cut[0] = f.p;
    cut[nchunks] = f.p + f.n;
    for (int c = 1; c < nchunks; c++) {
        const char *q = f.p + f.n / nchunks * c;
        const char *nl = (const char *)memchr(q, '\n', f.p + f.n - q);
        cut[c] = nl ? nl + 1 : f.p + f.n;
        if (cut[c] < cut[c - 1]) cut[c] = cut[c - 1];
    }
    std::vector<std::vector<SeqRanges>> part((size_t)nchunks);
    std::vector<size_t> rows((size_t)nchunks, 0);
    std::vector<ParseError> errs((size_t)nchunks);
    std::vector<char> rest_flag((size_t)nchunks, 0);
    parallel_for(nchunks, [&](int c) {
        bool ar = false;
        rows[c] = parse_chunk(cut[c], cut[c + 1], g, cols, skip_missing, part[c], &ar, errs[c]);
        rest_flag[c] = ar;
    });
    for (int c = 0; c < nchunks; c++)
        if (errs[c].set) die(errs[c].msg);
    size_t total = 0;
    for (int c = 0; c < nchunks; c++) {
        total += rows[c];
        if (rest_flag[c] && any_rest) *any_rest = true;
    }
    // stitch: per seqname, the chunks in file order
    const int nseq = (int)g.names.size();
    out.assign(nseq, SeqRanges());
    parallel_for(nseq, [&](int s) {
        size_t n = 0;
        for (int c = 0; c < nchunks; c++) n += part[c][s].starts.size();
        SeqRanges &o = out[s];
        o.starts.reserve(n);
        o.ends.reserve(n);
        if (cols == Cols::Bed5) {
            o.score.reserve(n);
            o.valid.reserve(n);
        }
        if (cols == Cols::Bedlike) o.rest.reserve(n);
        for (int c = 0; c < nchunks; c++) {
            SeqRanges &q = part[c][s];
            o.starts.insert(o.starts.end(), q.starts.begin(), q.starts.end());
            o.ends.insert(o.ends.end(), q.ends.begin(), q.ends.end());
            o.score.insert(o.score.end(), q.score.begin(), q.score.end());
            o.valid.insert(o.valid.end(), q.valid.begin(), q.valid.end());
            o.rest.insert(o.rest.end(), q.rest.begin(), q.rest.end());
            q = SeqRanges();
        }
    });
    return total;
}

// Rows [0, n) rendered by `row(i, p)` (returns the end of what it wrote at p; at most `max_row` bytes) on all host
// threads into per-block buffers, then written in order.
void write_rows(const char *path, size_t n, size_t max_row, const std::function<char *(size_t, char *)> &row) {
    FILE *f = path ? fopen(path, "wb") : stdout;
    if (!f) die(std::string(path) + ": " + strerror(errno));
    const size_t block = 1 << 15;
    const size_t nblocks = (n + block - 1) / block;
    const size_t wave = (size_t)n_threads() * 2;
    std::vector<std::vector<char>> bufs(wave);
    for (size_t b0 = 0; b0 < nblocks; b0 += wave) {
        const int nb = (int)std::min(wave, nblocks - b0);
        parallel_for(nb, [&](int w) {
            const size_t lo = (b0 + w) * block, hi = std::min(n, lo + block);
            std::vector<char> &buf = bufs[w];
            buf.clear();
            std::vector<char> tmp(max_row);
            for (size_t i = lo; i < hi; i++) {
                char *e = row(i, tmp.data());
                buf.insert(buf.end(), tmp.data(), e);
            }
        });
        for (int w = 0; w < nb; w++) fwrite(bufs[w].data(), 1, bufs[w].size(), f);
    }
    if (f != stdout)
        fclose(f);
    else
        fflush(stdout);
}

void check(grb_ctx *ctx, int rc) {
    if (rc != GRB_OK) die(ctx ? grb_last_error(ctx) : "cannot create a CUDA context (granges_b200 has no CPU fallback)");
}

constexpr int OP_COLLAPSE = 1000;  // FloatOperation::Collapse: not a grb_op (its result is text)

int parse_op(const std::string &s) {
    // clap ValueEnum names of FloatOperation (src/data/operations.rs:35-51), kebab-case
    if (s == "sum") return GRB_OP_SUM;
    if (s == "sum-not-empty") return GRB_OP_SUM_NOT_EMPTY;
    if (s == "min") return GRB_OP_MIN;
    if (s == "max") return GRB_OP_MAX;
    if (s == "mean") return GRB_OP_MEAN;
    if (s == "median") return GRB_OP_MEDIAN;
    if (s == "collapse") return OP_COLLAPSE;  // text column: grb_collapse_f64, order defined as sort_ranges order
    die("invalid value '" + s + "' for '--func <FUNC>'");
}

struct Args {
    std::string genome, left, right, output;
    bool has_output = false, skip_missing = false, chop = false;
    std::vector<int> ops;
    uint32_t width = 0, step = 0;
    bool has_width = false;
    int gpus = 1;  // --gpus N (an extension of this mirror): spread `map` over the first N GPUs of the box
};

int cmd_map(const Args &a) {
    if (a.ops.empty()) die("No operation was specified. See granges map --help.");
    Genome g = read_genome(a.genome);
    Mapped lf(a.left), rf(a.right);
    std::vector<SeqRanges> L, R;
    size_t nl, nr;
    grb_ctx *ctx = nullptr;
    {
        // CUDA start-up (driver + context, a second or more on a big box) runs beside the parsing
        int ctx_rc = GRB_OK;
        std::thread init([&] {
            Phase ph("cuda context");
            ctx_rc = grb_ctx_create(0, &ctx);
        });
        {
            Phase ph("parse left");
            nl = parse_bed(lf, g, Cols::Bed3, a.skip_missing, L);
        }
        {
            Phase ph("parse right");
            nr = parse_bed(rf, g, Cols::Bed5, a.skip_missing, R);
        }
        init.join();
        if (nl == 0 || nr == 0) die("The input file had no rows.");
        check(nullptr, ctx_rc);
    }
    const int nseq = (int)g.names.size();
    std::vector<uint64_t> off(nseq + 1, 0);
    std::vector<uint32_t> ls, le;
    ls.reserve(nl);
    le.reserve(nl);
    for (int s = 0; s < nseq; s++) {
        off[s] = ls.size();
        ls.insert(ls.end(), L[s].starts.begin(), L[s].starts.end());
        le.insert(le.end(), L[s].ends.begin(), L[s].ends.end());
    }
    off[nseq] = ls.size();
    const int k = (int)a.ops.size();
    // numeric reductions go through one fused call; Collapse columns are text and come from the materialised pairs
    std::vector<int> nops, col_num(k, -1);
    bool any_collapse = false;
    for (int c = 0; c < k; c++) {
        if (a.ops[c] == OP_COLLAPSE) {
            any_collapse = true;
        } else {
            col_num[c] = (int)nops.size();
            nops.push_back(a.ops[c]);
        }
    }
    const int kn = (int)nops.size();
    std::vector<double> out(nl * (size_t)kn);
    std::vector<uint8_t> ov(nl * (size_t)kn);
    std::vector<const uint32_t *> ps(nseq), pe(nseq);
    std::vector<const double *> psc(nseq);
    std::vector<const uint8_t *> pva(nseq);
    std::vector<size_t> pn(nseq);
    for (int s = 0; s < nseq; s++) {
        ps[s] = R[s].starts.data();
        pe[s] = R[s].ends.data();
        psc[s] = R[s].score.data();
        pva[s] = R[s].valid.data();
        pn[s] = R[s].starts.size();
    }
    if (kn) {
        Phase ph("index builds + map_joins");
        // into_coitrees -> left_overlaps -> map_joins in one pipelined call: uploads, index builds, map kernels and
        // downloads of different seqnames overlap.  A seqname the right file never mentions (n = 0) has no container
        // (its groups are empty either way).
        if (a.gpus > 1) {
            std::vector<int> devs(a.gpus);
            for (int d = 0; d < a.gpus; d++) devs[d] = d;
            char err[512];
            if (grb_map_joins_whole_f64_mgpu(devs.data(), a.gpus, nseq, ps.data(), pe.data(), pn.data(), psc.data(), pva.data(), off.data(), ls.data(),
                                             le.data(), nops.data(), kn, out.data(), ov.data(), err, sizeof(err)) != GRB_OK)
                die(err);
        } else {
            check(ctx, grb_map_joins_whole_f64(ctx, nseq, ps.data(), pe.data(), pn.data(), psc.data(), pva.data(), off.data(), ls.data(), le.data(),
                                               nops.data(), kn, out.data(), ov.data()));
        }
    }
    std::vector<uint64_t> ctext_off;  // Collapse: row i is ctext[ctext_off[i] .. ctext_off[i + 1])
    std::string ctext;
    size_t max_collapse = 0;
    if (any_collapse) {
        Phase ph("collapse");
        ctext_off.assign(nl + 1, 0);
        for (int s = 0; s < nseq; s++) {
            const size_t n_s = (size_t)(off[s + 1] - off[s]);
            grb_index *ix = nullptr;
            if (pn[s]) {
                check(ctx, grb_index_build(ctx, ps[s], pe[s], nullptr, pn[s], &ix));
                check(ctx, grb_index_set_scores(ix, psc[s], pva[s], pn[s]));
            }
            grb_text *t = nullptr;
            check(ctx, grb_collapse_f64(ctx, ix, ls.data() + off[s], le.data() + off[s], n_s, &t));
            std::vector<uint64_t> to(n_s + 1);
            const size_t base = ctext.size();
            ctext.resize(base + grb_text_bytes(t));
            check(ctx, grb_text_copy(t, to.data(), ctext.data() + base));
            for (size_t i = 0; i <= n_s; i++) ctext_off[off[s] + i] = base + to[i];
            for (size_t i = 0; i < n_s; i++) max_collapse = std::max(max_collapse, (size_t)(to[i + 1] - to[i]));
            grb_text_destroy(t);
            grb_index_destroy(ix);
        }
    }
    Phase phw("write");
    size_t max_name = 0;
    for (auto &nm : g.names) max_name = std::max(max_name, nm.size());
    write_rows(a.has_output ? a.output.c_str() : nullptr, nl, max_name + 64 + (size_t)k * (420 + max_collapse), [&](size_t i, char *q) {
        const int sq = (int)(std::upper_bound(off.begin(), off.end(), (uint64_t)i) - off.begin()) - 1;
        const std::string &name = g.names[sq];
        memcpy(q, name.data(), name.size());
        q += name.size();
        *q++ = '\t';
        q = grh::format_u32(q, ls[i]);
        *q++ = '\t';
        q = grh::format_u32(q, le[i]);
        for (int c = 0; c < k; c++) {
            *q++ = '\t';
            if (col_num[c] < 0) {
                const size_t len = (size_t)(ctext_off[i + 1] - ctext_off[i]);
                memcpy(q, ctext.data() + ctext_off[i], len);
                q += len;
            } else if (ov[i * kn + col_num[c]])
                q = grh::format_f64(q, out[i * kn + col_num[c]]);
            else
                *q++ = '.';
        }
        *q++ = '\n';
        return q;
    });
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
    // the kept rows, in order, straight from the device-side compaction
    std::vector<uint64_t> kept(nl);
    uint64_t nk = 0;
    if (nl) check(ctx, grb_filter_overlaps_select(ctx, idx.data(), off.data(), nseq, ls.data(), le.data(), 0, kept.data(), &nk));
    size_t max_name = 0, max_rest = 0;
    for (auto &nm : g.names) max_name = std::max(max_name, nm.size());
    for (auto &sr : L)
        for (auto &r : sr.rest) max_rest = std::max(max_rest, r.size());
    write_rows(a.has_output ? a.output.c_str() : nullptr, nk, max_name + 64 + max_rest, [&](size_t r, char *q) {
        const uint64_t i = kept[r];
        const int sq = (int)(std::upper_bound(off.begin(), off.end(), i) - off.begin()) - 1;
        const std::string &name = g.names[sq];
        std::string_view rest = L[sq].rest[i - off[sq]];
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
        return q;
    });
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
    size_t max_name = 0;
    for (auto &nm : g.names) max_name = std::max(max_name, nm.size());
    write_rows(a.has_output ? a.output.c_str() : nullptr, n, max_name + 40, [&](size_t i, char *q) {
        const std::string &name = g.names[seq[i]];
        memcpy(q, name.data(), name.size());
        q += name.size();
        *q++ = '\t';
        q = grh::format_u32(q, st[i]);
        *q++ = '\t';
        q = grh::format_u32(q, en[i]);
        *q++ = '\n';
        return q;
    });
    grb_ranges_destroy(r);
    grb_ctx_destroy(ctx);
    return 0;
}

// ---- file-order records (merge, feature-density): the merging iterators stream the file as it is ----------

struct Records {
    std::vector<std::string_view> seqname, name;  // name: 4th column (BED4 / BED5)
    std::vector<uint32_t> starts, ends;
    std::vector<double> score;  // BED5
    std::vector<uint8_t> valid;
    int ncols = 0;  // of the first record: 3, 4 or 5 (src/io/parsers/detect.rs:39-53 tries BED5, BED4, BED3)
};

// one chunk of the file (cut at line boundaries); ncols = 0: take it from the first record
Records parse_records_chunk(const char *p, const char *end, int want_cols /* 0 = detect */) {
    Records r;
    while (p < end) {
        const char *line = p;
        const char *nl = (const char *)memchr(p, '\n', end - p);
        const char *le = nl ? nl : end;
        p = le + 1;
        if (le > line && le[-1] == '\r') le--;
        if (le == line || *line == '#') continue;
        const char *col[6];
        int nc = 0;
        col[nc++] = line;
        for (const char *q = line; q < le && nc < 6; q++)
            if (*q == '\t') col[nc++] = q + 1;
        int total_cols = nc;
        if (nc == 6) {
            total_cols = 6;
            for (const char *q = col[5]; q < le; q++)
                if (*q == '\t') total_cols++;
        }
        auto col_end = [&](int c) { return c + 1 < nc ? col[c + 1] - 1 : le; };
        if (nc < 3) die("expected at least 3 tab-separated columns");
        if (r.ncols == 0) {
            r.ncols = want_cols ? want_cols : total_cols;
            if (!want_cols && r.ncols > 5) die("Unsupported genomic ranges file format: BED-like input with more than five columns (the reference's merge stops at todo!())");
        }
        if (total_cols < r.ncols) die("a record has fewer columns than the first one");
        uint64_t s = 0, e = 0;
        bool bad = col_end(1) == col[1] || col_end(2) == col[2];
        for (const char *q = col[1]; q < col_end(1); q++) {
            bad |= *q < '0' || *q > '9';
            s = s * 10 + (uint64_t)(*q - '0');
        }
        for (const char *q = col[2]; q < col_end(2); q++) {
            bad |= *q < '0' || *q > '9';
            e = e * 10 + (uint64_t)(*q - '0');
        }
        if (bad) die("invalid start / end coordinate");
        if (e <= s || e > 0x7fffffffull) die("invalid range: end must be > start and fit i32");
        r.seqname.emplace_back(col[0], col_end(0) - col[0]);
        r.starts.push_back((uint32_t)s);
        r.ends.push_back((uint32_t)e);
        if (r.ncols >= 4) r.name.emplace_back(col[3], col_end(3) - col[3]);
        if (r.ncols >= 5) {
            const char *a = col[4], *b = col_end(4);
            if (b - a == 1 && *a == '.') {
                r.score.push_back(0.0);
                r.valid.push_back(0);
            } else {
                char tmp[64];
                size_t len = (size_t)(b - a);
                if (len == 0 || len >= sizeof(tmp)) die("parsing error: invalid float literal");
                memcpy(tmp, a, len);
                tmp[len] = 0;
                char *ep;
                double v = strtod(tmp, &ep);
                if (ep == tmp || *ep) die("parsing error: invalid float literal");
                r.score.push_back(v);
                r.valid.push_back(1);
            }
        }
    }
    return r;
}

// The file is cut into chunks at line boundaries, the chunks are parsed on all host threads and stitched back in file
// order; the column count is that of the first record (src/io/parsers/detect.rs:39-53).
Records parse_records(const Mapped &f, int want_cols /* 0 = detect */) {
    const char *p = f.p, *end = f.p + f.n;
    if (want_cols == 0) {
        // detect on the first record alone
        const char *q = p;
        while (q < end) {
            const char *nl = (const char *)memchr(q, '\n', end - q);
            const char *le = nl ? nl : end;
            if (le > q && *q != '#' && !(le - q == 1 && *q == '\r')) {
                Records one = parse_records_chunk(q, le, 0);
                want_cols = one.ncols;
                break;
            }
            q = le + 1;
        }
        if (want_cols == 0) return Records();
    }
    const int nchunks = f.n < (1u << 20) ? 1 : n_threads() * 4;
    std::vector<const char *> cut((size_t)nchunks + 1);
    cut[0] = p;
    cut[nchunks] = end;
    for (int c = 1; c < nchunks; c++) {
        const char *q = p + f.n / nchunks * c;
        const char *nl = (const char *)memchr(q, '\n', end - q);
        cut[c] = nl ? nl + 1 : end;
        if (cut[c] < cut[c - 1]) cut[c] = cut[c - 1];
    }
    std::vector<Records> part((size_t)nchunks);
    parallel_for(nchunks, [&](int c) { part[c] = parse_records_chunk(cut[c], cut[c + 1], want_cols); });
    if (nchunks == 1) {
        part[0].ncols = want_cols;
        return std::move(part[0]);
    }
    Records r;
    r.ncols = want_cols;
    size_t total = 0;
    for (auto &q : part) total += q.starts.size();
    r.seqname.reserve(total);
    r.starts.reserve(total);
    r.ends.reserve(total);
    if (want_cols >= 4) r.name.reserve(total);
    if (want_cols >= 5) {
        r.score.reserve(total);
        r.valid.reserve(total);
    }
    for (auto &q : part) {
        r.seqname.insert(r.seqname.end(), q.seqname.begin(), q.seqname.end());
        r.name.insert(r.name.end(), q.name.begin(), q.name.end());
        r.starts.insert(r.starts.end(), q.starts.begin(), q.starts.end());
        r.ends.insert(r.ends.end(), q.ends.begin(), q.ends.end());
        r.score.insert(r.score.end(), q.score.begin(), q.score.end());
        r.valid.insert(r.valid.end(), q.valid.begin(), q.valid.end());
        q = Records();
    }
    return r;
}

// maximal stretches of consecutive records with the same seqname
std::vector<uint64_t> seqname_runs(const std::vector<std::string_view> &seqname) {
    std::vector<uint64_t> ro{0};
    for (size_t i = 1; i < seqname.size(); i++)
        if (seqname[i] != seqname[i - 1]) ro.push_back(i);
    if (!seqname.empty()) ro.push_back(seqname.size());
    return ro;
}

// `granges merge` -- Merge::run, src/commands.rs:616-673: BED3 -> merged ranges; BED4 -> names joined with ",";
// BED5 -> FloatOperation over the merged records' scores.
int cmd_merge(const std::string &bedfile, long long distance, const std::vector<int> &ops, const Args &a) {
    Mapped f(bedfile);
    grb_ctx *ctx = nullptr;
    int ctx_rc = GRB_OK;
    std::thread init([&] { ctx_rc = grb_ctx_create(0, &ctx); });  // CUDA start-up runs beside the parsing
    Records r = parse_records(f, 0);
    init.join();
    if (r.starts.empty()) die("File '" + bedfile + "' is empty.");
    if (r.ncols == 5 && ops.size() != 1) die("BED5 input needs exactly one --func (the reference unwraps it: src/commands.rs:656)");
    if (r.ncols == 5 && ops[0] == OP_COLLAPSE) die("merge --func collapse is not offered by this mirror (map --func collapse is)");
    check(nullptr, ctx_rc);
    std::vector<uint64_t> ro = seqname_runs(r.seqname);
    grb_merged *mg = nullptr;
    check(ctx, grb_merge_ranges(ctx, r.starts.data(), r.ends.data(), ro.data(), (int)ro.size() - 1, distance, &mg));
    const size_t m = grb_merged_len(mg);
    std::vector<uint32_t> ms(m), me(m), first(m + 1);
    check(ctx, grb_merged_copy(mg, ms.data(), me.data(), first.data(), nullptr));
    std::vector<double> val;
    std::vector<uint8_t> ok;
    if (r.ncols == 5) {
        val.resize(m);
        ok.resize(m);
        check(ctx, grb_merged_map_f64(ctx, mg, r.score.data(), r.valid.data(), ops.data(), 1, val.data(), ok.data()));
    }
    size_t max_row = 64;
    for (auto &nm : r.seqname) max_row = std::max(max_row, nm.size() + 96);
    if (r.ncols == 5) max_row += 420;  // grh::format_f64: positional notation, up to ~330 characters (1e-300, 1e300)
    if (r.ncols == 4) {
        size_t worst = 0;
        for (size_t g = 0; g < m; g++) {
            size_t len = 0;
            for (uint32_t i = first[g]; i < first[g + 1]; i++) len += r.name[i].size() + 1;
            worst = std::max(worst, len);
        }
        max_row += worst;
    }
    write_rows(a.has_output ? a.output.c_str() : nullptr, m, max_row, [&](size_t g, char *q) {
        std::string_view nm = r.seqname[first[g]];
        memcpy(q, nm.data(), nm.size());
        q += nm.size();
        *q++ = '\t';
        q = grh::format_u32(q, ms[g]);
        *q++ = '\t';
        q = grh::format_u32(q, me[g]);
        if (r.ncols == 4) {
            *q++ = '\t';
            for (uint32_t i = first[g]; i < first[g + 1]; i++) {
                if (i > first[g]) *q++ = ',';
                memcpy(q, r.name[i].data(), r.name[i].size());
                q += r.name[i].size();
            }
        } else if (r.ncols == 5) {
            *q++ = '\t';
            if (ok[g])
                q = grh::format_f64(q, val[g]);
            else
                *q++ = '.';
        }
        *q++ = '\n';
        return q;
    });
    grb_merged_destroy(mg);
    grb_ctx_destroy(ctx);
    return 0;
}

// `granges feature-density` (non-exclusive) -- FeatureDensity::feature_density, src/commands.rs:796-858: the BED4
// records are split by feature name, each feature's ranges are bookend-merged in file order, and every window
// gets the number of basepairs each feature covers in it (sum of overlap widths over merged = disjoint ranges).
// Columns are the features in sorted order (the reference iterates a HashMap: its order is unspecified).
int cmd_feature_density(const std::string &bedfile, bool headers, const Args &a) {
    if (!a.has_width) die("the following required arguments were not provided: --width <WIDTH>");
    Genome g = read_genome(a.genome);
    Mapped f(bedfile);
    grb_ctx *ctx = nullptr;
    int ctx_rc = GRB_OK;
    std::thread init([&] { ctx_rc = grb_ctx_create(0, &ctx); });  // CUDA start-up runs beside the parsing
    Records r = parse_records(f, 4);
    const int nseq = (int)g.names.size();
    std::map<std::string_view, std::vector<uint32_t>> by_feature;  // feature -> record positions, file order
    for (size_t i = 0; i < r.name.size(); i++) by_feature[r.name[i]].push_back((uint32_t)i);
    init.join();
    check(nullptr, ctx_rc);
    grb_ranges *w = nullptr;
    check(ctx, grb_windows(ctx, g.lens.data(), nseq, a.width, a.step, a.chop ? 1 : 0, &w));
    const size_t nw = grb_ranges_len(w);
    std::vector<uint64_t> woff(nseq + 1);
    check(ctx, grb_ranges_seq_offsets(w, woff.data()));
    std::vector<uint32_t> wseq(nw), wst(nw), wen(nw);
    check(ctx, grb_ranges_copy(w, wseq.data(), wst.data(), wen.data()));
    const size_t nf = by_feature.size();
    std::vector<std::vector<double>> cols(nf);
    std::vector<std::string_view> feature_names;
    size_t fi = 0;
    for (auto &kv : by_feature) {
        feature_names.push_back(kv.first);
        const std::vector<uint32_t> &pos = kv.second;
        std::vector<uint32_t> st(pos.size()), en(pos.size());
        std::vector<std::string_view> sq(pos.size());
        for (size_t j = 0; j < pos.size(); j++) {
            st[j] = r.starts[pos[j]];
            en[j] = r.ends[pos[j]];
            sq[j] = r.seqname[pos[j]];
        }
        std::vector<uint64_t> ro = seqname_runs(sq);
        grb_merged *mg = nullptr;
        check(ctx, grb_merge_ranges(ctx, st.data(), en.data(), ro.data(), (int)ro.size() - 1, 0, &mg));
        const size_t m = grb_merged_len(mg);
        std::vector<uint32_t> ms(m), me(m), first(m + 1);
        check(ctx, grb_merged_copy(mg, ms.data(), me.data(), first.data(), nullptr));
        grb_merged_destroy(mg);
        // into the genome-keyed container (GRangesEmpty::from_iter_ok): per seqname, in arrival order
        std::vector<std::vector<uint32_t>> ss(nseq), ee(nseq);
        for (size_t k = 0; k < m; k++) {
            auto it = g.id.find(sq[first[k]]);
            if (it == g.id.end())
                die("The sequence name '" + std::string(sq[first[k]]) +
                    "' is not found within the provided ranges container. Check the sequence names for typos or missing entries.");
            ss[it->second].push_back(ms[k]);
            ee[it->second].push_back(me[k]);
        }
        std::vector<grb_index *> idx(nseq, nullptr);
        for (int s = 0; s < nseq; s++)
            if (!ss[s].empty()) check(ctx, grb_index_build(ctx, ss[s].data(), ee[s].data(), nullptr, ss[s].size(), &idx[s]));
        cols[fi].resize(nw);
        std::vector<uint8_t> okv(nw);
        const int op = GRB_OP_OVERLAP_BP;
        if (nw)
            check(ctx, grb_map_joins_f64_batch(ctx, idx.data(), woff.data(), nseq, grb_ranges_dev_starts(w), grb_ranges_dev_ends(w), &op, 1,
                                               cols[fi].data(), okv.data()));
        for (auto ix : idx) grb_index_destroy(ix);
        fi++;
    }
    FILE *out = a.has_output ? fopen(a.output.c_str(), "wb") : stdout;
    if (!out) die(a.output + ": " + strerror(errno));
    if (headers) {
        fputs("chrom\tstart\tend", out);
        for (auto &nm : feature_names) {
            fputc('\t', out);
            fwrite(nm.data(), 1, nm.size(), out);
        }
        fputc('\n', out);
        fflush(out);
    }
    if (out != stdout) fclose(out);
    size_t max_name = 0;
    for (auto &nm : g.names) max_name = std::max(max_name, nm.size());
    // write_rows reopens the file: append after the header
    struct Appender {
        static void rows(const char *path, bool append, size_t n, size_t max_row, const std::function<char *(size_t, char *)> &row) {
            FILE *fo = path ? fopen(path, append ? "ab" : "wb") : stdout;
            if (!fo) die(std::string(path) + ": " + strerror(errno));
            std::vector<char> tmp(max_row);
            for (size_t i = 0; i < n; i++) {
                char *e = row(i, tmp.data());
                fwrite(tmp.data(), 1, (size_t)(e - tmp.data()), fo);
            }
            if (fo != stdout)
                fclose(fo);
            else
                fflush(stdout);
        }
    };
    Appender::rows(a.has_output ? a.output.c_str() : nullptr, headers, nw, max_name + 40 + 12 * nf, [&](size_t i, char *q) {
        const std::string &name = g.names[wseq[i]];
        memcpy(q, name.data(), name.size());
        q += name.size();
        *q++ = '\t';
        q = grh::format_u32(q, wst[i]);
        *q++ = '\t';
        q = grh::format_u32(q, wen[i]);
        for (size_t c = 0; c < nf; c++) {
            *q++ = '\t';
            q = grh::format_u32(q, (uint32_t)cols[c][i]);
        }
        *q++ = '\n';
        return q;
    });
    grb_ranges_destroy(w);
    grb_ctx_destroy(ctx);
    return 0;
}

// `granges random-bed` (dev command of the reference, src/commands.rs:563-588 + src/test_utilities.rs:60-164): seqname
// uniform over the genome file's names, length U{1..999}, start uniform, optional BED5 name + score U[0,1).  The
// generator is a counter-based splitmix64 (the reference's StdRng stream is not reproducible without the crate).
uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

int cmd_random_bed(const Args &a, size_t num, bool sort, bool scores, uint64_t seed) {
    Genome g = read_genome(a.genome);
    const size_t nseq = g.names.size();
    if (nseq == 0) die("The input file had no rows.");
    struct Row {
        uint32_t seq, start, end;
        uint64_t id;
    };
    std::vector<Row> rows(num);
    parallel_for(64, [&](int c) {
        for (size_t i = (size_t)c * num / 64; i < (size_t)(c + 1) * num / 64; i++) {
            const uint64_t b = seed * 0x100000001B3ull + i * 4;
            const uint32_t sq = (uint32_t)(splitmix(b) % nseq);
            const uint32_t L = g.lens[sq];
            uint32_t len = 1 + (uint32_t)(splitmix(b + 1) % 999);
            if (len > L) len = L ? L : 1;
            const uint32_t st = (uint32_t)(splitmix(b + 2) % ((uint64_t)L - len + 1));
            rows[i] = Row{sq, st, st + len, i};
        }
    });
    if (sort)
        std::sort(rows.begin(), rows.end(), [](const Row &x, const Row &y) {
            if (x.seq != y.seq) return x.seq < y.seq;
            if (x.start != y.start) return x.start < y.start;
            if (x.end != y.end) return x.end < y.end;
            return x.id < y.id;
        });
    size_t max_name = 0;
    for (auto &nm : g.names) max_name = std::max(max_name, nm.size());
    write_rows(a.has_output ? a.output.c_str() : nullptr, num, max_name + 96, [&](size_t i, char *q) {
        const Row &r = rows[i];
        const std::string &name = g.names[r.seq];
        memcpy(q, name.data(), name.size());
        q += name.size();
        *q++ = '\t';
        q = grh::format_u32(q, r.start);
        *q++ = '\t';
        q = grh::format_u32(q, r.end);
        if (scores) {
            uint64_t h = splitmix(seed * 0x100000001B3ull + r.id * 4 + 3);
            *q++ = '\t';
            for (int c = 0; c < 8; c++) {
                *q++ = (char)('a' + (h % 26));
                h /= 26;
            }
            *q++ = '\t';
            const double sc = (double)(splitmix(h + r.id) >> 11) * (1.0 / 9007199254740992.0);
            q = grh::format_f64(q, sc);
        }
        *q++ = '\n';
        return q;
    });
    return 0;
}

void usage() {
    fprintf(stderr,
            "granges (B200 mirror of the overlap-join commands)\n"
            "  granges map     -g <genome> -l <left.bed> -r <right.bed5> -f <op[,op..]> [-o <out>] [-s] [--gpus N]\n"
            "  granges filter  -g <genome> -l <left.bed> -r <right.bed> [-o <out>] [-s]\n"
            "  granges windows -g <genome> -w <width> [-s <step>] [-c] [-o <out>]\n"
            "  granges merge   -b <bedfile> [-d <distance>] [-f <op>] [-o <out>]\n"
            "  granges feature-density -g <genome> -b <bed4> -w <width> [-s <step>] [-c] [-l] [-o <out>]\n"
            "  granges random-bed <genome> -n <num> [-o <out>] [-s] [-c] [--seed <n>]\n"
            "ops: sum, sum-not-empty, min, max, mean, median\n");
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) {
        usage();
        return 2;
    }
    const std::string cmd = argv[1];
    const bool windows = cmd == "windows" || cmd == "feature-density";
    Args a;
    std::string bedfile;
    long long distance = 0;
    bool headers = false, exclusive = false;
    if (cmd == "random-bed") {
        size_t num = 0;
        bool sort = false, scores = false;
        uint64_t seed = 13;  // TEST_SEED default of the reference (src/test_utilities.rs:34-40)
        for (int i = 2; i < argc; i++) {
            std::string f = argv[i];
            if ((f == "-n" || f == "--num") && i + 1 < argc)
                num = strtoull(argv[++i], nullptr, 10);
            else if ((f == "-o" || f == "--output") && i + 1 < argc) {
                a.output = argv[++i];
                a.has_output = true;
            } else if (f == "--seed" && i + 1 < argc)
                seed = strtoull(argv[++i], nullptr, 10);
            else if (f == "-s" || f == "--sort")
                sort = true;
            else if (f == "-c" || f == "--scores")
                scores = true;
            else if (f[0] != '-' && a.genome.empty())
                a.genome = f;
            else
                die("unexpected argument '" + f + "' found");
        }
        if (a.genome.empty()) die("the following required arguments were not provided: <GENOME>");
        return cmd_random_bed(a, num, sort, scores, seed);
    }
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
        else if (cmd == "feature-density" && (f == "-l" || f == "--headers"))
            headers = true;
        else if (cmd == "feature-density" && (f == "-e" || f == "--exclusive"))
            exclusive = true;
        else if (f == "-b" || f == "--bedfile")
            bedfile = need();
        else if (cmd == "merge" && (f == "-d" || f == "--distance"))
            distance = strtoll(need().c_str(), nullptr, 10);
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
        else if (cmd == "map" && f == "--gpus") {
            a.gpus = (int)strtol(need().c_str(), nullptr, 10);
            if (a.gpus < 1 || a.gpus > 64) die("invalid value for '--gpus <N>'");
        }
        else if (f == "-h" || f == "--help") {
            usage();
            return 0;
        } else
            die("unexpected argument '" + f + "' found");
    }
    if (cmd == "merge") {
        if (bedfile.empty()) die("the following required arguments were not provided: --bedfile <BEDFILE>");
        return cmd_merge(bedfile, distance, a.ops, a);
    }
    if (a.genome.empty()) die("the following required arguments were not provided: --genome <GENOME>");
    if (cmd == "feature-density") {
        if (bedfile.empty()) die("the following required arguments were not provided: --bedfile <BEDFILE>");
        if (exclusive) die("--exclusive is experimental and untested in the reference (src/commands.rs:779-780); this mirror does not offer it");
        return cmd_feature_density(bedfile, headers, a);
    }
    if (cmd == "map") {
        if (a.left.empty() || a.right.empty()) die("the following required arguments were not provided: --left <LEFT> --right <RIGHT>");
        return cmd_map(a);
    }
    if (cmd == "filter") {
        if (a.left.empty() || a.right.empty()) die("the following required arguments were not provided: --left <LEFT> --right <RIGHT>");
        return cmd_filter(a);
    }
    if (cmd == "windows") return cmd_windows(a);
    die("unrecognized subcommand '" + cmd + "' (this mirror implements map, filter, windows, merge and feature-density)");
}
