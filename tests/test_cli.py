"""The `granges` CLI mirror (map / filter / windows) against text rendered from the oracle's results:
byte-for-byte TSV, the reference's row order (genome-file order, then input order), "." for NoValue,
Rust-style floats.  Flags as in the reference's clap definitions (src/main/mod.rs:96-214)."""
import os
import subprocess

import numpy as np
import pytest

from tests.test_host_format import rust_display

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "granges_b200", "granges")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(CLI):
        import __graft_entry__ as g

        g.build()


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle

    return oracle


def run(*args, check=True):
    r = subprocess.run([CLI, *map(str, args)], capture_output=True, text=True, timeout=120)
    if check:
        assert r.returncode == 0, r.stderr
    return r


def make_files(tmp_path, seed=5, missing_scores=True, extra_left_cols=False):
    rng = np.random.default_rng(seed)
    genome = {"chr1": 300_000, "chr2": 150_000, "chrX": 90_000, "chrEmpty": 5000}
    (tmp_path / "genome.tsv").write_text("".join(f"{k}\t{v}\n" for k, v in genome.items()))
    left, right = {}, {}
    llines, rlines = ["# a comment line"], []
    for name, L in genome.items():
        if name == "chrEmpty":
            continue
        nl, nr = int(rng.integers(200, 400)), int(rng.integers(300, 600))
        if name == "chrX":
            nr = 0  # the right file never mentions chrX
        ls = rng.integers(0, L - 1000, nl)
        le = ls + rng.integers(1, 1000, nl)
        rs = rng.integers(0, L - 1000, nr)
        re = rs + rng.integers(1, 1000, nr)
        sc = np.round(rng.random(nr) * 100, int(rng.integers(0, 4)))
        valid = np.ones(nr, bool) if not missing_scores else rng.random(nr) > 0.15
        left[name] = (ls.astype(np.uint32), le.astype(np.uint32))
        right[name] = (rs.astype(np.uint32), re.astype(np.uint32), sc, valid)
    # interleave seqnames in the files: output order must still be genome order, then input order
    order = [(n, i) for n in left for i in range(len(left[n][0]))]
    rng.shuffle(order)
    lrest = {}
    for n, i in order:
        rest = f"\tfeat{i}\t+" if extra_left_cols else ""
        lrest[(n, i)] = rest
        llines.append(f"{n}\t{left[n][0][i]}\t{left[n][1][i]}{rest}")
    # per-seqname insertion order = order of appearance in the file
    lorder = {n: [i for (m, i) in order if m == n] for n in left}
    rorder = {}
    ro = [(n, i) for n in right for i in range(len(right[n][0]))]
    rng.shuffle(ro)
    for n, i in ro:
        rs, re, sc, valid = right[n]
        s = repr(float(sc[i])) if valid[i] else "."
        rlines.append(f"{n}\t{rs[i]}\t{re[i]}\tr{i}\t{s}")
    rorder = {n: [i for (m, i) in ro if m == n] for n in right}
    (tmp_path / "left.bed").write_text("\n".join(llines) + "\n")
    (tmp_path / "right.bed").write_text("\n".join(rlines) + "\n")
    return genome, left, right, lorder, rorder, lrest


def test_map_matches_oracle_text(tmp_path, orc):
    genome, left, right, lorder, rorder, _ = make_files(tmp_path)
    ops = ["min", "max", "mean", "sum", "sum-not-empty", "median"]
    out = tmp_path / "out.tsv"
    run("map", "--genome", tmp_path / "genome.tsv", "--left", tmp_path / "left.bed", "--right", tmp_path / "right.bed", "--func",
        ",".join(ops), "--output", out)
    want = []
    for name in genome:
        if name not in left:
            continue
        ls, le = (a[lorder[name]] for a in left[name])
        rs, re, sc, valid = (a[rorder[name]] for a in right[name])
        t = orc.Tree(rs, re) if len(rs) else None
        o, ov = orc.map_joins(t, sc if t is not None else None, valid.astype(np.uint8) if t is not None else None, ls, le, ops)
        for i in range(len(ls)):
            vals = [rust_display(float(o[i, c])) if ov[i, c] else "." for c in range(len(ops))]
            want.append("\t".join([name, str(ls[i]), str(le[i])] + vals))
    got = out.read_text().splitlines()
    assert len(got) == len(want)
    # sums / means: exactly-rounded on the GPU vs sequential on the CPU may differ in the last digit
    for g, w in zip(got, want):
        if g != w:
            gf, wf = g.split("\t"), w.split("\t")
            assert gf[:3] == wf[:3] and len(gf) == len(wf)
            for c, (a, b) in enumerate(zip(gf[3:], wf[3:])):
                if a != b:
                    assert ops[c] in ("mean", "sum", "sum-not-empty") and abs(float(a) - float(b)) <= 1e-12 * abs(float(b)), (g, w)


def test_map_collapse_and_gpus(tmp_path, orc):
    """`map --func max,collapse,mean`: the Collapse column is text -- the group's non-missing scores in sort_ranges order
    (src/join.rs:121-128; the reference leaves the order open), joined by "," -- beside numeric columns; `--gpus N` (the
    in-process multi-GPU call) prints the same bytes as one GPU."""
    import torch

    genome, left, right, lorder, rorder, _ = make_files(tmp_path, seed=12)
    base = ["map", "--genome", tmp_path / "genome.tsv", "--left", tmp_path / "left.bed", "--right", tmp_path / "right.bed", "--func", "max,collapse,mean"]
    got = run(*base).stdout
    want = []
    for name in genome:
        if name not in left:
            continue
        ls, le = (a[lorder[name]] for a in left[name])
        rs, re, sc, valid = (a[rorder[name]] for a in right[name])
        t = orc.Tree(rs, re) if len(rs) else None
        o, ov = orc.map_joins(t, sc if t is not None else None, valid.astype(np.uint8) if t is not None else None, ls, le, ["max", "mean"])
        off, idx, _, _ = orc.left_overlaps(t, ls, le)
        for i in range(len(ls)):
            members = [int(d) for d in idx[int(off[i]):int(off[i + 1])] if valid[int(d)]]
            cols = [rust_display(float(o[i, 0])) if ov[i, 0] else ".", ",".join(rust_display(float(sc[d])) for d in members),
                    rust_display(float(o[i, 1])) if ov[i, 1] else "."]
            want.append("\t".join([name, str(ls[i]), str(le[i])] + cols))
    rows = got.splitlines()
    assert len(rows) == len(want)
    for g, w in zip(rows, want):
        gf, wf = g.split("\t"), w.split("\t")
        assert gf[:5] == wf[:5], (g, w)
        assert gf[5] == wf[5] or abs(float(gf[5]) - float(wf[5])) <= 1e-12 * abs(float(wf[5]))
    for n in sorted({1, torch.cuda.device_count()}):
        assert run(*base, "--gpus", n).stdout == got


def test_map_bedtools_doc_example(tmp_path, goldens):
    g = goldens["bedtools_map_example"]
    (tmp_path / "g.tsv").write_text("chr1\t1000\n")
    (tmp_path / "a.bed").write_text("".join(f"chr1\t{s}\t{e}\n" for s, e, *_ in g["left"]["chr1"]))
    (tmp_path / "b.bed").write_text("".join(f"chr1\t{s}\t{e}\tb{i}\t{sc}\n" for i, (s, e, sc) in enumerate(g["right"]["chr1"])))
    r = run("map", "-g", tmp_path / "g.tsv", "-l", tmp_path / "a.bed", "-r", tmp_path / "b.bed", "-f", "mean")
    rows = [line.split("\t") for line in r.stdout.splitlines()]
    assert [row[3] for row in rows] == ["." if m is None else rust_display(m) for m in g["mean"]]


def test_filter_matches_oracle(tmp_path, orc):
    genome, left, right, lorder, rorder, lrest = make_files(tmp_path, seed=9, extra_left_cols=True)
    r = run("filter", "-g", tmp_path / "genome.tsv", "-l", tmp_path / "left.bed", "-r", tmp_path / "right.bed")
    want = []
    for name in genome:
        if name not in left or len(right[name][0]) == 0:
            continue  # no right container for this seqname: the semi-join drops its lefts
        ls, le = (a[lorder[name]] for a in left[name])
        t = orc.Tree(right[name][0], right[name][1])
        keep = orc.filter_overlaps(t, ls, le)
        for j, i in enumerate(lorder[name]):
            if keep[j]:
                want.append(f"{name}\t{ls[j]}\t{le[j]}{lrest[(name, i)]}")
    assert r.stdout.splitlines() == want


def test_windows_matches_oracle(tmp_path, orc):
    lens = {"chr1": 100_003, "chr2": 999, "chr3": 1000, "chr4": 1}
    (tmp_path / "g.tsv").write_text("".join(f"{k}\t{v}\n" for k, v in lens.items()))
    names = list(lens)
    for width, step, chop in [(1000, 500, False), (1000, 500, True), (1000, None, False), (333, 100, True)]:
        args = ["windows", "-g", tmp_path / "g.tsv", "-w", width] + (["-s", step] if step else []) + (["--chop"] if chop else [])
        got = run(*args).stdout.splitlines()
        seq, s, e = orc.windows(list(lens.values()), width, step, chop)
        assert got == [f"{names[q]}\t{a}\t{b}" for q, a, b in zip(seq, s, e)]


def test_cli_errors(tmp_path):
    (tmp_path / "g.tsv").write_text("chr1\t1000\n")
    (tmp_path / "a.bed").write_text("chr1\t1\t5\nchr9\t1\t5\n")
    (tmp_path / "b.bed").write_text("chr1\t2\t6\tx\t1.5\n")
    r = run("map", "-g", tmp_path / "g.tsv", "-l", tmp_path / "a.bed", "-r", tmp_path / "b.bed", "-f", "sum", check=False)
    assert r.returncode == 1 and r.stderr.startswith("Error: The sequence name 'chr9' is not found")
    r = run("map", "-g", tmp_path / "g.tsv", "-l", tmp_path / "a.bed", "-r", tmp_path / "b.bed", "-f", "sum", "--skip-missing")
    assert r.stdout == "chr1\t1\t5\t1.5\n"
    r = run("map", "-g", tmp_path / "g.tsv", "-l", tmp_path / "a.bed", "-r", tmp_path / "b.bed", "-s", check=False)
    assert r.returncode == 1 and "No operation was specified" in r.stderr
    (tmp_path / "e.bed").write_text("")
    r = run("map", "-g", tmp_path / "g.tsv", "-l", tmp_path / "e.bed", "-r", tmp_path / "b.bed", "-f", "sum", check=False)
    assert r.returncode == 1 and "The input file had no rows." in r.stderr


# ---- merge / feature-density -------------------------------------------------------------------

def _merge_file(tmp_path, ncols, seed=9, n=4000):
    rng = np.random.default_rng(seed)
    names = ["chr1", "chr2", "chr10", "chrX"]
    rows = []
    for nm in names:
        k = int(rng.integers(n // 8, n // 3))
        st = np.sort(rng.integers(0, 40 * k, k))
        en = st + rng.integers(1, 90, k)
        for a, b in zip(st, en):
            sc = "." if rng.random() < 0.1 else repr(float(np.round(rng.standard_normal() * 10, 2)))
            rows.append((nm, int(a), int(b), f"f{int(rng.integers(0, 5))}", sc))
    path = tmp_path / f"in{ncols}.bed"
    path.write_text("".join("\t".join(map(str, r[:ncols])) + "\n" for r in rows))
    return path, rows


@pytest.mark.parametrize("distance", [0, 30, -5])
def test_cli_merge_bed3_and_bed4(tmp_path, orc, distance):
    """BED3 -> merged ranges; BED4 -> the merged records' names joined with "," (src/commands.rs:629-645)."""
    for ncols in (3, 4):
        path, rows = _merge_file(tmp_path, ncols)
        names = sorted({r[0] for r in rows})
        seq = np.array([names.index(r[0]) for r in rows], np.uint32)
        s, e = np.array([r[1] for r in rows], np.uint32), np.array([r[2] for r in rows], np.uint32)
        mseq, ms, me, first = orc.merge(seq, s, e, distance)
        bounds = list(first) + [len(rows)]
        want = []
        for g in range(len(ms)):
            line = f"{names[mseq[g]]}\t{ms[g]}\t{me[g]}"
            if ncols == 4:
                line += "\t" + ",".join(rows[i][3] for i in range(int(bounds[g]), int(bounds[g + 1])))
            want.append(line)
        got = run("merge", "--bedfile", path, "--distance", distance).stdout
        assert got == "\n".join(want) + "\n"


@pytest.mark.parametrize("op", ["sum", "mean", "median", "min", "max"])
def test_cli_merge_bed5(tmp_path, orc, op):
    """BED5 -> FloatOperation over the merged records' non-missing scores (src/commands.rs:646-662)."""
    path, rows = _merge_file(tmp_path, 5)
    names = sorted({r[0] for r in rows})
    seq = np.array([names.index(r[0]) for r in rows], np.uint32)
    s, e = np.array([r[1] for r in rows], np.uint32), np.array([r[2] for r in rows], np.uint32)
    sc = np.array([0.0 if r[4] == "." else float(r[4]) for r in rows])
    valid = np.array([r[4] != "." for r in rows], np.uint8)
    mseq, ms, me, first, val, ok = orc.merge(seq, s, e, 0, sc, valid, op)
    out_path = tmp_path / "merged.bed"
    run("merge", "-b", path, "-f", op, "-o", out_path)
    lines = out_path.read_text().splitlines()
    assert len(lines) == len(ms)
    for g, line in enumerate(lines):
        f = line.split("\t")
        assert f[:3] == [names[mseq[g]], str(ms[g]), str(me[g])]
        if not ok[g]:
            assert f[3] == "."
        elif op in ("sum", "mean"):
            assert abs(float(f[3]) - val[g]) <= 1e-12 * max(1.0, abs(val[g])) * 50
        else:
            assert f[3] == rust_display(float(val[g]))
    r = run("merge", "-b", path, check=False)
    assert r.returncode == 1 and "func" in r.stderr


def test_cli_merge_bed5_extreme_scores(tmp_path):
    """Scores of very small / very large magnitude print as hundreds of characters in Rust's positional notation
    (f64::to_string never uses an exponent): the row buffer must hold them."""
    path = tmp_path / "x.bed"
    path.write_text("chr1\t0\t10\ta\t1e-100\nchr1\t5\t20\tb\t1e-100\nchr1\t100\t200\tc\t1e300\nchr2\t7\t9\td\t-2.5e-300\n")
    for op, want in (("sum", [2e-100, 1e300, -2.5e-300]), ("max", [1e-100, 1e300, -2.5e-300])):
        lines = run("merge", "-b", path, "-f", op).stdout.splitlines()
        assert [ln.split("\t")[:3] for ln in lines] == [["chr1", "0", "20"], ["chr1", "100", "200"], ["chr2", "7", "9"]]
        assert [ln.split("\t")[3] for ln in lines] == [rust_display(v) for v in want]
        assert len(lines[1].split("\t")[3]) == 301 and len(lines[2].split("\t")[3]) > 300


def test_cli_feature_density(tmp_path, orc):
    """Per window and feature: basepairs covered by the feature's bookend-merged ranges (src/commands.rs:796-858)."""
    rng = np.random.default_rng(21)
    genome = {"chr1": 50_000, "chr2": 20_000, "chr3": 7_000}
    (tmp_path / "genome.tsv").write_text("".join(f"{k}\t{v}\n" for k, v in genome.items()))
    feats = ["CDS", "exon", "five_prime_UTR"]
    rows = []
    for nm, L in genome.items():
        if nm == "chr3":
            continue
        k = 400
        st = np.sort(rng.integers(0, L - 300, k))
        for a in st:
            rows.append((nm, int(a), int(a + rng.integers(1, 300)), feats[int(rng.integers(0, 3))]))
    (tmp_path / "feat.bed").write_text("".join("\t".join(map(str, r)) + "\n" for r in rows))
    width, step = 1000, 400
    got = run("feature-density", "-g", tmp_path / "genome.tsv", "-b", tmp_path / "feat.bed", "-w", width, "-s", step, "--headers").stdout
    lines = got.splitlines()
    assert lines[0].split("\t") == ["chrom", "start", "end"] + sorted(feats)
    names = list(genome)
    wseq, ws, we = orc.windows(list(genome.values()), width, step, False)
    assert len(lines) - 1 == len(ws)
    cover = {}
    for nm, L in genome.items():
        for ft in feats:
            c = np.zeros(L + 400, bool)
            for r in rows:
                if r[0] == nm and r[3] == ft:
                    c[r[1]:r[2]] = True
            cover[(nm, ft)] = np.concatenate([[0], np.cumsum(c)])
    for i, line in enumerate(lines[1:]):
        f = line.split("\t")
        nm = names[wseq[i]]
        assert f[:3] == [nm, str(ws[i]), str(we[i])]
        for c, ft in enumerate(sorted(feats)):
            assert int(f[3 + c]) == cover[(nm, ft)][we[i]] - cover[(nm, ft)][ws[i]]
    # without --headers: same rows, no header line
    got2 = run("feature-density", "-g", tmp_path / "genome.tsv", "-b", tmp_path / "feat.bed", "-w", width, "-s", step).stdout
    assert got2.splitlines() == lines[1:]


def test_cli_gzip_input(tmp_path, orc):
    """gzip input is detected by its magic bytes (src/io/parsers/tsv.rs:27-45), also multi-member (bgzip-style) files:
    the output equals that of the plain files."""
    import gzip

    make_files(tmp_path, seed=12)
    plain = run("map", "-g", tmp_path / "genome.tsv", "-l", tmp_path / "left.bed", "-r", tmp_path / "right.bed", "-f", "mean,median,max").stdout
    with open(tmp_path / "left.bed", "rb") as f, gzip.open(tmp_path / "left.bed.gz", "wb") as g:
        g.write(f.read())
    data = (tmp_path / "right.bed").read_bytes()
    cut = data.index(b"\n", len(data) // 2) + 1
    with open(tmp_path / "right.bed.gz", "wb") as g:  # two gzip members back to back
        g.write(gzip.compress(data[:cut]))
        g.write(gzip.compress(data[cut:]))
    got = run("map", "-g", tmp_path / "genome.tsv", "-l", tmp_path / "left.bed.gz", "-r", tmp_path / "right.bed.gz", "-f", "mean,median,max").stdout
    assert got == plain and len(plain) > 1000
    merged_plain = run("merge", "-b", tmp_path / "right.bed", "-f", "sum").stdout
    assert run("merge", "-b", tmp_path / "right.bed.gz", "-f", "sum").stdout == merged_plain
    (tmp_path / "bad.bed.gz").write_bytes(gzip.compress(data)[:200])
    r = run("filter", "-g", tmp_path / "genome.tsv", "-l", tmp_path / "bad.bed.gz", "-r", tmp_path / "right.bed", check=False)
    assert r.returncode == 1 and "gzip" in r.stderr
