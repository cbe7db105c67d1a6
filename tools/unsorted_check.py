import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
rng = np.random.default_rng(99)
nl, nr, span = 1_300_000, 400_000, 30_000_000
cases = []
for s in range(2):
    rs = rng.integers(0, span - 1000, nr).astype(np.uint32)
    re = (rs + rng.integers(1, 1000, nr)).astype(np.uint32)
    ls = np.sort(rng.integers(0, span - 1000, nl)).astype(np.uint32)
    le = (ls + rng.integers(1, 1000, nl)).astype(np.uint32)
    cases.append(dict(rs=rs, re=re, sc=rng.random(nr), ls=ls, le=le, perm=rng.permutation(nl)))
off = np.array([0, nl, 2 * nl], np.uint64)
ls_sorted = np.concatenate([c["ls"] for c in cases]); le_sorted = np.concatenate([c["le"] for c in cases])
ls_shuf = np.concatenate([c["ls"][c["perm"]] for c in cases]); le_shuf = np.concatenate([c["le"][c["perm"]] for c in cases])
perm_all = np.concatenate([c["perm"] + i * nl for i, c in enumerate(cases)])
res = {}
for mode in ("1", "0"):
    os.environ["GRB_PIPE_MODE"] = mode
    import granges_b200 as gb
    ctx = gb.Context(0)
    idxs = [ctx.index_build(c["rs"], c["re"]).set_scores(c["sc"]) for c in cases]
    res["shuf" + mode] = ctx.map_joins_batch(idxs, off, ls_shuf, le_shuf, ["mean"])
    res["sort" + mode] = ctx.map_joins_batch(idxs, off, ls_sorted, le_sorted, ["mean"])
ref = res["sort0"]
for k, (o, v) in res.items():
    if k.startswith("shuf"):
        ro, rv = ref[0][perm_all], ref[1][perm_all]
    else:
        ro, rv = ref
    m = rv == 1
    print(k, "valid equal", bool((v == rv).all()), "values equal", bool((o[m] == ro[m]).all()), "max", float(o[v == 1].max()))
