"""CPU model of the tile-local wavelet-tree median (k_sweep3's wavelet runs, csrc/sweep3.cuh) against a brute force.
Mirrors the kernel's data structures one to one: slice [wlo, whi) of the rights in start order, dense local ranks,
B sequence = C ++ end-order range ++ R, levelwise wavelet trees over permutations of 0..W-1, one rank query per level
and structure.  Run: python tools/wavelet_median_model.py"""
import numpy as np


def build_tree(seq, W, L):
    """bits[l][i] of the sequence sorted stably by the top l bits; returns list of bit arrays (as python lists of 0/1)."""
    levels = []
    cur = list(seq)
    for l in range(L):
        sh = L - 1 - l
        bits = [(k >> sh) & 1 for k in cur]
        levels.append(bits)
        nxt = [None] * W
        # dest = child start (arithmetic from the key) + earlier elements of the same child
        cnt = {}
        for k in cur:
            child = k >> sh
            start = min(child << sh, W)
            c = cnt.get(child, 0)
            nxt[start + c] = k
            cnt[child] = c + 1
        cur = nxt
    assert cur == sorted(seq)
    return levels


def rank0(bits, pos):
    return pos - sum(bits[:pos])


def kth(levA, levB, W, L, x, y, k):
    p, xa, xb = 0, x, y
    for l in range(L):
        sh = L - l
        s = p << sh
        z0 = p << (sh - 1)
        za = rank0(levA[l], s + xa) - z0
        zb = rank0(levB[l], s + xb) - z0
        if k < za - zb:
            p, xa, xb = 2 * p, za, zb
        else:
            k -= za - zb
            p, xa, xb = 2 * p + 1, xa - za, xb - zb
    return p


def run_case(rng, n, nl, maxlen, span):
    rs = np.sort(rng.integers(0, span, n))
    re = rs + rng.integers(1, maxlen, n)
    order = np.lexsort((re, rs))
    rs, re = rs[order], re[order]
    rank = rng.permutation(n)          # global ranks, all different
    ls = np.sort(rng.integers(0, span, nl))
    le = ls + rng.integers(1, maxlen, nl)
    eorder = np.lexsort((np.arange(n), re))   # perm_e: position in start order of the i-th smallest end
    ends_sorted = re[eorder]
    ub = np.searchsorted(rs, le, side="left")       # #{start < lend}
    pp = np.searchsorted(rs, ls, side="left")
    lb = np.searchsorted(ends_sorted, ls, side="right")  # #{end <= lstart}
    # one run = all the lefts
    pmin, umax = int(pp.min()), int(max(ub.max(), pp.max()))
    # fb: first position with end > starts[64 b]
    b = (pmin - 1) >> 6 if pmin else 0
    x0 = rs[64 * b] if pmin else None
    wlo = 0
    if pmin:
        cand = np.nonzero(re > x0)[0]
        wlo = int(cand[0]) if len(cand) and cand[0] < 64 * b else 64 * b
        wlo &= ~3
    whi = umax
    W = whi - wlo
    if W <= 0:
        return 0
    L = max(1, int(np.ceil(np.log2(W)))) if W > 1 else 1
    while (1 << L) < W:
        L += 1
    sl_rank = rank[wlo:whi]
    sl_ends = re[wlo:whi]
    sorted_r = np.sort(sl_rank)
    lrank = np.searchsorted(sorted_r, sl_rank)
    lsmin, lsmax = int(ls.min()), int(ls.max())
    lbmin, lbmax = int(lb.min()), int(lb.max())
    C = [int(lrank[j]) for j in range(W) if sl_ends[j] <= lsmin]
    R = [int(lrank[j]) for j in range(W) if sl_ends[j] > lsmax]
    mid = [int(lrank[eorder[q] - wlo]) for q in range(lbmin, lbmax)]
    assert all(0 <= eorder[q] - wlo < W for q in range(lbmin, lbmax))
    seqB = C + mid + R
    assert len(seqB) == W and sorted(seqB) == list(range(W)), (len(C), len(mid), len(R), W)
    assert wlo + len(C) == lbmin
    levA = build_tree([int(v) for v in lrank], W, L)
    levB = build_tree(seqB, W, L)
    checked = 0
    for i in range(nl):
        m = int(ub[i] - lb[i])
        mem = [int(rank[j]) for j in range(n) if rs[j] < le[i] and re[j] > ls[i]]
        assert len(mem) == m
        if m == 0:
            continue
        mem.sort()
        x = int(ub[i]) - wlo
        y = len(C) + int(lb[i]) - lbmin
        for k in {0, (m - 1) // 2, m // 2, m - 1}:
            p = kth(levA, levB, W, L, x, y, k)
            assert int(sorted_r[p]) == mem[k], (i, k, p)
            checked += 1
    return checked


if __name__ == "__main__":
    rng = np.random.default_rng(5)
    tot = 0
    for t in range(200):
        n = int(rng.integers(1, 400))
        nl = int(rng.integers(1, 60))
        tot += run_case(rng, n, nl, int(rng.integers(2, 300)), int(rng.integers(50, 3000)))
    print("ok", tot, "selections checked")
