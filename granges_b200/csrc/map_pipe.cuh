// map_pipe.cuh -- the pipelined single-reduction map kernel (headline path: `granges map --func mean`).
// Included by query.cu after the PTX / search helpers.
//
// Same arithmetic as k_tile<M_MAP, FAST>, organised as a persistent, warp-specialised TMA pipeline
// so that no warp ever waits on a dependent chain of global loads.  One CTA per SM, tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...; per tile (1024 consecutive lefts of one seqname):
//
//   warp 0  lefts loader   A: TMA the lefts (ls, le) of tile j into the lefts ring, as far ahead as
//                             the ring allows                                                 -> fullL
//   warp 1  surveyor       B: min / max of the tile's lefts out of shared memory -> coordinate bins
//                             -> four bin-table entries by plain loads, published one iteration
//                             later so the loads have a whole iteration to land               -> statsL
//   warp 2  window loader  C: TMA the slices of starts, ends_sorted, both bin tables and both
//                             prefix arrays that the tile can touch into the window ring (six
//                             bulk copies issued by six lanes at once, one mbarrier)          -> fullW
//   warps 3.. consumers:      wait fullW, resolve ub / lb of 2 lefts per thread out of shared
//                             memory, exact sums from the staged prefixes, write results,
//                             release the slots                                               -> emptyW, emptyL
//
// Tiles whose windows do not fit the ring slots (sparse lefts over dense rights) are resolved by
// per-left global searches inside the consumers; results are identical.

constexpr int PIPE_GROUPS = 1;  // consumer groups; group g takes tiles g, g + GROUPS, ... so one group computes while another waits
constexpr int PIPE_NCG = 16;    // warps per consumer group
constexpr int PIPE_NC = PIPE_GROUPS * PIPE_NCG;  // consumer warps
constexpr int PIPE_LPT = TILE / (PIPE_NCG * 32);  // consecutive lefts per consumer thread
constexpr int PIPE_NSURV = 4;  // surveyor warps (tile j goes to surveyor j % PIPE_NSURV)
constexpr int PIPE_THREADS = (PIPE_NC + 2 + PIPE_NSURV) * 32;
constexpr int PIPE_LSTAGES = 8;
constexpr int PIPE_WSTAGES = 4;
constexpr u32 PIPE_WIN = 1280;  // keys (and prefixes) per staged window
constexpr u32 PIPE_LUT = 640;   // bin-table entries per staged slice
constexpr int PIPE_MAXSEQ = 32; // seqnames per launch (their descriptors live in shared memory)

struct PipeSeq {
    const u32 *starts, *ends_sorted, *lut_a, *lut_b;
    const u64 *pfx_a, *pfx_b;
    u64 left_begin, left_end;
    double scale;
    u32 tile_begin, bmax, fast_cmax;
    int sh;
};

struct PipeTile {      // one per lefts-ring slot
    u64 g_first;       // global index of the tile's slot 0                (loader)
    u32 vlo, vhi;      // valid slots of the tile: [vlo, vhi)              (loader)
    int seq;           //                                                  (loader)
    u32 bin[4];        // bins of min l.end, max l.end, min l.start+1, max l.start+1   (surveyor)
    u32 bound[4];      // lut_a[bin0], lut_a[bin1 + 1], lut_b[bin2], lut_b[bin3 + 1]     (surveyor)
};

struct PipeDesc {      // one per window-ring slot (loader -> consumers)
    u64 g_first;
    u32 vlo, vhi;
    u32 a0, b0, la0, lb0;
    u32 staged;        // 1: windows staged in the slot; 0: search in global memory (windows too large, or invalid lefts)
    int seq;
};

struct PipeSmem {
    alignas(16) u32 ls[PIPE_LSTAGES][TILE];
    alignas(16) u32 le[PIPE_LSTAGES][TILE];
    alignas(16) u64 pa[PIPE_WSTAGES][PIPE_WIN + 8];
    alignas(16) u64 pb[PIPE_WSTAGES][PIPE_WIN + 8];
    alignas(16) u32 ka[PIPE_WSTAGES][PIPE_WIN + 8];
    alignas(16) u32 kb[PIPE_WSTAGES][PIPE_WIN + 8];
    alignas(16) u32 la[PIPE_WSTAGES][PIPE_LUT + 8];
    alignas(16) u32 lb[PIPE_WSTAGES][PIPE_LUT + 8];
    PipeSeq seq[PIPE_MAXSEQ];
    PipeTile tile[PIPE_LSTAGES];
    PipeDesc desc[PIPE_WSTAGES];
    alignas(8) u64 fullL[PIPE_LSTAGES];
    alignas(8) u64 emptyL[PIPE_LSTAGES];
    alignas(8) u64 statsL[PIPE_LSTAGES];
    alignas(8) u64 fullW[PIPE_WSTAGES];
    alignas(8) u64 emptyW[PIPE_WSTAGES];
};

__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

constexpr int PIPE_MULTI = 100;  // FAST value: several prefix-sum reductions (sum / sum-not-empty / mean / count) per left

constexpr int PIPE_UBLB = 102;   // FAST value: PIPE_MULTI for the columns the prefix sums answer (others are left alone) + the three
                                 // counts ub = #{start < l.end}, pp = #{start < l.start}, lb = #{end <= l.start} written out for the
                                 // member sweep / the closed forms that follow

struct PipeOps {
    int k;
    int ops[GRB_MAX_OPS];
    u32 *ub, *pp, *lb;  // PIPE_UBLB
    int need_pfx;       // PIPE_UBLB: some column needs the score prefixes
};

// flags: bit 0 = stage with TMA; bit 1 = debug, consumers skip the arithmetic (pipeline-only timing)
template <int FAST>
__global__ void __launch_bounds__(PIPE_THREADS, 1) k_map_pipe(const SeqDev *__restrict__ seqs, int nseq, u32 tile_base, u32 ntiles,
                                                              const u32 *__restrict__ ls, const u32 *__restrict__ le,
                                                              double *__restrict__ out, u8 *__restrict__ out_valid, int flags,
                                                              u32 *__restrict__ ctx_flags, const PipeOps mops) {
    const bool want_prefixes = FAST != GRB_OP_COUNT && (FAST != PIPE_UBLB || mops.need_pfx != 0);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeSmem &S = *reinterpret_cast<PipeSmem *>(smem_raw);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool use_tma = flags & 1;
    if (tid == 0) {
        for (int i = 0; i < PIPE_LSTAGES; i++) {
            mbar_init(&S.fullL[i], 1);
            mbar_init(&S.emptyL[i], PIPE_NCG);
            mbar_init(&S.statsL[i], 1);
        }
        for (int i = 0; i < PIPE_WSTAGES; i++) {
            mbar_init(&S.fullW[i], 1);
            mbar_init(&S.emptyW[i], PIPE_NCG);
        }
    }
    if ((int)tid < nseq) {
        const SeqDev *sd = &seqs[tid];
        PipeSeq &q = S.seq[tid];
        q.starts = sd->starts;
        q.ends_sorted = sd->ends_sorted;
        q.lut_a = sd->lut_a;
        q.lut_b = sd->lut_b;
        q.pfx_a = sd->pfx_a;
        q.pfx_b = sd->pfx_b;
        q.left_begin = sd->left_begin;
        q.left_end = sd->left_end;
        q.scale = sd->scale;
        q.tile_begin = sd->tile_begin;
        q.bmax = sd->lut_bmax;
        q.fast_cmax = sd->fast_cmax;
        q.sh = sd->lut_shift;
    }
    __syncthreads();
    const u32 nmine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) {
        // ================= lefts loader (A) =================
        const u64 pol = l2_evict_first_policy();
        const bool vec_in = ((((uintptr_t)ls) | ((uintptr_t)le)) & 15) == 0;
        int sq = 0;  // tiles only move forward, so does the seqname
        for (u32 j = 0; j < nmine; j++) {
            const u32 sl = j % PIPE_LSTAGES;
            const u32 t = tile_base + blockIdx.x + j * gridDim.x;
            while (sq + 1 < nseq && S.seq[sq + 1].tile_begin <= t) sq++;
            const PipeSeq &q = S.seq[sq];
            const u64 g_first = ((q.left_begin >> TILE_SHIFT) + (t - q.tile_begin)) << TILE_SHIFT;
            const u64 lo_g = q.left_begin > g_first ? q.left_begin : g_first;
            const u64 hi_g = q.left_end < g_first + TILE ? q.left_end : g_first + TILE;
            const bool whole = vec_in && lo_g == g_first && hi_g == g_first + TILE;
            mbar_wait(&S.emptyL[sl], ((j / PIPE_LSTAGES) & 1) ^ 1);
            if (lane == 0) {
                PipeTile &ti = S.tile[sl];
                ti.g_first = g_first;
                ti.vlo = (u32)(lo_g - g_first);
                ti.vhi = (u32)(hi_g - g_first);
                ti.seq = sq;
            }
            if (whole && use_tma) {
                if (lane == 0) mbar_arrive_expect_tx(&S.fullL[sl], 2 * TILE * 4);
                __syncwarp();
                if (lane < 2) tma_load_1d_hint(lane ? S.le[sl] : S.ls[sl], (lane ? le : ls) + g_first, TILE * 4, &S.fullL[sl], pol);
            } else {
                // ragged tile (seqname boundary, end of the arrays, unaligned pointers): plain copies
                for (u32 i = lane; i < TILE; i += 32) {
                    const u64 g = g_first + i;
                    const bool ok = g >= lo_g && g < hi_g;
                    S.ls[sl][i] = ok ? ls[g] : 0u;
                    S.le[sl][i] = ok ? le[g] : 0u;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.fullL[sl]);
            }
            __syncwarp();
        }
        return;
    }

    if (warp == 2) {
        // ================= window loader (C) =================
        const u64 pol = l2_evict_first_policy();
        for (u32 j = 0; j < nmine; j++) {
            const u32 sw = j % PIPE_WSTAGES, sl = j % PIPE_LSTAGES;
            mbar_wait(&S.statsL[sl], (j / PIPE_LSTAGES) & 1);
            mbar_wait(&S.emptyW[sw], ((j / PIPE_WSTAGES) & 1) ^ 1);
            const PipeTile &ti = S.tile[sl];
            const PipeSeq &q = S.seq[ti.seq];
            const u32 a0 = ti.bound[0] & ~3u, b0 = ti.bound[2] & ~3u, la0 = ti.bin[0] & ~3u, lb0 = ti.bin[2] & ~3u;
            // + 8: a search reads the two aligned quads around its answer, i.e. up to 7 keys past it
            const u32 alen = ti.bound[1] - a0 + 8, blen = ti.bound[3] - b0 + 8;
            const u32 lalen = ti.bin[1] + 1 - la0, lblen = ti.bin[3] + 1 - lb0;
            // bound[1] == 0xffffffff: the surveyor found invalid lefts in this tile (no index is that long)
            const bool staged = use_tma && ti.bound[1] != 0xffffffffu && alen <= PIPE_WIN && blen <= PIPE_WIN && lalen <= PIPE_LUT &&
                                lblen <= PIPE_LUT;
            const bool want_pfx = want_prefixes;
            // lane l issues copy l: 0 lut_a, 1 lut_b, 2 starts, 3 ends_sorted, 4 pfx_a, 5 pfx_b
            const void *src = nullptr;
            void *dst = nullptr;
            u32 bytes = 0;
            switch (lane) {
                case 0: src = q.lut_a + la0; dst = S.la[sw]; bytes = lalen * 4; break;
                case 1: src = q.lut_b + lb0; dst = S.lb[sw]; bytes = lblen * 4; break;
                case 2: src = q.starts + a0; dst = S.ka[sw]; bytes = alen * 4; break;
                case 3: src = q.ends_sorted + b0; dst = S.kb[sw]; bytes = blen * 4; break;
                case 4: src = q.pfx_a + a0; dst = S.pa[sw]; bytes = want_pfx ? alen * 8 : 0; break;
                case 5: src = q.pfx_b + b0; dst = S.pb[sw]; bytes = want_pfx ? blen * 8 : 0; break;
                default: break;
            }
            bytes = (bytes + 15) & ~15u;
            const u32 total_bytes = warp_sum(bytes);
            if (lane == 0) {
                PipeDesc &d = S.desc[sw];
                d.g_first = ti.g_first;
                d.vlo = ti.vlo;
                d.vhi = ti.vhi;
                d.seq = ti.seq;
                d.a0 = a0;
                d.b0 = b0;
                d.la0 = la0;
                d.lb0 = lb0;
                d.staged = staged ? 1u : 0u;
                if (staged)
                    mbar_arrive_expect_tx(&S.fullW[sw], total_bytes);
                else
                    mbar_arrive(&S.fullW[sw]);
            }
            __syncwarp();
            if (staged && bytes) tma_load_1d_hint(dst, src, bytes, &S.fullW[sw], pol);
            __syncwarp();
        }
        return;
    }

    if (warp == 1 || (warp >= 3 && warp < 2 + PIPE_NSURV)) {
        // ================= surveyors =================
        const u32 sid = warp == 1 ? 0u : warp - 2;
        for (u32 j = sid; j < nmine; j += PIPE_NSURV) {
            u32 bin_new = 0, bound_new = 0;
            if (j < nmine) {
                const u32 sl = j % PIPE_LSTAGES;
                mbar_wait(&S.fullL[sl], (j / PIPE_LSTAGES) & 1);
                const PipeTile &ti = S.tile[sl];
                const u32 vlo = ti.vlo, vhi = ti.vhi;
                const PipeSeq &q = S.seq[ti.seq];
                u32 mn_s = 0xffffffffu, mx_s = 0, mn_e = 0xffffffffu, mx_e = 0;
                // widest (end - start - 1) mod 2^32: it reaches 2^31-1 iff some left has end <= start (ends <= 2^31-1)
                u32 mx_w = 0;
                if (vlo == 0 && vhi == TILE) {
                    const uint4 *vs = reinterpret_cast<const uint4 *>(S.ls[sl]);
                    const uint4 *ve = reinterpret_cast<const uint4 *>(S.le[sl]);
#pragma unroll
                    for (int r = 0; r < TILE / 128; r++) {
                        const uint4 x = vs[r * 32 + lane], y = ve[r * 32 + lane];
                        mx_w = max(mx_w, max(max(y.x - x.x - 1, y.y - x.y - 1), max(y.z - x.z - 1, y.w - x.w - 1)));
                        mn_s = min(mn_s, min(min(x.x, x.y), min(x.z, x.w)));
                        mx_s = max(mx_s, max(max(x.x, x.y), max(x.z, x.w)));
                        mn_e = min(mn_e, min(min(y.x, y.y), min(y.z, y.w)));
                        mx_e = max(mx_e, max(max(y.x, y.y), max(y.z, y.w)));
                    }
                } else {
                    for (u32 i = vlo + lane; i < vhi; i += 32) {
                        const u32 x = S.ls[sl][i], y = S.le[sl][i];
                        mx_w = max(mx_w, y - x - 1);
                        mn_s = min(mn_s, x);
                        mx_s = max(mx_s, x);
                        mn_e = min(mn_e, y);
                        mx_e = max(mx_e, y);
                    }
                }
                mn_e = warp_min(mn_e);
                mx_e = warp_max(mx_e);
                mn_s = warp_min(mn_s);
                mx_s = warp_max(mx_s);
                // invalid lefts cannot exist in the reference (it panics: ranges/mod.rs:30, coitrees.rs:53-54): make the call fail
                mx_w = warp_max(mx_w);
                const bool tile_bad = mx_w >= 0x7fffffffu || mx_e > 0x7fffffffu;
                if (tile_bad && lane == 0) atomicOr(ctx_flags, 1u);
                const u32 qq = lane & 3;
                // (PIPE_UBLB also searches the starts for l.start: the low end of that window comes from min l.start)
                const u32 v = qq == 0 ? (FAST == PIPE_UBLB ? mn_s : mn_e) : qq == 1 ? mx_e : qq == 2 ? mn_s : mx_s;
                const u32 x = v + (qq >> 1);  // l.end for the starts table, l.start + 1 for the ends table
                bin_new = min(x >> q.sh, q.bmax);
                bound_new = ((qq < 2) ? q.lut_a : q.lut_b)[bin_new + (qq & 1)];
                // publish at once: this warp now waits for its four table entries, the other surveyors carry on
                // a tile with invalid lefts must not be staged (their bins may fall outside the slices): an impossible
                // upper bound tells the window loader, and the consumers' global path skips those lefts
                if (tile_bad && qq == 1) bound_new = 0xffffffffu;
                if (lane < 4) {
                    S.tile[sl].bin[lane] = bin_new;
                    S.tile[sl].bound[lane] = bound_new;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.statsL[sl]);
            }
        }
        return;
    }

    // ================= consumers =================
    const u32 cgroup = (warp - 2 - PIPE_NSURV) / PIPE_NCG, cw = (warp - 2 - PIPE_NSURV) % PIPE_NCG;
    const u32 i0 = (cw * 32 + lane) * PIPE_LPT;  // first slot of this thread inside a tile
    const bool ovec = ((((uintptr_t)out) & 15) | (((uintptr_t)out_valid) & 3)) == 0;
    const bool skip = flags & 2;
    const bool nostore = flags & 4;  // debug: do the arithmetic, keep the stores out (timing experiments)
    for (u32 j = cgroup; j < nmine; j += PIPE_GROUPS) {
        const u32 sw = j % PIPE_WSTAGES, sl = j % PIPE_LSTAGES;
        mbar_wait(&S.fullW[sw], (j / PIPE_WSTAGES) & 1);
        if (!skip) {
            const PipeDesc &d = S.desc[sw];
            const PipeSeq &q = S.seq[d.seq];
            const u32 vlo = d.vlo, vhi = d.vhi;
            const int sh = q.sh;
            const u32 bmax = q.bmax, fast_cmax = q.fast_cmax;
            const double scale = q.scale;
            u32 a[PIPE_LPT], b[PIPE_LPT];
            if (PIPE_LPT == 4) {
                const uint4 va = *reinterpret_cast<const uint4 *>(&S.ls[sl][i0]);
                const uint4 vb = *reinterpret_cast<const uint4 *>(&S.le[sl][i0]);
                a[0] = va.x, a[1 % PIPE_LPT] = va.y, a[2 % PIPE_LPT] = va.z, a[3 % PIPE_LPT] = va.w;
                b[0] = vb.x, b[1 % PIPE_LPT] = vb.y, b[2 % PIPE_LPT] = vb.z, b[3 % PIPE_LPT] = vb.w;
            } else if (PIPE_LPT == 2) {
                const uint2 va = *reinterpret_cast<const uint2 *>(&S.ls[sl][i0]);
                const uint2 vb = *reinterpret_cast<const uint2 *>(&S.le[sl][i0]);
                a[0] = va.x, a[1 % PIPE_LPT] = va.y;
                b[0] = vb.x, b[1 % PIPE_LPT] = vb.y;
            } else {
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) a[k] = S.ls[sl][i0 + k], b[k] = S.le[sl][i0 + k];
            }
            bool okv[PIPE_LPT];
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++) {
                okv[k] = i0 + k >= vlo && i0 + k < vhi;
            }
            // (A left with end <= start or end > 2^31-1 -- impossible in the reference, which panics -- was flagged by
            // the surveyor, which also keeps its tile off the staged path.)
            const u32 a0 = d.a0, b0 = d.b0, la0 = d.la0, lb0 = d.lb0;
            const SeqDev *__restrict__ sd = &seqs[d.seq];  // only dereferenced on the rare paths
            u32 ub[PIPE_LPT], lbk[PIPE_LPT], ppk[PIPE_LPT];  // absolute positions
            u64 pa[PIPE_LPT], pb[PIPE_LPT];
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++) ppk[k] = 0;
            if (d.staged) {
                const u32 *__restrict__ KA = S.ka[sw];
                const u32 *__restrict__ KB = S.kb[sw];
                u32 pA[PIPE_LPT], pB[PIPE_LPT], cA[PIPE_LPT], cB[PIPE_LPT];
                // bin -> first position of the bin (relative to the staged slice); invalid slots scan with x = 0
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) {
                    const u32 ba = min(b[k] >> sh, bmax), bb = min((a[k] + 1) >> sh, bmax);
                    pA[k] = okv[k] ? S.la[sw][ba - la0] - a0 : 0u;
                    pB[k] = okv[k] ? S.lb[sw][bb - lb0] - b0 : 0u;
                }
                // one unconditional round per search: the 8 keys of the two aligned quads around the bin's first
                // position, two 128-bit loads.  Keys before that position are < x by construction, so the answer
                // is (quad base) + #(keys < x) unless all 8 are smaller (a crowded bin: finish with the scan).
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) {
                    const u32 xa = okv[k] ? b[k] : 0u, xb = okv[k] ? a[k] + 1 : 0u;
                    const u32 qa = pA[k] & ~3u, qb = pB[k] & ~3u;
                    const uint4 a0v = *reinterpret_cast<const uint4 *>(KA + qa), a1v = *reinterpret_cast<const uint4 *>(KA + qa + 4);
                    const uint4 b0v = *reinterpret_cast<const uint4 *>(KB + qb), b1v = *reinterpret_cast<const uint4 *>(KB + qb + 4);
                    cA[k] = (a0v.x < xa) + (a0v.y < xa) + (a0v.z < xa) + (a0v.w < xa) + (a1v.x < xa) + (a1v.y < xa) + (a1v.z < xa) +
                            (a1v.w < xa);
                    cB[k] = (b0v.x < xb) + (b0v.y < xb) + (b0v.z < xb) + (b0v.w < xb) + (b1v.x < xb) + (b1v.y < xb) + (b1v.z < xb) +
                            (b1v.w < xb);
                    pA[k] = okv[k] ? qa + cA[k] : 0u;
                    pB[k] = okv[k] ? qb + cB[k] : 0u;
                }
                if (FAST == PIPE_UBLB) {
                    // third search: #{start < l.start}, same scheme on the starts window
#pragma unroll
                    for (int k = 0; k < PIPE_LPT; k++) {
                        const u32 xp = okv[k] ? a[k] : 0u;
                        u32 pP = okv[k] ? S.la[sw][min(a[k] >> sh, bmax) - la0] - a0 : 0u;
                        const u32 qp = pP & ~3u;
                        const uint4 p0v = *reinterpret_cast<const uint4 *>(KA + qp), p1v = *reinterpret_cast<const uint4 *>(KA + qp + 4);
                        const u32 cP = (p0v.x < xp) + (p0v.y < xp) + (p0v.z < xp) + (p0v.w < xp) + (p1v.x < xp) + (p1v.y < xp) + (p1v.z < xp) +
                                       (p1v.w < xp);
                        pP = okv[k] ? qp + cP : 0u;
                        if (cP == 8) pP = scan_lower_bound(KA, pP, a[k]);
                        ppk[k] = a0 + pP;
                    }
                }
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) {
                    if (cA[k] == 8) pA[k] = scan_lower_bound(KA, pA[k], b[k]);      // rare: a crowded bin
                    if (cB[k] == 8) pB[k] = scan_lower_bound(KB, pB[k], a[k] + 1);
                    pa[k] = want_prefixes ? S.pa[sw][pA[k]] : 0ull;
                    pb[k] = want_prefixes ? S.pb[sw][pB[k]] : 0ull;
                    ub[k] = a0 + pA[k];
                    lbk[k] = b0 + pB[k];
                }
            } else {
                // the tile's windows do not fit the ring slot: per-left searches in global memory
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) {
                    ub[k] = lbk[k] = 0;
                    pa[k] = pb[k] = 0;
                    if (!okv[k]) continue;
                    if (b[k] <= a[k] || b[k] > 0x7fffffffu) continue;  // invalid left (flagged by the surveyor): an empty group
                    const u32 xa = b[k], xb = a[k] + 1;
                    ub[k] = scan_lower_bound(q.starts, q.lut_a[min(xa >> sh, bmax)], xa);
                    lbk[k] = scan_lower_bound(q.ends_sorted, q.lut_b[min(xb >> sh, bmax)], xb);
                    if (want_prefixes) {
                        pa[k] = q.pfx_a[ub[k]];
                        pb[k] = q.pfx_b[lbk[k]];
                    }
                    if (FAST == PIPE_UBLB) ppk[k] = scan_lower_bound(q.starts, q.lut_a[min(a[k] >> sh, bmax)], a[k]);
                }
            }
            // ---- everything this warp needs from the two ring slots is in registers now: hand them back
            // before the arithmetic and the stores, so the loaders can refill them meanwhile
            const u64 g0 = d.g_first + i0;
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.emptyW[sw]);
                mbar_arrive(&S.emptyL[sl]);
            }
            // ---- reductions
            if (FAST == PIPE_UBLB) {
                if (PIPE_LPT == 2 && okv[0] && okv[1]) {
                    *reinterpret_cast<uint2 *>(mops.ub + g0) = make_uint2(ub[0], ub[1 % PIPE_LPT]);
                    *reinterpret_cast<uint2 *>(mops.pp + g0) = make_uint2(ppk[0], ppk[1 % PIPE_LPT]);
                    *reinterpret_cast<uint2 *>(mops.lb + g0) = make_uint2(lbk[0], lbk[1 % PIPE_LPT]);
                } else {
#pragma unroll
                    for (int k = 0; k < PIPE_LPT; k++)
                        if (okv[k]) {
                            mops.ub[g0 + k] = ub[k];
                            mops.pp[g0 + k] = ppk[k];
                            mops.lb[g0 + k] = lbk[k];
                        }
                }
            }
            if (FAST == PIPE_MULTI || FAST == PIPE_UBLB) {
                // several reductions from the same (count, exact sum): row-major out[left][column]
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) {
                    if (!okv[k]) continue;
                    const u32 c = ub[k] - lbk[k];
                    double sum = 0.0;
                    if (!want_prefixes) {
                        // no column needs a sum
                    } else if (c < fast_cmax) {
                        sum = (double)(i64)(pa[k] - pb[k]) * scale;
                    } else {
                        const i128 ba128 = sd->base_a[ub[k] >> GRB_BASE_SHIFT], bb128 = sd->base_b[lbk[k] >> GRB_BASE_SHIFT];
                        const i128 PA = ba128 + (i128)(i64)(pa[k] - (u64)(u128)ba128);
                        const i128 PB = bb128 + (i128)(i64)(pb[k] - (u64)(u128)bb128);
                        sum = fx_to_double(PA - PB, sd->e0);
                    }
                    const double mean = c ? sum / (double)c : 0.0;
                    double *po = out + (g0 + k) * (u64)mops.k;
                    u8 *pv = out_valid + (g0 + k) * (u64)mops.k;
                    for (int col = 0; col < mops.k; col++) {
                        const int op = mops.ops[col];
                        if (FAST == PIPE_UBLB && op != GRB_OP_COUNT && op != GRB_OP_MEAN && op != GRB_OP_SUM && op != GRB_OP_SUM_NOT_EMPTY)
                            continue;  // a member reduction / closed form: the kernels that follow own this column
                        const double rr = op == GRB_OP_COUNT ? (double)c : op == GRB_OP_MEAN ? mean : sum;
                        const u8 vv = (op == GRB_OP_SUM || op == GRB_OP_COUNT) ? (u8)1 : (u8)(c != 0);
                        __stcs(po + col, rr);
                        pv[col] = vv;
                    }
                }
                continue;
            }
            double r[PIPE_LPT];
            u8 v[PIPE_LPT];
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++) {
                const u32 c = ub[k] - lbk[k];
                if (FAST == GRB_OP_COUNT) {
                    r[k] = (double)c;
                    v[k] = 1;
                    continue;
                }
                double sum;
                if (c < fast_cmax) {
                    sum = (double)(i64)(pa[k] - pb[k]) * scale;
                } else {
                    const i128 ba128 = sd->base_a[ub[k] >> GRB_BASE_SHIFT], bb128 = sd->base_b[lbk[k] >> GRB_BASE_SHIFT];
                    const i128 PA = ba128 + (i128)(i64)(pa[k] - (u64)(u128)ba128);
                    const i128 PB = bb128 + (i128)(i64)(pb[k] - (u64)(u128)bb128);
                    sum = fx_to_double(PA - PB, sd->e0);
                }
                if (FAST == GRB_OP_SUM) {
                    r[k] = sum;
                    v[k] = 1;
                } else if (FAST == GRB_OP_SUM_NOT_EMPTY) {
                    r[k] = sum;
                    v[k] = c != 0;
                } else {
                    v[k] = c != 0;
                    r[k] = c ? sum / (double)c : 0.0;
                }
            }
            if (nostore && r[0] != 1.2345e300) {
                // nothing
            } else if (ovec && okv[0] && okv[PIPE_LPT - 1] && (PIPE_LPT == 4 || PIPE_LPT == 2)) {
                double2 *po = reinterpret_cast<double2 *>(out + g0);
                __stcs(po, make_double2(r[0], r[1 % PIPE_LPT]));  // streaming stores: results are not read back here
                if (PIPE_LPT == 4) {
                    __stcs(po + 1, make_double2(r[2 % PIPE_LPT], r[3 % PIPE_LPT]));
                    __stcs(reinterpret_cast<uchar4 *>(out_valid + g0), make_uchar4(v[0], v[1 % PIPE_LPT], v[2 % PIPE_LPT], v[3 % PIPE_LPT]));
                } else {
                    __stcs(reinterpret_cast<uchar2 *>(out_valid + g0), make_uchar2(v[0], v[1 % PIPE_LPT]));
                }
            } else {
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) {
                    if (okv[k]) {
                        out[g0 + k] = r[k];
                        out_valid[g0 + k] = v[k];
                    }
                }
            }
        } else {
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.emptyW[sw]);
                mbar_arrive(&S.emptyL[sl]);
            }
        }
    }
}
