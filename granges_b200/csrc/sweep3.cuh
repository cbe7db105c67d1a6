// sweep3.cuh -- MIN / MAX / MEDIAN over score *ranks* (src/data/operations.rs:17-32,71-98): per run of lefts either wavelet
// trees over the staged slice (the median as an order statistic of a prefix difference: "wavelet runs" below, round 2) or the
// staged member sweep (round 1; sparse lefts over dense rights, unsorted lefts).  Included by query.cu.
//
// The member sweep:
// Same group structure as sweep.cuh (k_tile<M_UBLB> has resolved ub, pp, lb per left; the members are
// [pp, ub) plus pp - lb straddlers below pp), but
//   * a CTA stages the slice of `ends`, `rank_a` and `bme` its lefts can touch into shared memory with
//     bulk copies: consecutive (sorted) lefts share almost all of it, and no lane ever waits on a chain of
//     dependent global loads.  The slice starts at fb[..] (index.cu: k_first_back), the precomputed bound on
//     how far back a straddler can sit, and ends at the largest ub.  A run of lefts whose slice does not
//     fit is halved until it does; below 32 lefts the groups walk global memory instead (same code);
//   * the members' scores are replaced by their ranks (32-bit, all different): MIN / MAX are integer
//     min / max, the median is a sorting network over 32-bit keys, and only the answers are translated back
//     through score_sorted[] -- by loads that are issued one left ahead of the stores that need them.
// Used when every index of the batch has its member tables, there are at most 8 reductions and no non-finite score (MIN / MAX skip
// those: src/data/operations.rs:71-86); k_sweep2 keeps the overlap-width reductions and the rest.

constexpr int SW3_TPB = 256;
constexpr u32 SW3_WCAP = 4096;      // rights per staged slice
constexpr u32 SW3_KBUF = 4096;      // rank slots per CTA for the medians: 32 groups x 128, 16 x 256 or 8 x 512
constexpr u32 SW3_MIN_RUN = 32;     // runs shorter than this are not halved any further

// ---- wavelet runs: the median as an order statistic of a prefix difference -------------------------------------
// A run of lefts whose slice of rights S = start-order positions [wlo, whi) fits WV_CAP builds, once for the run,
//   * the dense local ranks of S (bitonic sort of the 32-bit score ranks in registers / shuffles / shared memory,
//     then one binary search per right): keys 0 .. W-1, all different, sorted[] maps a key back to the score rank;
//   * two levelwise wavelet trees over permutations of those keys: A = S in start order, B = S as
//     C ++ end-order range [min lb, max lb) ++ R, C = the rights of S that end at or before every left start of the
//     run, R = those that end after all of them (neither can separate two lefts of the run, so any order will do).
// A left's members are then the first ub - wlo elements of A minus the first |C| + lb - min lb elements of B (the
// counting identity of DESIGN.md section 3 inside the slice: everything below wlo ends before every left starts), and
// its k-th smallest member is one descent: per level one rank query in each tree (the keys are dense, so node
// boundaries are arithmetic), whatever the group's size.  MIN / MAX beside a median are the descents k = 0 and m - 1.
// Cost per run: ~45 warp-instructions per right of the slice + ~8 per left and descent, against ~270 per left for the
// member sweep at 43 members per left (growing like m log^2 m): taken when the run has at least W / wave_ratio lefts.
constexpr u32 WV_CAP = 2048;           // rights per wavelet slice
constexpr int WV_LMAX = 11;            // levels: keys below 2^11
constexpr u32 WV_ROW = WV_CAP / 32 + 1;  // words per level (+1: a query may ask for the position one past the end)

struct WaveSmem {
    union {
        struct {
            alignas(16) u32 rank[WV_CAP];        // staged score ranks (None -> 0x80000000 | local position)
            alignas(16) u32 perm[WV_CAP + 8];    // staged perm_e[min lb .. max lb)
        };
        uint2 tree[2][WV_LMAX][WV_ROW];          // then, per tree / level / 32 positions: x = bitmap (set bit = key bit is one),
    };                                           // y = zeros of the level before the word
    union {
        alignas(16) u32 ends[WV_CAP];            // staged ends ...
        u16 seq1[2][WV_CAP];                     // ... then the second halves of the two ping-pong sequences
    };
    alignas(16) u32 sorted[WV_CAP];              // the slice's ranks in ascending order: key -> score rank
    u16 seq0[2][WV_CAP];                         // A and B at the current level (start: local rank of each position)
    u32 wtot[2][SW3_TPB / 32];                   // zeros per warp of the level being built
    u32 cnt[2];                                  // cursors of the C / R fill
};

struct Sweep3Smem {
    union {
        struct {
            alignas(16) u32 ends[SW3_WCAP];
            alignas(16) u32 rank[SW3_WCAP];
            alignas(16) u32 bme[SW3_WCAP / 64 + 8];
            u32 kbuf[SW3_KBUF];
        };
        WaveSmem w;
    };
    alignas(8) u64 bar;
    u32 red[7][SW3_TPB / 32];
    u32 bcast[8];
};

// A prefix-sum column (count / sum / sum-not-empty / mean) from what the pipelined pass left per left: exactly its own
// arithmetic (pipe.cu, PIPE_MULTI epilogue), so the column is the same whoever writes it.
__device__ __forceinline__ bool sw3_is_pipe_op(int op) {
    return op == GRB_OP_COUNT || op == GRB_OP_SUM || op == GRB_OP_SUM_NOT_EMPTY || op == GRB_OP_MEAN;
}
__device__ __forceinline__ double sw3_pipe_col(int op, const uint4 pl, u8 &valid) {
    const double sum = __hiloint2double((int)pl.y, (int)pl.x);
    const u32 cnt = pl.z, nval = pl.w;
    valid = (op == GRB_OP_SUM || op == GRB_OP_COUNT) ? (u8)1 : (u8)(nval != 0);
    if (op == GRB_OP_COUNT) return (double)cnt;
    if (op == GRB_OP_MEAN) return nval ? sum / (double)nval : 0.0;
    return sum;
}

// Median of m <= G * R different 32-bit keys held in kbuf[0, m) by a group of G lanes: a bitonic sorting network
// over R registers per lane (element position = lane * R + register: small partner distances stay inside a
// lane, the large ones are one shuffle per register), padded with low / high sentinels so that the middle pair
// lands on positions N/2 - 1 and N/2; the last merge stops after its first step, where the lower half of the
// lanes holds the N/2 smallest keys.  No data-dependent branch: the four groups of a warp stay in lockstep.
// Keys are shifted by one so that 0 and 0xffffffff are free for the sentinels.  Returns key + 1 of the lower
// and the upper middle (equal ranks are impossible; for odd m only `lo` is meaningful).
template <int G, int R>
__device__ __forceinline__ void bitonic_median(const u32 *kbuf, u32 m, u32 glane, u32 &lo, u32 &hi) {
    constexpr u32 gmask = 0xffffffffu;  // the whole warp executes this together; the xor distances stay inside a group
    constexpr int N = G * R;
    const u32 L = (m & 1) ? (u32)(N / 2 - 1) - ((m - 1) >> 1) : (u32)(N / 2) - (m >> 1);  // low sentinels
    u32 v[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const u32 i = (u32)r * G + glane;  // any assignment of the elements to the slots will do: coalesced reads
        v[r] = i < m ? kbuf[i] + 1u : (i < m + L ? 0u : 0xffffffffu);
    }
    // sort inside the lane (ascending): bitonic network over the registers
#pragma unroll
    for (int k = 2; k <= R; k <<= 1) {
#pragma unroll
        for (int r = 0; r < R; r++) {  // flip step: r <-> r ^ (k - 1)
            const int q = r ^ (k - 1);
            if (q > r) {
                const u32 a = v[r], b = v[q];
                v[r] = min(a, b);
                v[q] = max(a, b);
            }
        }
#pragma unroll
        for (int j = k >> 2; j >= 1; j >>= 1) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int q = r ^ j;
                if (q > r) {
                    const u32 a = v[r], b = v[q];
                    v[r] = min(a, b);
                    v[q] = max(a, b);
                }
            }
        }
    }
    // merges across lanes: blocks of k = 2R, 4R, ... elements
#pragma unroll
    for (int k = 2 * R; k <= N; k <<= 1) {
        {
            // flip step: position e <-> e ^ (k - 1): lane ^ (k / R - 1), register R - 1 - r
            const int lm = k / R - 1;
            const bool lower = (glane & (u32)(k / (2 * R))) == 0;
#pragma unroll
            for (int r = 0; r < R / 2; r++) {
                const u32 a = __shfl_xor_sync(gmask, v[R - 1 - r], lm);
                const u32 b = __shfl_xor_sync(gmask, v[r], lm);
                v[r] = lower ? min(v[r], a) : max(v[r], a);
                v[R - 1 - r] = lower ? min(v[R - 1 - r], b) : max(v[R - 1 - r], b);
            }
        }
        if (k == N) break;  // the halves are separated: no need to finish the merge
#pragma unroll
        for (int j = k >> 2; j >= R; j >>= 1) {
            const int lm = j / R;
            const bool lower = (glane & (u32)lm) == 0;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const u32 w = __shfl_xor_sync(gmask, v[r], lm);
                v[r] = lower ? min(v[r], w) : max(v[r], w);
            }
        }
#pragma unroll
        for (int j = (R >> 1); j >= 1; j >>= 1) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int q = r ^ j;
                if (q > r) {
                    const u32 a = v[r], b = v[q];
                    v[r] = min(a, b);
                    v[q] = max(a, b);
                }
            }
        }
    }
    const bool lower = glane < G / 2;
    u32 x = v[0];
#pragma unroll
    for (int r = 1; r < R; r++) x = lower ? max(x, v[r]) : min(x, v[r]);
#pragma unroll
    for (int sft = G / 4; sft >= 1; sft >>= 1) {
        const u32 y = __shfl_xor_sync(gmask, x, sft);
        x = lower ? max(x, y) : min(x, y);
    }
    const u32 y = __shfl_xor_sync(gmask, x, G / 2);
    lo = lower ? x : y;
    hi = lower ? y : x;
}

template <int G, bool STAGED>
__device__ __forceinline__ void sweep3_run(const SeqDev *__restrict__ sd, u64 c_lo, u64 c_hi, const u32 *__restrict__ ls,
                                           const u32 *__restrict__ ubv, const u32 *__restrict__ ppv, const u32 *__restrict__ lbv,
                                           const SweepOut &o, u64 ops_packed, const u32 *ends, const u32 *rank, const u32 *bme,
                                           u32 *kbuf_all) {
    // Warp-synchronous: the 32 / G groups of a warp take every step together (a group with nothing left to do
    // is masked by its predicates), so every vote / shuffle is ONE full-warp instruction.  Per-group masks
    // would make the hardware run each group's copy of the instruction in turn.
    constexpr u32 FULL = 0xffffffffu;
    constexpr u32 CAP = SW3_KBUF / (SW3_TPB / G);
    constexpr int NGROUPS = SW3_TPB / G;
    constexpr u32 GM = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    const u32 glane = threadIdx.x & (G - 1), gidx = threadIdx.x / G;
    const u32 gshift = (threadIdx.x & 31) & ~(G - 1);
    const u32 lt = (1u << glane) - 1;
    const bool has_nulls = sd->has_nulls != 0;
    const double *__restrict__ sorted = sd->score_sorted;
    u32 *kbuf = kbuf_all + (size_t)gidx * CAP;
    const bool want_median = o.want_median != 0;
    const int k = o.k;
    // this lane's output column, if any (k <= 8 <= G)
    const int my_op = (int)glane < k ? (int)((ops_packed >> (4 * glane)) & 15u) : -1;
    // answers of the previous left: the loads are in flight while the next left is processed
    u64 pend_g = ~0ull;
    double pend_v1 = 0.0, pend_v2 = 0.0;
    bool pend_ok = false, pend_two = false;

    const bool pipe_lane = o.planar != nullptr && my_op >= 0 && sw3_is_pipe_op(my_op);
    uint4 pl = make_uint4(0, 0, 0, 0);
    u64 g = c_lo + gidx;
    u32 u = 0, p = 0, l = 0, lstart = 0;
    if (g < c_hi) {
        u = ubv[g];
        p = ppv[g];
        l = lbv[g];
        lstart = ls[g];
        if (pipe_lane) pl = o.planar[g];
    }
    const u64 g_warp = c_lo + (threadIdx.x >> 5) * (32 / G);
    const u32 rounds = g_warp < c_hi ? (u32)((c_hi - g_warp + NGROUPS - 1) / NGROUPS) : 0u;
    for (u32 round = 0; round < rounds; round++) {
        const bool live = g < c_hi;
        // parameters of the next left of this group
        const u64 gn = g + NGROUPS;
        u32 nu = 0, np = 0, nl = 0, nls = 0;
        uint4 npl = make_uint4(0, 0, 0, 0);
        if (gn < c_hi) {
            nu = ubv[gn];
            np = ppv[gn];
            nl = lbv[gn];
            nls = ls[gn];
            if (pipe_lane) npl = o.planar[gn];
        }
        u32 nvalid = 0;
        u32 rmin = GRB_RANK_NONE;  // unsigned min ignores NONE
        int rmax = -1;             // signed max ignores NONE (ranks are below 2^31)
        // A group that cannot fit the in-group median buffer (its size is known before any member is read) is left
        // entirely to k_median_heavy3, a whole CTA per left: one warp walking thousands of members would hold up
        // its CTA, and the ranks it collected would be thrown away.
        const bool pre_heavy = want_median && u > l && u - l > CAP;
        // ---- members that start inside the left: [p, u), ascending
        const u32 fw = (u > p && !pre_heavy) ? u - p : 0u;
        const u32 fw_steps = __reduce_max_sync(FULL, (fw + G - 1) / G);
        for (u32 it = 0; it < fw_steps; it++) {
            const u32 j = p + it * G + glane;
            const u32 r = j < u ? rank[j] : GRB_RANK_NONE;
            rmin = min(rmin, r);
            rmax = max(rmax, (int)r);
            if (has_nulls) {
                const bool v = r != GRB_RANK_NONE;
                const u32 vm = (__ballot_sync(FULL, v) >> gshift) & GM;
                if (want_median && v) {
                    const u32 slot = nvalid + __popc(vm & lt);
                    if (slot < CAP) kbuf[slot] = r;
                }
                nvalid += __popc(vm);
            } else if (want_median) {
                const u32 slot = j - p;
                if (j < u && slot < CAP) kbuf[slot] = r;
            }
        }
        if (!has_nulls) nvalid = fw;
        // ---- straddlers: below p, end > l.start; exactly p - l of them
        const u32 need = pre_heavy ? 0u : p - l;
        u32 found = 0;
        int cb = p ? (int)((p - 1) & ~(u32)(G - 1)) : -1;
        while (true) {
            const bool act = found < need && cb >= 0;
            if (!__any_sync(FULL, act)) break;
            // a group either looks at G rights or skips the rest of a 64-block whose largest end cannot overlap
            const bool look = act && bme[cb >> GRB_BME_SHIFT] > lstart;
            const u32 j = (u32)cb + glane;
            const bool hit = look && j < p && ends[j] > lstart;
            const u32 hm = (__ballot_sync(FULL, hit) >> gshift) & GM;
            const u32 r = hit ? rank[j] : GRB_RANK_NONE;
            rmin = min(rmin, r);
            rmax = max(rmax, (int)r);
            u32 vm = hm;
            bool v = hit;
            if (has_nulls) {
                v = r != GRB_RANK_NONE;
                vm = (__ballot_sync(FULL, v) >> gshift) & GM;
            }
            if (want_median && v) {
                const u32 slot = nvalid + __popc(vm & lt);
                if (slot < CAP) kbuf[slot] = r;
            }
            found += __popc(hm);
            nvalid += __popc(vm);
            if (act) cb = look ? cb - G : ((cb >> GRB_BME_SHIFT) << GRB_BME_SHIFT) - G;
        }
#pragma unroll
        for (int s = G / 2; s; s >>= 1) {
            rmin = min(rmin, __shfl_xor_sync(FULL, rmin, s));
            rmax = max(rmax, __shfl_xor_sync(FULL, rmax, s));
        }
        // ---- median: sorting network over the (all different) ranks; its size is the warp's largest group
        bool heavy = pre_heavy;
        u32 med_lo = 0, med_hi = 0;
        if (want_median) {
            heavy = pre_heavy || nvalid > CAP;  // (nvalid <= the group size, so the second test never adds anything)
            const u32 m = heavy ? 0u : nvalid;
            const u32 mx = __reduce_max_sync(FULL, m);
            __syncwarp();
            if (mx) {
                u32 a = 1, b = 1;
                if (mx <= 2 * G)
                    bitonic_median<G, 2>(kbuf, m, glane, a, b);
                else if (mx <= 4 * G)
                    bitonic_median<G, 4>(kbuf, m, glane, a, b);
                else if (mx <= 8 * G)
                    bitonic_median<G, 8>(kbuf, m, glane, a, b);
                else
                    bitonic_median<G, 16>(kbuf, m, glane, a, b);
                med_lo = a - 1u;
                med_hi = (m & 1) ? med_lo : b - 1u;
            }
        }
        // ---- the previous left's answers have arrived by now: store them
        if (pend_g != ~0ull) {
            o.out[pend_g * (u64)k + glane] = pend_ok ? (pend_two ? (pend_v1 + pend_v2) / 2.0 : pend_v1) : 0.0;
            o.out_valid[pend_g * (u64)k + glane] = pend_ok ? 1 : 0;
            pend_g = ~0ull;
        }
        // ---- this left's answers: issue the translations rank -> score
        {
            u32 r1 = 0, r2 = 0;
            bool ok = false, mine = false;
            if (my_op == GRB_OP_MIN) {
                mine = !heavy;
                ok = rmin != GRB_RANK_NONE;
                r1 = r2 = rmin;
            } else if (my_op == GRB_OP_MAX) {
                mine = !heavy;
                ok = rmax >= 0;
                r1 = r2 = (u32)rmax;
            } else if (my_op == GRB_OP_MEDIAN) {
                mine = !heavy;
                ok = nvalid != 0;
                r1 = med_lo;
                r2 = med_hi;
            }
            if (mine && live) {
                pend_g = g;
                pend_ok = ok;
                pend_two = r1 != r2;
                if (ok) {
                    pend_v1 = sorted[r1];
                    pend_v2 = sorted[r2];
                }
            }
        }
        if (pipe_lane && live) {
            u8 vv;
            o.out[g * (u64)k + glane] = sw3_pipe_col(my_op, pl, vv);
            o.out_valid[g * (u64)k + glane] = vv;
        }
        pl = npl;
        if (heavy && live && glane == 0) {
            const u32 slot = atomicAdd(o.heavy_count, 1u);
            o.heavy_list[slot] = (u32)g;  // the host checks total < 2^32 when a median is requested
        }
        __syncwarp();  // the next left reuses kbuf
        g = gn;
        u = nu;
        p = np;
        l = nl;
        lstart = nls;
    }
    if (pend_g != ~0ull) {
        o.out[pend_g * (u64)k + glane] = pend_ok ? (pend_two ? (pend_v1 + pend_v2) / 2.0 : pend_v1) : 0.0;
        o.out_valid[pend_g * (u64)k + glane] = pend_ok ? 1 : 0;
    }
}

// ---- wavelet runs: device code ---------------------------------------------------------------------------------

// Bitonic sort of P = 256 * nblk keys (nblk a power of two <= 8) in shared memory, ascending.  Warp w owns the block of 256
// consecutive keys [256 w, 256 w + 256): element 8 * lane + r sits in register r of `lane`, so partner distances below 8
// are register pairs, 8 .. 128 one shuffle per register, and only the distances >= 256 (6 of the 66 steps at P = 2048) go
// through shared memory.
__device__ __forceinline__ void wv_ce_regs(u32 (&v)[8], const int j, const u32 ebase, const u32 k) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int q = r ^ j;
        if (q > r) {
            const bool asc = k < 8 ? ((u32)r & k) == 0 : (ebase & k) == 0;
            const u32 a = v[r], b = v[q];
            const u32 lo = min(a, b), hi = max(a, b);
            v[r] = asc ? lo : hi;
            v[q] = asc ? hi : lo;
        }
    }
}
__device__ __forceinline__ void wv_ce_lanes(u32 (&v)[8], const int j, const u32 ebase, const u32 k, const u32 lane) {
    const int lm = j >> 3;
    const bool keep_min = ((ebase & k) == 0) == ((lane & (u32)lm) == 0);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const u32 w = __shfl_xor_sync(GRB_FULL, v[r], lm);
        v[r] = keep_min ? min(v[r], w) : max(v[r], w);
    }
}
__device__ __forceinline__ void wv_sort(u32 *keys, const u32 P) {
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 nblk = P >> 8;
    const u32 ebase = warp * 256 + lane * 8;
    u32 v[8];
    if (warp < nblk) {
        const uint4 a = *reinterpret_cast<const uint4 *>(keys + ebase), b = *reinterpret_cast<const uint4 *>(keys + ebase + 4);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
#pragma unroll
        for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j >= 1; j >>= 1) {
                if (j >= 8)
                    wv_ce_lanes(v, j, ebase, (u32)k, lane);
                else
                    wv_ce_regs(v, j, ebase, (u32)k);
            }
        }
        *reinterpret_cast<uint4 *>(keys + ebase) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4 *>(keys + ebase + 4) = make_uint4(v[4], v[5], v[6], v[7]);
    }
    for (u32 k = 512; k <= P; k <<= 1) {
        for (u32 j = k >> 1; j >= 256; j >>= 1) {
            __syncthreads();
            for (u32 t = tid; t < (P >> 1); t += SW3_TPB) {
                const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const u32 a = keys[i], b = keys[i | j];
                const bool asc = (i & k) == 0;
                if ((a > b) == asc) {
                    keys[i] = b;
                    keys[i | j] = a;
                }
            }
        }
        __syncthreads();
        if (warp < nblk) {
            const uint4 a = *reinterpret_cast<const uint4 *>(keys + ebase), b = *reinterpret_cast<const uint4 *>(keys + ebase + 4);
            v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
#pragma unroll
            for (int j = 128; j >= 1; j >>= 1) {
                if (j >= 8)
                    wv_ce_lanes(v, j, ebase, k, lane);
                else
                    wv_ce_regs(v, j, ebase, k);
            }
            *reinterpret_cast<uint4 *>(keys + ebase) = make_uint4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<uint4 *>(keys + ebase + 4) = make_uint4(v[4], v[5], v[6], v[7]);
        }
    }
    __syncthreads();
}

// k-th smallest key (0-based) of: the first x elements of A minus the first y elements of B.  Per level: the node starts
// at position s of both level sequences (dense keys), s / 2 zeros precede it, and one word + one count per tree give the
// zeros among the node's first xa (xb) elements.
__device__ __forceinline__ u32 wv_kth(const WaveSmem &w, const int L, const u32 x, const u32 y, u32 k) {
    u32 s = 0, xa = x, xb = y;
    const uint2 *ta = w.tree[0][0], *tb = w.tree[1][0];
    for (u32 half = 1u << (L - 1); half; half >>= 1, ta += WV_ROW, tb += WV_ROW) {
        const u32 pa = s + xa, pb = s + xb;
        const uint2 qa = ta[pa >> 5], qb = tb[pb >> 5];
        const u32 z0 = s >> 1;
        const u32 za = qa.y + __popc(~qa.x & ((1u << (pa & 31)) - 1u)) - z0;
        const u32 zb = qb.y + __popc(~qb.x & ((1u << (pb & 31)) - 1u)) - z0;
        const u32 dz = za - zb;
        const bool one = k >= dz;
        k -= one ? dz : 0u;
        s += one ? half : 0u;
        xa = one ? xa - za : za;
        xb = one ? xb - zb : zb;
    }
    return s;
}

struct WaveRun {
    u32 wlo, W;        // slice [wlo, wlo + W) in start order
    u32 lbmin, lbmax;  // end-order range of the run
    u32 qlo;           // first staged entry of perm_e (lbmin rounded down to 4)
    u32 lsmin, lsmax;  // smallest / largest left start of the run
};

// The slice (rank, ends, perm) is in shared memory; builds the trees and answers the run's lefts.
__device__ __forceinline__ void wave_run(const SeqDev *__restrict__ sd, WaveSmem &w, const WaveRun q, u64 c_lo, u64 c_hi,
                                         const u32 *__restrict__ ubv, const u32 *__restrict__ lbv, const SweepOut &o, u64 ops_packed) {
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 W = q.W;
    u32 P = 256;
    while (P < W) P <<= 1;
    int L = 1;
    while ((1u << L) < W) L++;
    const u32 nC = q.lbmin - q.wlo, nM = q.lbmax - q.lbmin;
    // ---- keys: None scores sort last, each on its own
    for (u32 i = tid; i < P; i += SW3_TPB) {
        u32 r = 0xffffffffu;
        if (i < W) {
            r = w.rank[i];
            if (r == GRB_RANK_NONE) r = 0x80000000u | i;
            w.rank[i] = r;
        }
        w.sorted[i] = r;
    }
    if (tid < 2) w.cnt[tid] = 0;
    __syncthreads();
    wv_sort(w.sorted, P);
    // ---- local rank of every position = where its key sits in the sorted slice: A's first sequence
    for (u32 i = tid; i < W; i += SW3_TPB) {
        const u32 r = w.rank[i];
        u32 lo = 0;
        for (u32 step = P >> 1; step; step >>= 1)
            if (w.sorted[lo + step - 1] < r) lo += step;
        w.seq0[0][i] = (u16)lo;
    }
    __syncthreads();
    // ---- B's first sequence: C ++ end-order range ++ R
    for (u32 i = tid; i < W; i += SW3_TPB) {
        const u32 e = w.ends[i];
        if (e <= q.lsmin) {
            const u32 slot = atomicAdd(&w.cnt[0], 1u);
            if (slot < nC) w.seq0[1][slot] = w.seq0[0][i];
        } else if (e > q.lsmax) {
            const u32 d = nC + nM + atomicAdd(&w.cnt[1], 1u);
            if (d < W) w.seq0[1][d] = w.seq0[0][i];
        }
    }
    for (u32 i = tid; i < nM; i += SW3_TPB) {
        const u32 j = w.perm[i + (q.lbmin - q.qlo)] - q.wlo;
        if (j < W && nC + i < W) w.seq0[1][nC + i] = w.seq0[0][j];
    }
    __syncthreads();
    // ---- the levels.  Warp w owns the chunks (32 positions) [w cpw, (w + 1) cpw) of both level sequences: it keeps their keys
    // and bitmaps in registers, publishes its zero count, and after one barrier knows the zeros before each of its chunks:
    // the words' counts are written and the keys move to their places in the next level (inside every node the zeros first,
    // then the ones, both in order).  Two barriers per level.
    const u32 nch = (W + 31) >> 5;
    const u32 cpw = (nch + SW3_TPB / 32 - 1) / (SW3_TPB / 32);  // <= 8
    const u32 ltmask = (1u << lane) - 1u;
    for (int l = 0; l < L; l++) {
        const int sh = L - 1 - l;
        u32 word[2][8];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const u16 *cur = (l & 1) ? w.seq1[t] : w.seq0[t];
            u32 zsum = 0;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const u32 i = (warp * cpw + e) * 32 + lane;
                const u32 key = (u32)e < cpw && i < W ? cur[i] : 0xffffu;  // beyond the end: ones, never counted
                word[t][e] = __ballot_sync(GRB_FULL, (key >> sh) & 1u);
                zsum += 32 - __popc(word[t][e]);
            }
            if (lane == 0) w.wtot[t][warp] = zsum;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const u16 *cur = (l & 1) ? w.seq1[t] : w.seq0[t];
            u16 *nxt = (l & 1) ? w.seq0[t] : w.seq1[t];
            uint2 *row = w.tree[t][l];
            u32 zrun = 0;  // zeros of the level before this warp's first chunk
#pragma unroll
            for (int v = 0; v < SW3_TPB / 32; v++) zrun += (u32)v < warp ? w.wtot[t][v] : 0u;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const u32 c = warp * cpw + e;
                if ((u32)e < cpw && c < nch) {
                    const u32 wd = word[t][e];
                    const u32 zc = 32 - __popc(wd);
                    if (lane == 0) {
                        row[c] = make_uint2(wd, zrun);
                        if (c == nch - 1) row[nch] = make_uint2(0xffffffffu, zrun + zc);  // a query may ask for position W
                    }
                    const u32 i = c * 32 + lane;
                    if (sh && i < W) {
                        const u32 key = cur[i];
                        const u32 zb = zrun + __popc(~wd & ltmask);            // zeros before i on this level
                        const u32 s = (key >> (sh + 1)) << (sh + 1);           // the node of this key starts here
                        const u32 zin = zb - (s >> 1);                         // zeros of the node before i
                        const u32 dest = (wd >> lane) & 1u ? s + (1u << sh) + (i - s - zin) : s + zin;
                        if (dest < W) nxt[dest] = (u16)key;  // (always, for the permutations valid lefts produce)
                    }
                    zrun += zc;
                }
            }
        }
        __syncthreads();
    }
    // ---- the lefts: one thread each
    const bool has_nulls = sd->has_nulls != 0;
    const double *__restrict__ sorted_scores = sd->score_sorted;
    const int k = o.k;
    bool want_min = false, want_max = false;
    for (int c = 0; c < k; c++) {
        const int op = (int)((ops_packed >> (4 * c)) & 15u);
        want_min |= op == GRB_OP_MIN;
        want_max |= op == GRB_OP_MAX;
    }
    for (u64 g = c_lo + tid; g < c_hi; g += SW3_TPB) {
        const u32 u = ubv[g], l = lbv[g];
        u32 m = 0;
        if (u > l && u >= q.wlo && u - q.wlo <= W && l >= q.lbmin && l <= q.lbmax) m = has_nulls ? sd->vcnt_a[u] - sd->vcnt_b[l] : u - l;
        const u32 x = u - q.wlo, y = nC + (l - q.lbmin);
        double v_min = 0.0, v_max = 0.0, v_med = 0.0;
        if (m) {
            if (want_min) v_min = sorted_scores[w.sorted[wv_kth(w, L, x, y, 0u)]];
            if (want_max) v_max = sorted_scores[w.sorted[wv_kth(w, L, x, y, m - 1u)]];
            if (o.want_median) {
                v_med = sorted_scores[w.sorted[wv_kth(w, L, x, y, (m - 1u) >> 1)]];
                if (!(m & 1u)) v_med = (v_med + sorted_scores[w.sorted[wv_kth(w, L, x, y, m >> 1)]]) / 2.0;
            }
        }
        uint4 pl = make_uint4(0, 0, 0, 0);
        if (o.planar != nullptr) pl = o.planar[g];
        for (int c = 0; c < k; c++) {
            const int op = (int)((ops_packed >> (4 * c)) & 15u);
            if (op == GRB_OP_MIN || op == GRB_OP_MAX || op == GRB_OP_MEDIAN) {
                o.out[g * (u64)k + c] = op == GRB_OP_MIN ? v_min : op == GRB_OP_MAX ? v_max : v_med;
                o.out_valid[g * (u64)k + c] = m ? 1 : 0;
            } else if (o.planar != nullptr && sw3_is_pipe_op(op)) {
                u8 vv;
                o.out[g * (u64)k + c] = sw3_pipe_col(op, pl, vv);
                o.out_valid[g * (u64)k + c] = vv;
            }
        }
    }
}

// One CTA per tile of 1024 lefts (the tiling of k_tile).
__global__ void __launch_bounds__(SW3_TPB, 4) k_sweep3(const SeqDev *__restrict__ seqs, const u32 *__restrict__ tile_begin, int nseq,
                                                    const u32 *__restrict__ ls, const u32 *__restrict__ ubv,
                                                    const u32 *__restrict__ ppv, const u32 *__restrict__ lbv, const SweepOut o,
                                                    int stage_mode) {
    extern __shared__ __align__(128) unsigned char sw3_raw[];
    Sweep3Smem &S = *reinterpret_cast<Sweep3Smem *>(sw3_raw);
    const u32 t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const SeqDev *__restrict__ sd = &seqs[find_seq_by_tile(tile_begin, nseq, t)];
    const u64 g_first = ((sd->left_begin >> TILE_SHIFT) + (t - sd->tile_begin)) << TILE_SHIFT;
    const u64 g_lo = sd->left_begin > g_first ? sd->left_begin : g_first;
    const u64 g_hi = sd->left_end < g_first + TILE ? sd->left_end : g_first + TILE;
    u64 ops_packed = 0;  // 4 bits per reduction (k <= 8): no dynamic indexing of the parameter block
#pragma unroll
    for (int c = 0; c < 8; c++)
        if (c < o.k) ops_packed |= (u64)(o.ops[c] & 15) << (4 * c);
    if (sd->n == 0 || !sd->has_scores) {
        // absent seqname / no rights: every group is empty
        for (u64 g = g_lo + tid; g < g_hi; g += SW3_TPB)
            for (int c = 0; c < o.k; c++) {
                const int op = (int)((ops_packed >> (4 * c)) & 15u);
                if (op == GRB_OP_MIN || op == GRB_OP_MAX || op == GRB_OP_MEDIAN) {
                    o.out[g * (u64)o.k + c] = 0.0;
                    o.out_valid[g * (u64)o.k + c] = 0;
                } else if (o.planar != nullptr && sw3_is_pipe_op(op)) {
                    u8 vv;
                    o.out[g * (u64)o.k + c] = sw3_pipe_col(op, o.planar[g], vv);
                    o.out_valid[g * (u64)o.k + c] = vv;
                }
            }
        return;
    }
    if (tid == 0) {
        mbar_init(&S.bar, 1);
    }
    __syncthreads();
    u32 phase = 0;
    u64 c_lo = g_lo;
    u64 run = g_hi - g_lo;
    while (c_lo < g_hi) {
        const u64 c_hi = c_lo + run < g_hi ? c_lo + run : g_hi;
        // the slice of rights this run of lefts can touch, and its members per left
        u32 pmin = 0xffffffffu, umax = 0, mem = 0, lbmin = 0xffffffffu, lbmax = 0, lsmin = 0xffffffffu, lsmax = 0;
        for (u64 g = c_lo + tid; g < c_hi; g += SW3_TPB) {
            const u32 u = ubv[g], p = ppv[g], l = lbv[g], x = ls[g];
            pmin = min(pmin, p);
            umax = max(umax, max(u, p));  // p > u only for an invalid left (flagged by k_tile); stay inside the slice
            mem += u - l;
            lbmin = min(lbmin, l);
            lbmax = max(lbmax, l);
            lsmin = min(lsmin, x);
            lsmax = max(lsmax, x);
        }
        pmin = warp_min(pmin);
        umax = warp_max(umax);
        mem = warp_sum(mem);
        lbmin = warp_min(lbmin);
        lbmax = warp_max(lbmax);
        lsmin = warp_min(lsmin);
        lsmax = warp_max(lsmax);
        if (lane == 0) {
            S.red[0][warp] = pmin;
            S.red[1][warp] = umax;
            S.red[2][warp] = mem;
            S.red[3][warp] = lbmin;
            S.red[4][warp] = lbmax;
            S.red[5][warp] = lsmin;
            S.red[6][warp] = lsmax;
        }
        __syncthreads();
        if (warp == 0) {
            const bool in = lane < SW3_TPB / 32;
            u32 a = in ? S.red[0][lane] : 0xffffffffu;
            u32 b = in ? S.red[1][lane] : 0u;
            u32 c = in ? S.red[2][lane] : 0u;
            u32 d = in ? S.red[3][lane] : 0xffffffffu;
            u32 e = in ? S.red[4][lane] : 0u;
            u32 f = in ? S.red[5][lane] : 0xffffffffu;
            u32 h = in ? S.red[6][lane] : 0u;
            a = warp_min(a);
            b = warp_max(b);
            c = warp_sum(c);
            d = warp_min(d);
            e = warp_max(e);
            f = warp_min(f);
            h = warp_max(h);
            if (lane == 0) {
                u32 wlo = a ? sd->fb[(a - 1) >> GRB_BME_SHIFT] : 0u;
                S.bcast[7] = wlo & ~3u;  // wavelet runs: bulk copies need 16-byte aligned sources
                wlo &= ~255u;            // member sweep: also for the block-max-end slice
                S.bcast[0] = wlo;
                S.bcast[1] = b;
                S.bcast[2] = c;
                S.bcast[3] = d;
                S.bcast[4] = e;
                S.bcast[5] = f;
                S.bcast[6] = h;
            }
        }
        __syncthreads();
        const u32 wlo = S.bcast[0], whi = S.bcast[1];
        // lanes per left: 8 up to ~64 members per left on average, 16 up to ~150, else a whole warp (the in-group
        // median buffers hold 16 ranks per lane, and a left's members range from half to twice the average)
        const u64 avg_den = c_hi - c_lo;
        const int gsel = (u64)S.bcast[2] > 150ull * avg_den ? 32 : (u64)S.bcast[2] > 64ull * avg_den ? 16 : 8;
        const u32 W = whi > wlo ? whi - wlo : 0u;
        // a wavelet run: the slice fits WV_CAP, the run is long enough to pay for the trees, and the end-order range is
        // consistent with the slice (it always is for valid lefts; an invalid one is flagged by k_tile and answered anyhow)
        WaveRun wq;
        wq.wlo = S.bcast[7], wq.W = whi > wq.wlo ? whi - wq.wlo : 0u;
        wq.lbmin = S.bcast[3], wq.lbmax = S.bcast[4], wq.qlo = S.bcast[3] & ~3u, wq.lsmin = S.bcast[5], wq.lsmax = S.bcast[6];
        const bool wave_on = stage_mode && o.wave_ratio > 0 && o.want_median;
        const bool wave_dense = (c_hi - c_lo) * (u64)o.wave_ratio >= wq.W;
        const bool wave = wave_on && wave_dense && wq.W && wq.W <= WV_CAP && wq.lbmin >= wq.wlo && wq.lbmax >= wq.lbmin &&
                          wq.lbmax - wq.wlo <= wq.W;
        if (wave) {
            if (tid == 0) {
                const u32 bytes = ((wq.W + 3) & ~3u) * 4;
                const u32 pbytes = ((wq.lbmax - wq.qlo + 3) & ~3u) * 4;
                mbar_arrive_expect_tx(&S.bar, 2 * bytes + pbytes);
                tma_load_1d(S.w.ends, sd->ends + wq.wlo, bytes, &S.bar);
                tma_load_1d(S.w.rank, sd->rank_a + wq.wlo, bytes, &S.bar);
                if (pbytes) tma_load_1d(S.w.perm, sd->perm_e + wq.qlo, pbytes, &S.bar);
            }
            mbar_wait(&S.bar, phase);
            phase ^= 1;
            wave_run(sd, S.w, wq, c_lo, c_hi, ubv, lbv, o, ops_packed);
            __syncthreads();  // the slice and S.bcast are rewritten by the next run
            c_lo = c_hi;
            continue;
        }
        if (wave_on && wave_dense && wq.W > WV_CAP && c_hi - c_lo > SW3_MIN_RUN) {
            run = (c_hi - c_lo + 1) >> 1;  // dense lefts over a slice too wide for the trees: halve the run, not the method
            __syncthreads();
            continue;
        }
        const bool fits = stage_mode && W <= SW3_WCAP;
        if (!fits && stage_mode && c_hi - c_lo > SW3_MIN_RUN) {
            run = (c_hi - c_lo + 1) >> 1;
            __syncthreads();  // S.bcast is rewritten by the next attempt
            continue;
        }
        if (fits) {
            if (tid == 0 && W) {
                const u32 bytes = ((W + 3) & ~3u) * 4;
                const u32 nb = ((whi + 63) >> GRB_BME_SHIFT) - (wlo >> GRB_BME_SHIFT);
                const u32 bbytes = ((nb + 3) & ~3u) * 4;
                mbar_arrive_expect_tx(&S.bar, 2 * bytes + bbytes);
                tma_load_1d(S.ends, sd->ends + wlo, bytes, &S.bar);
                tma_load_1d(S.rank, sd->rank_a + wlo, bytes, &S.bar);
                tma_load_1d(S.bme, sd->bme + (wlo >> GRB_BME_SHIFT), bbytes, &S.bar);
            }
            if (W) {
                mbar_wait(&S.bar, phase);
                phase ^= 1;
            }
            const u32 *e_p = S.ends - wlo, *r_p = S.rank - wlo, *b_p = S.bme - (wlo >> GRB_BME_SHIFT);
            if (gsel == 32)
                sweep3_run<32, true>(sd, c_lo, c_hi, ls, ubv, ppv, lbv, o, ops_packed, e_p, r_p, b_p, S.kbuf);
            else if (gsel == 16)
                sweep3_run<16, true>(sd, c_lo, c_hi, ls, ubv, ppv, lbv, o, ops_packed, e_p, r_p, b_p, S.kbuf);
            else
                sweep3_run<8, true>(sd, c_lo, c_hi, ls, ubv, ppv, lbv, o, ops_packed, e_p, r_p, b_p, S.kbuf);
        } else {
            if (gsel == 32)
                sweep3_run<32, false>(sd, c_lo, c_hi, ls, ubv, ppv, lbv, o, ops_packed, sd->ends, sd->rank_a, sd->bme, S.kbuf);
            else if (gsel == 16)
                sweep3_run<16, false>(sd, c_lo, c_hi, ls, ubv, ppv, lbv, o, ops_packed, sd->ends, sd->rank_a, sd->bme, S.kbuf);
            else
                sweep3_run<8, false>(sd, c_lo, c_hi, ls, ubv, ppv, lbv, o, ops_packed, sd->ends, sd->rank_a, sd->bme, S.kbuf);
        }
        __syncthreads();  // the slice and S.bcast are rewritten by the next run
        c_lo = c_hi;
    }
}


// Median of the groups too large for the in-group networks of k_sweep3 (more than 16 ranks per lane): one CTA per such
// left.  The straddlers' positions are found once (backward walk with block-max-end skipping) and kept in shared
// memory; then an MSB-first radix select over the 32-bit ranks, 10 bits per pass, reads the members -- the run
// [pp, ub) and the cached straddlers -- once per pass, and one more pass finds the upper middle of an even group.
// (k_median_heavy, the score-based version, needs 8 passes per middle element and re-walks the candidates each time.)
constexpr int MH3_TPB = 256;
constexpr u32 MH3_STRAD = 6144;  // straddler positions kept in shared memory; beyond that the walk is repeated per pass
constexpr u32 MH3_BINS = 1024;

template <typename F>
__device__ __forceinline__ void mh3_for_straddlers(const u32 *__restrict__ ends, const u32 *__restrict__ bme, u32 p, u32 need, u32 lstart,
                                                   F &&visit) {
    // CTA-wide backward walk: 256 positions per step, whole 64-blocks skipped through their block-max-end
    u32 found = 0;
    long long top = p;  // positions [top - 256, top) are examined next
    while (found < need && top > 0) {
        const long long j = top - 1 - threadIdx.x;
        bool hit = false;
        if (j >= 0 && bme[j >> GRB_BME_SHIFT] > lstart) hit = ends[j] > lstart;
        if (hit) visit((u32)j);
        found += __syncthreads_count(hit);
        top -= MH3_TPB;
    }
}

__global__ void __launch_bounds__(MH3_TPB) k_median_heavy3(const SeqDev *__restrict__ seqs, int nseq, const u32 *__restrict__ ls,
                                                           const u32 *__restrict__ ubv, const u32 *__restrict__ ppv,
                                                           const u32 *__restrict__ lbv, const u32 *__restrict__ heavy_list,
                                                           const u32 *__restrict__ heavy_count, const SweepOut o) {
    __shared__ u32 s_strad[MH3_STRAD];
    __shared__ u32 hist[MH3_BINS];
    __shared__ u32 s_n, s_prefix, s_k, s_total, s_red[MH3_TPB / 32];
    __shared__ int s_redmax[MH3_TPB / 32];
    const u32 nheavy = *heavy_count;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (u32 h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const u64 g = heavy_list[h];
        const SeqDev *__restrict__ sd = &seqs[find_seq_by_left(seqs, nseq, g)];
        const u32 u = ubv[g], p = ppv[g], l = lbv[g], lstart = ls[g];
        const u32 need = p - l;
        const u32 *__restrict__ ends = sd->ends;
        const u32 *__restrict__ bme = sd->bme;
        const u32 *__restrict__ rank = sd->rank_a;
        if (tid == 0) s_n = 0;
        __syncthreads();
        // ---- the straddlers, once
        mh3_for_straddlers(ends, bme, p, need, lstart, [&](u32 j) {
            const u32 slot = atomicAdd(&s_n, 1u);
            if (slot < MH3_STRAD) s_strad[slot] = j;
        });
        __syncthreads();
        const bool cached = s_n <= MH3_STRAD;
        const u32 nstrad = cached ? s_n : 0;
        // visit every member's rank: [p, u) and the straddlers (from the cache, or by walking again)
        auto for_members = [&](auto &&f) {
            for (u32 j = p + tid; j < u; j += MH3_TPB) f(rank[j]);
            if (cached) {
                for (u32 q = tid; q < nstrad; q += MH3_TPB) f(rank[s_strad[q]]);
            } else {
                mh3_for_straddlers(ends, bme, p, need, lstart, [&](u32 j) { f(rank[j]); });
            }
        };
        // ---- radix select, most significant digit first
        int bits = 1;
        while (bits < 32 && (sd->n >> bits)) bits++;  // ranks are below n
        const int passes = (bits + 9) / 10;
        if (tid == 0) {
            s_prefix = 0;
            s_k = 0;
        }
        u32 m = 0;  // members that carry a score
        u32 t_min = GRB_RANK_NONE;
        int t_max = -1;
        for (int pass = 0; pass < passes; pass++) {
            const int shift = 10 * (passes - 1 - pass);
            for (u32 i = tid; i < MH3_BINS; i += MH3_TPB) hist[i] = 0;
            __syncthreads();
            const u32 prefix = s_prefix;
            for_members([&](u32 r) {
                if (r == GRB_RANK_NONE) return;
                if (pass == 0) {
                    t_min = min(t_min, r);
                    t_max = max(t_max, (int)r);
                }
                if (pass == 0 || (r >> (shift + 10)) == (prefix >> (shift + 10))) atomicAdd(&hist[(r >> shift) & (MH3_BINS - 1)], 1u);
            });
            __syncthreads();
            // one warp finds the digit: 32 bins per lane, then a scan of the lane totals
            if (warp == 0) {
                u32 mine = 0;
                for (u32 i = 0; i < 32; i++) mine += hist[lane * 32 + i];
                u32 incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 t = __shfl_up_sync(GRB_FULL, incl, d);
                    if (lane >= (u32)d) incl += t;
                }
                const u32 total = __shfl_sync(GRB_FULL, incl, 31);
                if (pass == 0) {
                    if (lane == 0) {
                        s_total = total;
                        s_k = total ? (total - 1) >> 1 : 0;  // lower middle
                    }
                }
                __syncwarp();
                const u32 kk = pass == 0 ? (total ? (total - 1) >> 1 : 0) : s_k;
                const u32 excl = incl - mine;
                const bool here = total && kk >= excl && kk < incl;
                if (here) {
                    u32 c = excl;
                    u32 digit = lane * 32;
                    for (u32 i = 0; i < 32; i++) {
                        const u32 hv = hist[lane * 32 + i];
                        if (kk < c + hv) {
                            digit = lane * 32 + i;
                            break;
                        }
                        c += hv;
                    }
                    s_k = kk - c;
                    s_prefix = prefix | (digit << shift);
                }
            }
            __syncthreads();
            if (pass == 0) m = s_total;
        }
        // smallest / largest rank of the group (k_sweep3 left this group's MIN / MAX columns to this kernel too)
        t_min = warp_min(t_min);
        t_max = __reduce_max_sync(GRB_FULL, t_max);
        __syncthreads();
        if (lane == 0) {
            s_red[warp] = t_min;
            s_redmax[warp] = t_max;
        }
        __syncthreads();
        t_min = GRB_RANK_NONE;
        t_max = -1;
        for (int w = 0; w < MH3_TPB / 32; w++) {
            t_min = min(t_min, s_red[w]);
            t_max = max(t_max, s_redmax[w]);
        }
        __syncthreads();
        const u32 med_lo = s_prefix;
        u32 med_hi = med_lo;
        if (m && !(m & 1)) {
            // even: the upper middle is the smallest rank above the lower middle
            u32 best = GRB_RANK_NONE;
            for_members([&](u32 r) {
                if (r != GRB_RANK_NONE && r > med_lo) best = min(best, r);
            });
            best = warp_min(best);
            if (lane == 0) s_red[warp] = best;
            __syncthreads();
            best = GRB_RANK_NONE;
            for (int w = 0; w < MH3_TPB / 32; w++) best = min(best, s_red[w]);
            med_hi = best;
        }
        if (tid == 0) {
            double r = 0.0;
            if (m) {
                const double a = sd->score_sorted[med_lo];
                r = med_hi != med_lo ? (a + sd->score_sorted[med_hi]) / 2.0 : a;
            }
            for (int c = 0; c < o.k; c++) {
                double v;
                if (o.ops[c] == GRB_OP_MEDIAN)
                    v = r;
                else if (o.ops[c] == GRB_OP_MIN)
                    v = m ? sd->score_sorted[t_min] : 0.0;
                else if (o.ops[c] == GRB_OP_MAX)
                    v = m ? sd->score_sorted[t_max] : 0.0;
                else
                    continue;
                o.out[g * (u64)o.k + c] = v;
                o.out_valid[g * (u64)o.k + c] = m ? 1 : 0;
            }
        }
        __syncthreads();
    }
}
