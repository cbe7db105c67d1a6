// pipe.cu -- the pipelined prefix-sum map kernel (headline path: `granges map --func mean`).
//
// Replaces, per left range, COITrees::query + the gather of JoinDataLeftEmpty::map + FloatOperation::run for the
// reductions a prefix sum answers (src/ranges/coitrees.rs:48-57, src/join.rs:343-361, src/commands.rs:524-542,
// src/data/operations.rs:60-70,87-95):  count = ub - lb,  sum = PA[ub] - PB[lb]  with
//     ub = #{r.start < l.end}     lb = #{r.end <= l.start}.
//
// A persistent, warp-specialised TMA pipeline so that no warp ever waits on a dependent chain of global loads.
// One CTA per SM, tiles blockIdx.x, blockIdx.x + gridDim.x, ...; per tile (1024 consecutive lefts of one seqname):
//
//   warp 0  lefts loader   A: TMA the lefts (ls, le) of tile j into the lefts ring, as far ahead as
//                             the ring allows                                                 -> fullL
//   warps 1, 3, 4, 5  surveyors (tile j goes to surveyor j mod 4)
//                          B: min / max of the tile's lefts out of shared memory -> coordinate bins
//                             -> four bin-table entries by plain loads                        -> statsL
//   warp 2  window loader  C: TMA the slices the tile can touch -- starts, ends_sorted, the low halves of both
//                             bin tables, both prefix arrays (and, per variant, the high planes of 96-bit
//                             prefixes and the low halves of the valid-score counts) -- into the window ring,
//                             one bulk copy per lane, one mbarrier                            -> fullW
//   warps 6.. consumers:      wait fullW, resolve ub / lb of 2 lefts per thread out of shared
//                             memory, exact sums from the staged prefixes, write results,
//                             release the slots                                               -> emptyW, emptyL
//
// Variants (template VAR, chosen per run of seqnames by what their indexes hold):
//   PIPE_VAR_NULLS  some scores are None (BED "."): the mean divides by the number of members that carry a score,
//                   vcnt_a[ub] - vcnt_b[lb]; only the low 16 bits of those counts are staged (groups below 65536).
//   PIPE_VAR_WIDE   the scores need more than 54 bits of fixed point (decimal fractions over a few decades):
//                   prefixes mod 2^96 in two planes, u64 low + u32 high.
// Non-finite scores (+-inf, NaN) ride in the NULLS variants: the seqname carries prefix counters nf_a / nf_b that are
// read from global memory for its lefts, and the IEEE outcome of the reference's left-to-right sum follows by rule.
//
// Tiles whose windows do not fit the ring slots (sparse lefts over dense rights, unsorted lefts) are resolved by
// per-left global searches inside the consumers; results are identical.
#include <type_traits>

#include "tilelib.cuh"

namespace {

#ifndef PIPE_LUT16
#define PIPE_LUT16 0    // A/B switch: 1 = stage the low halves of the bin tables in every variant (measured: 0.7 % slower, 0.13 GB less traffic)
#endif
#ifndef PIPE_ALIGN8
#define PIPE_ALIGN8 0   // A/B switch: 1 = slices start on multiples of 8 entries in every variant
#endif
// The staged bin tables: u16 low halves where PIPE_LUT16 says so, and always in the NULLS variants (whose extra planes
// need the room): within one staged window (< 65536 keys) the low halves of the table entries are enough.
template <int VAR>
struct PipeLut {
    static constexpr bool L16 = PIPE_LUT16 || (VAR & PIPE_VAR_NULLS);
    typedef typename std::conditional<L16, u16, u32>::type type;
    // slices start on multiples of AL entries: 16-byte aligned bulk-copy sources for 4-byte (and, with AL = 8, 2-byte) planes
    static constexpr u32 AL = (L16 || PIPE_ALIGN8) ? 8u : 4u;
};
constexpr int PIPE_NCG = 16;    // consumer warps
constexpr int PIPE_LPT = TILE / (PIPE_NCG * 32);  // consecutive lefts per consumer thread
static_assert(PIPE_LPT == 2, "consumer layout");
constexpr int PIPE_NSURV = 4;  // surveyor warps (tile j goes to surveyor j % PIPE_NSURV)
constexpr int PIPE_THREADS = (PIPE_NCG + 2 + PIPE_NSURV) * 32;
#ifndef PIPE_LSTAGES_N
#define PIPE_LSTAGES_N 8   // A/B switch: lefts-ring depth
#endif
#ifndef PIPE_WST_N
#define PIPE_WST_N 4       // A/B switch: window-ring depth of the variants without extra planes
#endif
constexpr int PIPE_LSTAGES = PIPE_LSTAGES_N;
constexpr u32 PIPE_WIN = 1280;  // keys (and prefixes) per staged window
constexpr u32 PIPE_LUT = 640;   // bin-table entries per staged slice
constexpr int PIPE_MAXSEQ = 32; // seqnames per launch (their descriptors live in shared memory)
#ifndef PIPE_ASSIGN_AB
#define PIPE_ASSIGN_AB 0        // A/B build: GRB_PIPE_ASSIGN=1 deals runs of PIPE_RUN consecutive tiles to the CTAs instead of single tiles
#endif
#ifndef PIPE_RUN
#define PIPE_RUN 4u
#endif

struct PipeSeq {
    const u32 *starts, *ends_sorted, *lut_a, *lut_b;
    const u16 *lut16_a, *lut16_b, *vc16_a, *vc16_b;
    const u64 *pfx_a, *pfx_b;
    const u32 *pfh_a, *pfh_b;
    const uint4 *nf_a, *nf_b;
    u64 left_begin, left_end;
    double scale;
    u32 tile_begin, bmax, fast_cmax, n;
    int sh;
};

struct PipeTile {      // one per lefts-ring slot
    u64 g_first;       // global index of the tile's slot 0                (loader)
    u32 vlo, vhi;      // valid slots of the tile: [vlo, vhi)              (loader)
    int seq;           //                                                  (loader)
    u32 bin[4];        // bins of min l.end, max l.end, min l.start+1, max l.start+1   (surveyor)
    u32 bound[4];      // lut_a[bin0], lut_a[bin1 + 1], lut_b[bin2], lut_b[bin3 + 1]     (surveyor)
    u32 nf;            // NULLS variants: 1 = some group of the tile may hold a non-finite score (surveyor)
};

struct PipeDesc {      // one per window-ring slot (loader -> consumers)
    u64 g_first;
    u32 vlo, vhi;
    u32 a0, b0, la0, lb0;
    u32 staged;        // 1: windows staged in the slot; 0: search in global memory (windows too large, or invalid lefts)
    u32 nf;            // copied from the tile
    int seq;
};

template <int VAR>
struct PipeSmemT {
    static constexpr int WST = (VAR & PIPE_VAR_WIDE) ? (PIPE_LSTAGES_N > 8 ? 2 : 3) : PIPE_WST_N;  // window-ring depth: what fits beside the extra planes
    static constexpr u32 NH = (VAR & PIPE_VAR_WIDE) ? PIPE_WIN + 8 : 4;
    static constexpr u32 NV = (VAR & PIPE_VAR_NULLS) ? PIPE_WIN + 8 : 8;
    alignas(16) u32 ls[PIPE_LSTAGES][TILE];
    alignas(16) u32 le[PIPE_LSTAGES][TILE];
    alignas(16) u64 pa[WST][PIPE_WIN + 8];
    alignas(16) u64 pb[WST][PIPE_WIN + 8];
    alignas(16) u32 ka[WST][PIPE_WIN + 8];
    alignas(16) u32 kb[WST][PIPE_WIN + 8];
    alignas(16) u32 pah[WST][NH];
    alignas(16) u32 pbh[WST][NH];
    typedef typename PipeLut<VAR>::type lut_t;
    alignas(16) lut_t la[WST][PIPE_LUT + 8];
    alignas(16) lut_t lb[WST][PIPE_LUT + 8];
    alignas(16) u16 va[WST][NV];
    alignas(16) u16 vb[WST][NV];
    PipeSeq seq[PIPE_MAXSEQ];
    PipeTile tile[PIPE_LSTAGES];
    PipeDesc desc[WST];
    alignas(8) u64 fullL[PIPE_LSTAGES];
    alignas(8) u64 emptyL[PIPE_LSTAGES];
    alignas(8) u64 statsL[PIPE_LSTAGES];
    alignas(8) u64 fullW[WST];
    alignas(8) u64 emptyW[WST];
};
static_assert(sizeof(PipeSmemT<0>) <= 232448 && sizeof(PipeSmemT<1>) <= 232448 && sizeof(PipeSmemT<2>) <= 232448 &&
                  sizeof(PipeSmemT<3>) <= 232448,
              "the rings must fit the 227 KB of shared memory a CTA can opt in to");

#ifndef PIPE_WAITMODE
#define PIPE_WAITMODE 1  // A/B switch: 0 = poll with 64 ns sleeps everywhere; 1 = try_wait with a suspend-time hint; 2 = role-sized sleeps
#endif
// try_wait that lets the hardware park the warp for up to `ns` nanoseconds: no instructions are issued while it waits
__device__ __forceinline__ bool mbar_try_wait_hint(u64 *bar, u32 parity, u32 ns) {
    u32 ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// ROLE_NS: how long this role can oversleep without holding anyone up (loaders and surveyors run tiles ahead of the consumers).
// In the ncu profile of round 1's kernel 17 % of all issued instructions were the polling loops of idle producers.
template <u32 ROLE_NS>
__device__ __forceinline__ void pipe_wait(u64 *bar, u32 parity) {
#if PIPE_WAITMODE == 0
    mbar_wait(bar, parity);
#elif PIPE_WAITMODE == 1
    while (!mbar_try_wait_hint(bar, parity, 4000u)) {
    }
#else
    if (mbar_try_wait(bar, parity)) return;
    while (!mbar_try_wait(bar, parity)) __nanosleep(ROLE_NS);
#endif
}

__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

struct PipeOps {
    int k;
    int ops[GRB_MAX_OPS];
    u32 *ub, *pp, *lb;  // PIPE_UBLB
    int need_pfx;       // PIPE_UBLB / PIPE_MULTI: some column needs the score prefixes
    u32 *valid_bits;    // single-reduction kernels: validity as bits (bit g of word g / 32) instead of bytes
    u32 *stats;         // see TileOut::stats
    u32 stats_gen;
    uint4 *planar;      // see TileOut::planar
};

// #(keys < x) among the 8 keys of the two aligned quads at q
#ifndef PIPE_COUNT_PTX
#define PIPE_COUNT_PTX 0  // A/B switch: compare + predicated add per key in PTX (2 instructions; nvcc's own code takes 3: compare, add, un-add)
#endif
#ifndef PIPE_MERGE_RARE
#define PIPE_MERGE_RARE 0  // A/B switch: one branch for "some search of this thread met a crowded bin" instead of one per search
#endif
__device__ __forceinline__ u32 count4_below(const uint4 v, u32 x) {
#if PIPE_COUNT_PTX
    u32 c;
    asm("{\n"
        ".reg .pred p0, p1, p2, p3;\n"
        "setp.lt.u32 p0, %1, %5;\n"
        "setp.lt.u32 p1, %2, %5;\n"
        "setp.lt.u32 p2, %3, %5;\n"
        "setp.lt.u32 p3, %4, %5;\n"
        "mov.u32 %0, 0;\n"
        "@p0 add.u32 %0, %0, 1;\n"
        "@p1 add.u32 %0, %0, 1;\n"
        "@p2 add.u32 %0, %0, 1;\n"
        "@p3 add.u32 %0, %0, 1;\n"
        "}\n"
        : "=r"(c)
        : "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(x));
    return c;
#else
    return (v.x < x) + (v.y < x) + (v.z < x) + (v.w < x);
#endif
}
__device__ __forceinline__ u32 count8_below(const u32 *__restrict__ K, u32 q, u32 x) {
    const uint4 v0 = *reinterpret_cast<const uint4 *>(K + q), v1 = *reinterpret_cast<const uint4 *>(K + q + 4);
    return count4_below(v0, x) + count4_below(v1, x);
}

// flags: bit 0 = stage with TMA; bit 1 = debug, consumers skip the arithmetic (pipeline-only timing); bit 2 = debug, no stores
template <int FAST, int VAR>
__global__ void __launch_bounds__(PIPE_THREADS, 1) k_map_pipe(const SeqDev *__restrict__ seqs, int nseq, u32 tile_base, u32 ntiles,
                                                              const u32 *__restrict__ ls, const u32 *__restrict__ le,
                                                              double *__restrict__ out, u8 *__restrict__ out_valid, int flags,
                                                              u32 *__restrict__ ctx_flags, const PipeOps mops) {
    using Smem = PipeSmemT<VAR>;
    constexpr int WST = Smem::WST;
    constexpr bool NULLS = (VAR & PIPE_VAR_NULLS) != 0, WIDE = (VAR & PIPE_VAR_WIDE) != 0;
    const bool want_prefixes = FAST != GRB_OP_COUNT && ((FAST != PIPE_UBLB && FAST != PIPE_MULTI) || mops.need_pfx != 0);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool use_tma = flags & 1;
    if (tid == 0) {
        for (int i = 0; i < PIPE_LSTAGES; i++) {
            mbar_init(&S.fullL[i], 1);
            mbar_init(&S.emptyL[i], PIPE_NCG);
            mbar_init(&S.statsL[i], 1);
        }
        for (int i = 0; i < WST; i++) {
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
        q.lut16_a = sd->lut16_a;
        q.lut16_b = sd->lut16_b;
        q.vc16_a = sd->vc16_a;
        q.vc16_b = sd->vc16_b;
        q.pfx_a = sd->pfx_a;
        q.pfx_b = sd->pfx_b;
        q.pfh_a = sd->pfh_a;
        q.pfh_b = sd->pfh_b;
        q.nf_a = sd->nf_a;
        q.nf_b = sd->nf_b;
        q.left_begin = sd->left_begin;
        q.left_end = sd->left_end;
        q.scale = sd->scale;
        q.tile_begin = sd->tile_begin;
        q.bmax = sd->lut_bmax;
        q.fast_cmax = sd->fast_cmax;
        q.n = sd->n;
        q.sh = sd->lut_shift;
    }
    __syncthreads();
#if PIPE_ASSIGN_AB
    // tile assignment: round-robin (tile blockIdx.x + j gridDim.x), or -- flags bit 3, A/B switch GRB_PIPE_ASSIGN=1 -- runs of
    // PIPE_RUN consecutive tiles dealt round-robin (run r of CTA b = run r gridDim.x + b).  Measured: no gain (profiles/r2/ab_pipe_kernel.txt).
    const bool blocked = (flags & 8) != 0;
    u32 nmine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (blocked) {
        const u32 nruns = (ntiles + PIPE_RUN - 1) / PIPE_RUN;
        const u32 myruns = blockIdx.x < nruns ? (nruns - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        nmine = myruns * PIPE_RUN;
        if (myruns) {
            const u32 last = (blockIdx.x + (myruns - 1) * gridDim.x) * PIPE_RUN;  // first tile of my last run
            if (last + PIPE_RUN > ntiles) nmine -= last + PIPE_RUN - ntiles;
        }
    }
#else
    const u32 nmine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
#endif

    if (warp == 0) {
        // ================= lefts loader (A) =================
        const u64 pol = l2_evict_first_policy();
        const bool vec_in = ((((uintptr_t)ls) | ((uintptr_t)le)) & 15) == 0;
        int sq = 0;  // tiles only move forward, so does the seqname
        for (u32 j = 0; j < nmine; j++) {
            const u32 sl = j % PIPE_LSTAGES;
#if PIPE_ASSIGN_AB
            const u32 t = tile_base + (blocked ? (blockIdx.x + (j / PIPE_RUN) * gridDim.x) * PIPE_RUN + j % PIPE_RUN : blockIdx.x + j * gridDim.x);
#else
            const u32 t = tile_base + blockIdx.x + j * gridDim.x;
#endif
            while (sq + 1 < nseq && S.seq[sq + 1].tile_begin <= t) sq++;
            const PipeSeq &q = S.seq[sq];
            const u64 g_first = ((q.left_begin >> TILE_SHIFT) + (t - q.tile_begin)) << TILE_SHIFT;
            const u64 lo_g = q.left_begin > g_first ? q.left_begin : g_first;
            const u64 hi_g = q.left_end < g_first + TILE ? q.left_end : g_first + TILE;
            const bool whole = vec_in && lo_g == g_first && hi_g == g_first + TILE;
            pipe_wait<400>(&S.emptyL[sl], ((j / PIPE_LSTAGES) & 1) ^ 1);
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
                // ragged tile (seqname boundary, end of the arrays, unaligned pointers): plain copies, 16 loads of each
                // array in flight per lane (one at a time costs a memory round trip per element: ~30 us per ragged tile)
#pragma unroll 1
                for (u32 i0 = 0; i0 < TILE; i0 += 32 * 16) {
                    u32 xs[16], xe[16];
#pragma unroll
                    for (int r = 0; r < 16; r++) {
                        const u64 g = g_first + i0 + r * 32 + lane;
                        const bool ok = g >= lo_g && g < hi_g;
                        xs[r] = ok ? ls[g] : 0u;
                        xe[r] = ok ? le[g] : 0u;
                    }
#pragma unroll
                    for (int r = 0; r < 16; r++) {
                        S.ls[sl][i0 + r * 32 + lane] = xs[r];
                        S.le[sl][i0 + r * 32 + lane] = xe[r];
                    }
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
        u32 unstaged = 0;
        for (u32 j = 0; j < nmine; j++) {
            const u32 sw = j % WST, sl = j % PIPE_LSTAGES;
            pipe_wait<64>(&S.statsL[sl], (j / PIPE_LSTAGES) & 1);
            pipe_wait<64>(&S.emptyW[sw], ((j / WST) & 1) ^ 1);
            const PipeTile &ti = S.tile[sl];
            const PipeSeq &q = S.seq[ti.seq];
            constexpr u32 AM = ~(PipeLut<VAR>::AL - 1u);
            const u32 a0 = ti.bound[0] & AM, b0 = ti.bound[2] & AM, la0 = ti.bin[0] & AM, lb0 = ti.bin[2] & AM;
            // + 8: a search reads the two aligned quads around its answer, i.e. up to 7 keys past it
            const u32 alen = ti.bound[1] - a0 + 8, blen = ti.bound[3] - b0 + 8;
            const u32 lalen = ti.bin[1] + 1 - la0, lblen = ti.bin[3] + 1 - lb0;
            // bound[1] == 0xffffffff: the surveyor found invalid lefts in this tile (no index is that long)
            const bool staged = use_tma && ti.bound[1] != 0xffffffffu && alen <= PIPE_WIN && blen <= PIPE_WIN && lalen <= PIPE_LUT &&
                                lblen <= PIPE_LUT;
            const bool want_pfx = want_prefixes;
            // lane l issues copy l
            const void *src = nullptr;
            void *dst = nullptr;
            u32 bytes = 0;
            switch (lane) {
                case 0:
                    dst = S.la[sw];
                    if (PipeLut<VAR>::L16) { src = q.lut16_a + la0; bytes = lalen * 2; } else { src = q.lut_a + la0; bytes = lalen * 4; }
                    break;
                case 1:
                    dst = S.lb[sw];
                    if (PipeLut<VAR>::L16) { src = q.lut16_b + lb0; bytes = lblen * 2; } else { src = q.lut_b + lb0; bytes = lblen * 4; }
                    break;
                case 2: src = q.starts + a0; dst = S.ka[sw]; bytes = alen * 4; break;
                case 3: src = q.ends_sorted + b0; dst = S.kb[sw]; bytes = blen * 4; break;
                case 4: src = q.pfx_a + a0; dst = S.pa[sw]; bytes = want_pfx ? alen * 8 : 0; break;
                case 5: src = q.pfx_b + b0; dst = S.pb[sw]; bytes = want_pfx ? blen * 8 : 0; break;
                case 6: if (WIDE) { src = q.pfh_a + a0; dst = S.pah[sw]; bytes = want_pfx ? alen * 4 : 0; } break;
                case 7: if (WIDE) { src = q.pfh_b + b0; dst = S.pbh[sw]; bytes = want_pfx ? blen * 4 : 0; } break;
                case 8: if (NULLS) { src = q.vc16_a + a0; dst = S.va[sw]; bytes = alen * 2; } break;
                case 9: if (NULLS) { src = q.vc16_b + b0; dst = S.vb[sw]; bytes = blen * 2; } break;
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
                d.nf = ti.nf;
                if (staged)
                    mbar_arrive_expect_tx(&S.fullW[sw], total_bytes);
                else
                    mbar_arrive(&S.fullW[sw]);
            }
            __syncwarp();
            if (staged && bytes) tma_load_1d_hint(dst, src, bytes, &S.fullW[sw], pol);
            unstaged += staged ? 0u : 1u;
            __syncwarp();
        }
        // feedback for the caller's order detection: a plain store to a mapped host word, nothing on the stream
        if (lane == 0 && mops.stats != nullptr && blockIdx.x < 1000) {
            mops.stats[16 + blockIdx.x] = unstaged;
            if (blockIdx.x == 0) {
                __threadfence_system();
                mops.stats[1] = mops.stats_gen;
            }
        }
        return;
    }

    if (warp == 1 || (warp >= 3 && warp < 2 + PIPE_NSURV)) {
        // ================= surveyors =================
        const u32 sid = warp == 1 ? 0u : warp - 2;
        for (u32 j = sid; j < nmine; j += PIPE_NSURV) {
            const u32 sl = j % PIPE_LSTAGES;
            pipe_wait<300>(&S.fullL[sl], (j / PIPE_LSTAGES) & 1);
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
            const u32 bin_new = min(x >> q.sh, q.bmax);
            u32 bound_new = ((qq < 2) ? q.lut_a : q.lut_b)[bin_new + (qq & 1)];
            // a tile with invalid lefts must not be staged (their bins may fall outside the slices): an impossible
            // upper bound tells the window loader, and the consumers' global path skips those lefts
            if (tile_bad && qq == 1) bound_new = 0xffffffffu;
            if (lane < 4) {
                S.tile[sl].bin[lane] = bin_new;
                S.tile[sl].bound[lane] = bound_new;
            }
            if (NULLS) {
                // Non-finite scores are rare: every ub of the tile lies in [bound0, bound1] and every lb in [bound2, bound3], so
                // if the non-finite prefix counters agree at those four positions no group of the tile holds one, and the
                // consumers skip their per-left counter loads (a second round trip here, where there is slack, not there).
                u32 nf = 0;
                if (q.nf_a != nullptr) {
                    const u32 pos = min(bound_new, q.n);  // (an impossible bound marks a tile with invalid lefts)
                    const uint4 c4 = qq < 2 ? q.nf_a[pos] : q.nf_b[pos];
                    const u32 x0 = __shfl_sync(GRB_FULL, c4.x, 0), y0 = __shfl_sync(GRB_FULL, c4.y, 0), z0 = __shfl_sync(GRB_FULL, c4.z, 0);
                    nf = __any_sync(GRB_FULL, lane < 4 && (c4.x != x0 || c4.y != y0 || c4.z != z0)) || tile_bad ? 1u : 0u;
                }
                if (lane == 0) S.tile[sl].nf = nf;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.statsL[sl]);
        }
        return;
    }

    // ================= consumers =================
    const u32 cw = warp - 2 - PIPE_NSURV;
    const u32 i0 = (cw * 32 + lane) * PIPE_LPT;  // first slot of this thread inside a tile
    const bool ovec = ((((uintptr_t)out) & 15) | (((uintptr_t)out_valid) & 3)) == 0;
    const bool skip = flags & 2;
    const bool nostore = flags & 4;  // debug: do the arithmetic, keep the stores out (timing experiments)
    for (u32 j = 0; j < nmine; j++) {
        const u32 sw = j % WST, sl = j % PIPE_LSTAGES;
        pipe_wait<32>(&S.fullW[sw], (j / WST) & 1);
        if (skip) {
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.emptyW[sw]);
                mbar_arrive(&S.emptyL[sl]);
            }
            continue;
        }
        const PipeDesc &d = S.desc[sw];
        const PipeSeq &q = S.seq[d.seq];
        const u32 vlo = d.vlo, vhi = d.vhi;
        const int sh = q.sh;
        const u32 bmax = q.bmax, fast_cmax = q.fast_cmax;
        const double scale = q.scale;
        const uint4 *__restrict__ nf_a = q.nf_a, *__restrict__ nf_b = q.nf_b;
        const bool tile_nf = NULLS && nf_a != nullptr && d.nf != 0;
        u32 a[PIPE_LPT], b[PIPE_LPT];
        bool okv[PIPE_LPT];
        {
            const uint2 va = *reinterpret_cast<const uint2 *>(&S.ls[sl][i0]);
            const uint2 vb = *reinterpret_cast<const uint2 *>(&S.le[sl][i0]);
            a[0] = va.x, a[1] = va.y;
            b[0] = vb.x, b[1] = vb.y;
        }
#pragma unroll
        for (int k = 0; k < PIPE_LPT; k++) okv[k] = i0 + k >= vlo && i0 + k < vhi;
        // (A left with end <= start or end > 2^31-1 -- impossible in the reference, which panics -- was flagged by
        // the surveyor, which also keeps its tile off the staged path.)
        const u32 a0 = d.a0, b0 = d.b0, la0 = d.la0, lb0 = d.lb0;
        const SeqDev *__restrict__ sd = &seqs[d.seq];  // only dereferenced on the rare paths
        u32 ub[PIPE_LPT], lbk[PIPE_LPT], ppk[PIPE_LPT];  // absolute positions
        u64 pa[PIPE_LPT], pb[PIPE_LPT];
        u32 pha[PIPE_LPT], phb[PIPE_LPT];  // WIDE: bits 64..95 of the prefixes
        u32 nv16[PIPE_LPT];                // NULLS: (valid-score count of the group) mod 2^16
#pragma unroll
        for (int k = 0; k < PIPE_LPT; k++) ppk[k] = pha[k] = phb[k] = nv16[k] = 0;
        if (d.staged) {
            const u32 *__restrict__ KA = S.ka[sw];
            const u32 *__restrict__ KB = S.kb[sw];
            const u32 lmask = PipeLut<VAR>::L16 ? 0xffffu : 0xffffffffu;
            const u32 a16 = a0 & lmask, b16 = b0 & lmask;
            u32 pA[PIPE_LPT], pB[PIPE_LPT], cA[PIPE_LPT], cB[PIPE_LPT];
            // bin -> first position of the bin relative to the staged slice (a window holds < 65536 keys, so the low
            // halves of the table entries are enough); invalid slots scan with x = 0
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++) {
                const u32 ba = min(b[k] >> sh, bmax), bb = min((a[k] + 1) >> sh, bmax);
                pA[k] = okv[k] ? ((u32)S.la[sw][ba - la0] - a16) & lmask : 0u;
                pB[k] = okv[k] ? ((u32)S.lb[sw][bb - lb0] - b16) & lmask : 0u;
            }
            // one unconditional round per search: the 8 keys of the two aligned quads around the bin's first
            // position, two 128-bit loads.  Keys before that position are < x by construction, so the answer
            // is (quad base) + #(keys < x) unless all 8 are smaller (a crowded bin: finish with the scan).
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++) {
                const u32 xa = okv[k] ? b[k] : 0u, xb = okv[k] ? a[k] + 1 : 0u;
                const u32 qa = pA[k] & ~3u, qb = pB[k] & ~3u;
                cA[k] = count8_below(KA, qa, xa);
                cB[k] = count8_below(KB, qb, xb);
                pA[k] = okv[k] ? qa + cA[k] : 0u;
                pB[k] = okv[k] ? qb + cB[k] : 0u;
            }
            if (FAST == PIPE_UBLB) {
                // third search: #{start < l.start}, same scheme on the starts window
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) {
                    const u32 xp = okv[k] ? a[k] : 0u;
                    u32 pP = okv[k] ? ((u32)S.la[sw][min(a[k] >> sh, bmax) - la0] - a16) & lmask : 0u;
                    const u32 qp = pP & ~3u;
                    const u32 cP = count8_below(KA, qp, xp);
                    pP = okv[k] ? qp + cP : 0u;
                    if (cP == 8) pP = scan_lower_bound(KA, pP, a[k]);
                    ppk[k] = a0 + pP;
                }
            }
#if PIPE_MERGE_RARE
            if (((cA[0] | cB[0] | cA[1] | cB[1]) & 8u) != 0) {  // rare: some search met a crowded bin, the scan goes on
#pragma unroll
                for (int k = 0; k < PIPE_LPT; k++) {
                    if (cA[k] == 8) pA[k] = scan_lower_bound(KA, pA[k], b[k]);
                    if (cB[k] == 8) pB[k] = scan_lower_bound(KB, pB[k], a[k] + 1);
                }
            }
#endif
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++) {
#if !PIPE_MERGE_RARE
                if (cA[k] == 8) pA[k] = scan_lower_bound(KA, pA[k], b[k]);      // rare: a crowded bin
                if (cB[k] == 8) pB[k] = scan_lower_bound(KB, pB[k], a[k] + 1);
#endif
                pa[k] = want_prefixes ? S.pa[sw][pA[k]] : 0ull;
                pb[k] = want_prefixes ? S.pb[sw][pB[k]] : 0ull;
                if (WIDE && want_prefixes) {
                    pha[k] = S.pah[sw][pA[k]];
                    phb[k] = S.pbh[sw][pB[k]];
                }
                if (NULLS) nv16[k] = ((u32)S.va[sw][pA[k]] - (u32)S.vb[sw][pB[k]]) & 0xffffu;
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
                    if (WIDE) {
                        pha[k] = q.pfh_a[ub[k]];
                        phb[k] = q.pfh_b[lbk[k]];
                    }
                }
                if (NULLS) nv16[k] = ((u32)q.vc16_a[ub[k]] - (u32)q.vc16_b[lbk[k]]) & 0xffffu;
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
        if constexpr (VAR == 0 && FAST != PIPE_MULTI && FAST != PIPE_UBLB) {
            // The headline path (every score Some and finite, 64-bit prefixes, one reduction, validity as bytes), kept
            // exactly as it was tuned in round 1: small rearrangements of this block cost 2-5 % (profiles/r2/ab_*.txt).
            if (mops.valid_bits == nullptr) {
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
                } else if (ovec && okv[0] && okv[PIPE_LPT - 1]) {
                    double2 *po = reinterpret_cast<double2 *>(out + g0);
                    __stcs(po, make_double2(r[0], r[1]));  // streaming stores: results are not read back here
                    __stcs(reinterpret_cast<uchar2 *>(out_valid + g0), make_uchar2(v[0], v[1]));
                } else {
#pragma unroll
                    for (int k = 0; k < PIPE_LPT; k++) {
                        if (okv[k]) {
                            out[g0 + k] = r[k];
                            out_valid[g0 + k] = v[k];
                        }
                    }
                }
                continue;
            }
        }
        // ---- count, number of members that carry a score, exact sum
        u32 cnt[PIPE_LPT], nval[PIPE_LPT];
        double sum[PIPE_LPT];
#pragma unroll
        for (int k = 0; k < PIPE_LPT; k++) {
            const u32 c = ub[k] - lbk[k];
            cnt[k] = c;
            nval[k] = c;
            if (NULLS) nval[k] = c < 65536u ? nv16[k] : (okv[k] ? sd->vcnt_a[ub[k]] - sd->vcnt_b[lbk[k]] : 0u);
            sum[k] = 0.0;
            if (!want_prefixes) continue;
            if (!WIDE) {
                if (c < fast_cmax) {
                    sum[k] = (double)(i64)(pa[k] - pb[k]) * scale;
                } else {
                    const i128 ba128 = sd->base_a[ub[k] >> GRB_BASE_SHIFT], bb128 = sd->base_b[lbk[k] >> GRB_BASE_SHIFT];
                    const i128 PA = ba128 + (i128)(i64)(pa[k] - (u64)(u128)ba128);
                    const i128 PB = bb128 + (i128)(i64)(pb[k] - (u64)(u128)bb128);
                    sum[k] = fx_to_double(PA - PB, sd->e0);
                }
            } else {
                if (c < fast_cmax) {
                    const u64 lo = pa[k] - pb[k];
                    const u32 hi = pha[k] - phb[k] - (pa[k] < pb[k] ? 1u : 0u);
                    const i64 lo_s = (i64)lo;
                    // most sums of modest groups fit 63 bits even in the wide format: one conversion instead of the 128-bit path
                    if (hi == (u32)(lo_s >> 63))
                        sum[k] = (double)lo_s * scale;
                    else
                        sum[k] = fx_to_double(sext96(hi, lo), sd->e0);
                } else {
                    const u128 m96 = (((u128)1) << 96) - 1;
                    const i128 ba128 = sd->base_a[ub[k] >> GRB_BASE_SHIFT], bb128 = sd->base_b[lbk[k] >> GRB_BASE_SHIFT];
                    const u128 da = ((((u128)pha[k] << 64) | pa[k]) - (u128)ba128) & m96;
                    const u128 db = ((((u128)phb[k] << 64) | pb[k]) - (u128)bb128) & m96;
                    const i128 PA = ba128 + sext96((u32)(da >> 64), (u64)da);
                    const i128 PB = bb128 + sext96((u32)(db >> 64), (u64)db);
                    sum[k] = fx_to_double(PA - PB, sd->e0);
                }
            }
            if (NULLS && tile_nf && okv[k]) {
                const uint4 na = nf_a[ub[k]], nb = nf_b[lbk[k]];
                sum[k] = nf_rule(sum[k], na.x - nb.x, na.y - nb.y, na.z - nb.z);
            }
        }
        // ---- outputs
        if (FAST == PIPE_UBLB) {
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++)
                if (okv[k]) {
                    mops.ub[g0 + k] = ub[k];
                    mops.pp[g0 + k] = ppk[k];
                    mops.lb[g0 + k] = lbk[k];
                }
        }
        if (FAST == PIPE_UBLB && mops.planar != nullptr) {
            // the member sweep that follows writes the rows: hand it (sum, count, scored members), one 16-byte store per left
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++)
                if (okv[k]) {
                    const u64 sb = (u64)__double_as_longlong(sum[k]);
                    __stcs(mops.planar + g0 + k, make_uint4((u32)sb, (u32)(sb >> 32), cnt[k], nval[k]));
                }
            continue;
        }
        if (FAST == PIPE_MULTI || FAST == PIPE_UBLB) {
            // several reductions from the same (count, exact sum): row-major out[left][column]
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++) {
                if (!okv[k]) continue;
                const double mean = nval[k] ? sum[k] / (double)nval[k] : 0.0;
                double *po = out + (g0 + k) * (u64)mops.k;
                u8 *pv = out_valid + (g0 + k) * (u64)mops.k;
                for (int col = 0; col < mops.k; col++) {
                    const int op = mops.ops[col];
                    if (FAST == PIPE_UBLB && op != GRB_OP_COUNT && op != GRB_OP_MEAN && op != GRB_OP_SUM && op != GRB_OP_SUM_NOT_EMPTY)
                        continue;  // a member reduction / closed form: the kernels that follow own this column
                    const double rr = op == GRB_OP_COUNT ? (double)cnt[k] : op == GRB_OP_MEAN ? mean : sum[k];
                    const u8 vv = (op == GRB_OP_SUM || op == GRB_OP_COUNT) ? (u8)1 : (u8)(nval[k] != 0);
                    __stcs(po + col, rr);
                    pv[col] = vv;
                }
            }
            continue;
        }
        double r[PIPE_LPT];
        bool v[PIPE_LPT];
#pragma unroll
        for (int k = 0; k < PIPE_LPT; k++) {
            if (FAST == GRB_OP_COUNT) {
                r[k] = (double)cnt[k];
                v[k] = true;
            } else if (FAST == GRB_OP_SUM) {
                r[k] = sum[k];
                v[k] = true;
            } else if (FAST == GRB_OP_SUM_NOT_EMPTY) {
                r[k] = sum[k];
                v[k] = nval[k] != 0;
            } else {
                v[k] = nval[k] != 0;
                r[k] = v[k] ? sum[k] / (double)nval[k] : 0.0;
            }
        }
        if (nostore && r[0] != 1.2345e300) continue;
        if (mops.valid_bits != nullptr) {
            // validity as bits.  Lane L owns slots 2 L, 2 L + 1: lanes 0..15 fill the first word of the warp's 64 slots,
            // lanes 16..31 the second.  (Measured at 100M x 100M: the two warp reductions cost more issue slots than the
            // byte stores they replace -- 0.94 vs 0.88 ms -- so the byte form stays the default.)
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++)
                if (okv[k]) out[g0 + k] = r[k];
            const u32 sh2 = 2 * (lane & 15);
            const u32 lv = ((okv[0] ? 1u : 0u) | (okv[1] ? 2u : 0u)) << sh2;
            const u32 bv = ((okv[0] && v[0] ? 1u : 0u) | (okv[1] && v[1] ? 2u : 0u)) << sh2;
            const u32 live0 = __reduce_or_sync(GRB_FULL, lane < 16 ? lv : 0u), live1 = __reduce_or_sync(GRB_FULL, lane < 16 ? 0u : lv);
            const u32 bits0 = __reduce_or_sync(GRB_FULL, lane < 16 ? bv : 0u), bits1 = __reduce_or_sync(GRB_FULL, lane < 16 ? 0u : bv);
            if (lane < 2) {
                const u32 live = lane ? live1 : live0, bits = lane ? bits1 : bits0;
                u32 *w = mops.valid_bits + ((d.g_first + cw * 64) >> 5) + lane;  // tiles start on multiples of 1024 lefts
                if (live == GRB_FULL) {
                    __stcs(w, bits);
                } else if (live) {
                    // a seqname boundary inside the word: the neighbouring tile owns the other bits
                    atomicAnd(w, ~live);
                    atomicOr(w, bits);
                }
            }
        } else if (okv[0] && okv[1] && ((((uintptr_t)out) & 15) | (((uintptr_t)out_valid) & 1)) == 0) {
            __stcs(reinterpret_cast<double2 *>(out + g0), make_double2(r[0], r[1]));  // streaming stores: results are not read back here
            __stcs(reinterpret_cast<uchar2 *>(out_valid + g0), make_uchar2(v[0] ? 1 : 0, v[1] ? 1 : 0));
        } else {
#pragma unroll
            for (int k = 0; k < PIPE_LPT; k++)
                if (okv[k]) {
                    out[g0 + k] = r[k];
                    out_valid[g0 + k] = v[k] ? (u8)1 : (u8)0;
                }
        }
    }
}

template <int FAST, int VAR>
int launch_var(grb_ctx *ctx, const Batch &b, int s0, int s1, const u32 *ls, const u32 *le, const TileOut &o, u32 *valid_bits) {
    using Smem = PipeSmemT<VAR>;
    // the opt-in to > 48 KB of dynamic shared memory is per device: remembered per device, not per process
    static bool configured[64] = {};
    if (ctx->device < 0 || ctx->device >= 64 || !configured[ctx->device]) {
        GRB_CUDA(ctx, cudaFuncSetAttribute(k_map_pipe<FAST, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        if (ctx->device >= 0 && ctx->device < 64) configured[ctx->device] = true;
    }
    const int flags = (ctx->tile_mode ? 1 : 0) | (ctx->pipe_mode == 2 ? 2 : 0) | (ctx->pipe_mode == 3 ? 4 : 0) | (ctx->pipe_assign ? 8 : 0);
    const u32 t0 = b.host_tile_begin[s0], t1 = b.host_tile_begin[s1];
    if (t1 == t0) return GRB_OK;
    u32 grid = (u32)ctx->sm_count;
    if (grid > t1 - t0) grid = t1 - t0;
    PipeOps mops;
    mops.k = o.k;
    for (int c = 0; c < GRB_MAX_OPS; c++) mops.ops[c] = c < o.k ? o.ops[c] : 0;
    mops.ub = o.ub;
    mops.pp = o.pp;
    mops.lb = o.lb;
    mops.need_pfx = o.need_sum;
    mops.valid_bits = valid_bits;
    mops.stats = o.stats;
    mops.stats_gen = o.stats_gen;
    mops.planar = o.planar;
    k_map_pipe<FAST, VAR><<<grid, PIPE_THREADS, sizeof(Smem), ctx->stream>>>(b.seqs + s0, s1 - s0, t0, t1 - t0, ls, le, o.out, o.out_valid, flags,
                                                                          o.flags, mops);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}

template <int FAST>
int launch_fast(grb_ctx *ctx, const Batch &b, int s0, int s1, int var, const u32 *ls, const u32 *le, const TileOut &o, u32 *valid_bits) {
    switch (var) {
        case 0: return launch_var<FAST, 0>(ctx, b, s0, s1, ls, le, o, valid_bits);
        case 1: return launch_var<FAST, 1>(ctx, b, s0, s1, ls, le, o, valid_bits);
        case 2: return launch_var<FAST, 2>(ctx, b, s0, s1, ls, le, o, valid_bits);
        default: return launch_var<FAST, 3>(ctx, b, s0, s1, ls, le, o, valid_bits);
    }
}

}  // namespace

int grb_launch_pipe(grb_ctx *ctx, const Batch &b, const u32 *ls, const u32 *le, const TileOut &o, int fast, u32 *valid_bits) {
    if (b.ntiles == 0) return GRB_OK;
    // Runs of consecutive seqnames that need the same variant (absent / empty seqnames fit any), at most PIPE_MAXSEQ
    // per launch: the kernel keeps its seqnames' descriptors in shared memory.
    int s0 = 0;
    while (s0 < b.nseq) {
        int var = -1, s1 = s0;
        while (s1 < b.nseq && s1 - s0 < PIPE_MAXSEQ) {
            const int v = b.var[s1];
            if (v >= 0 && var >= 0 && v != var) break;
            if (v >= 0) var = v;
            s1++;
        }
        if (var < 0) var = 0;
        // counts alone need no score planes at all
        if (fast == GRB_OP_COUNT || ((fast == PIPE_UBLB || fast == PIPE_MULTI) && !o.need_sum)) var = 0;
        int rc;
        switch (fast) {
            case GRB_OP_MEAN: rc = launch_fast<GRB_OP_MEAN>(ctx, b, s0, s1, var, ls, le, o, valid_bits); break;
            case GRB_OP_SUM: rc = launch_fast<GRB_OP_SUM>(ctx, b, s0, s1, var, ls, le, o, valid_bits); break;
            case GRB_OP_SUM_NOT_EMPTY: rc = launch_fast<GRB_OP_SUM_NOT_EMPTY>(ctx, b, s0, s1, var, ls, le, o, valid_bits); break;
            case GRB_OP_COUNT: rc = launch_var<GRB_OP_COUNT, 0>(ctx, b, s0, s1, ls, le, o, valid_bits); break;
            case PIPE_MULTI: rc = launch_fast<PIPE_MULTI>(ctx, b, s0, s1, var, ls, le, o, nullptr); break;
            case PIPE_UBLB: rc = launch_fast<PIPE_UBLB>(ctx, b, s0, s1, var, ls, le, o, nullptr); break;
            default: return grb_fail(ctx, GRB_ERR_INVALID_ARG, "pipe: unknown reduction %d", fast);
        }
        if (rc != GRB_OK) return rc;
        s0 = s1;
    }
    return GRB_OK;
}
