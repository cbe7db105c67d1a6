// scan.cu -- single-pass exclusive prefix scan with decoupled look-back.
//
// Used for (a) the CSR offsets of left_overlaps (count pass -> scan -> emit pass,
// replacing the per-left Vec growth of src/granges.rs:958-969) and (b) the digit
// offsets of the radix sort in the index build.  Each tile publishes
// {flag, value} packed in ONE 64-bit word (2 flag bits + 62 value bits), so a
// single relaxed 64-bit load observes a consistent pair and no fence is needed
// between flag and value.
#include "common.cuh"

namespace {

constexpr int SCAN_TPB = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_TPB * SCAN_ITEMS;

constexpr u64 FLAG_AGG = 1ull << 62;  // tile aggregate available
constexpr u64 FLAG_PFX = 2ull << 62;  // inclusive prefix available
constexpr u64 VAL_MASK = (1ull << 62) - 1;

template <typename TOut>
__global__ void __launch_bounds__(SCAN_TPB) k_scan_lookback(const u32 *__restrict__ in, TOut *__restrict__ out, size_t n,
                                                            u64 *__restrict__ tile_state, u32 *__restrict__ tile_counter,
                                                            int write_total) {
    __shared__ u32 s_tile;
    __shared__ u64 s_warp[SCAN_TPB / 32];
    __shared__ u64 s_excl;
    // dynamic tile id: tiles start in id order, so every predecessor is already running
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const size_t base = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;

    u32 v[SCAN_ITEMS];
    u64 tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t g = base + i;
        v[i] = g < n ? in[g] : 0u;
        tsum += v[i];
    }
    // block exclusive scan of thread sums
    u64 incl = tsum;
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(GRB_FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    u64 warp_off = 0, tile_agg = 0;
#pragma unroll
    for (int w = 0; w < SCAN_TPB / 32; w++) {
        u64 t = s_warp[w];
        if (w < (int)warp) warp_off += t;
        tile_agg += t;
    }
    u64 thread_excl = warp_off + incl - tsum;

    // publish + look back (warp 0)
    if (warp == 0) {
        u64 excl = 0;
        if (tile == 0) {
            if (lane == 0) st_relaxed_u64(&tile_state[0], FLAG_PFX | tile_agg);
        } else {
            if (lane == 0) st_relaxed_u64(&tile_state[tile], FLAG_AGG | tile_agg);
            long long look = (long long)tile - 1;
            while (true) {
                long long idx = look - lane;
                u64 s = FLAG_PFX;  // virtual tiles before 0: prefix 0
                if (idx >= 0) {
                    do {
                        s = ld_relaxed_u64(&tile_state[idx]);
                    } while ((s >> 62) == 0);
                }
                u32 pfx_mask = __ballot_sync(GRB_FULL, (s >> 62) == 2);
                u64 val = s & VAL_MASK;
                if (pfx_mask) {
                    int first = __ffs(pfx_mask) - 1;  // nearest tile holding an inclusive prefix
                    u64 contrib = (int)lane <= first ? val : 0;
                    excl += warp_sum64(contrib);
                    break;
                }
                excl += warp_sum64(val);
                look -= 32;
            }
            if (lane == 0) st_relaxed_u64(&tile_state[tile], FLAG_PFX | ((excl + tile_agg) & VAL_MASK));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    u64 run = s_excl + thread_excl;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        size_t g = base + i;
        if (g < n) out[g] = (TOut)run;
        run += v[i];
        if (write_total && g + 1 == n) out[n] = (TOut)run;
    }
}

template <typename TOut>
int scan_impl(grb_ctx *ctx, const u32 *in, TOut *out, size_t n, bool write_total) {
    if (n == 0) {
        if (write_total) GRB_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(TOut), ctx->stream));
        return GRB_OK;
    }
    size_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    TempBuf state;
    GRB_TRY(state.alloc(ctx, (ntiles + 1) * sizeof(u64)));
    GRB_CUDA(ctx, cudaMemsetAsync(state.p, 0, (ntiles + 1) * sizeof(u64), ctx->stream));
    u64 *tile_state = state.as<u64>();
    u32 *counter = (u32 *)(tile_state + ntiles);
    k_scan_lookback<TOut><<<(unsigned)ntiles, SCAN_TPB, 0, ctx->stream>>>(in, out, n, tile_state, counter, write_total ? 1 : 0);
    GRB_LAUNCHED(ctx);
    return GRB_OK;
}

}  // namespace

int grb_scan_u32_to_u64(grb_ctx *ctx, const u32 *in, u64 *out, size_t n, bool write_total) {
    return scan_impl<u64>(ctx, in, out, n, write_total);
}
int grb_scan_u32_to_u32(grb_ctx *ctx, const u32 *in, u32 *out, size_t n, bool write_total) {
    return scan_impl<u32>(ctx, in, out, n, write_total);
}
