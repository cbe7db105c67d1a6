// sort.cu -- stable LSD radix sort of (u64 key, u32 value) pairs, hand-written.
//
// Index build only (replaces the sort inside coitrees' BasicCOITree::new, reached from
// src/ranges/coitrees.rs:78-84).  8-bit digits; one pre-pass finds which digits vary at all
// (coordinates < 2^28 leave most of the 8 digits constant) and whether the keys are already
// sorted (BED files usually are), so typical inputs pay for 0-4 scatter passes.
//
// Per pass: tile histograms (digit-major) -> decoupled look-back scan -> stable scatter.  In
// the scatter a warp ranks its keys with match_any (all lanes holding the same digit elect a
// leader that bumps the warp's shared-memory counter once), which keeps equal digits in input
// order without atomics.
#include "common.cuh"

namespace {

constexpr int SORT_TPB = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_TPB * SORT_ITEMS;
constexpr int SORT_WARPS = SORT_TPB / 32;

struct KeyStats {
    u64 or_all;
    u64 and_all;
    u32 unsorted;
    u32 pad;
};

template <typename K>
__global__ void __launch_bounds__(256) k_key_stats(const K *__restrict__ keys, size_t n, KeyStats *st) {
    __shared__ u64 s_or[8], s_and[8];
    __shared__ u32 s_bad[8];
    u64 o = 0, a = ~0ull;
    u32 bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        u64 k = (u64)keys[i];
        o |= k;
        a &= k;
        if (i + 1 < n && keys[i + 1] < k) bad = 1;
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        o |= __shfl_xor_sync(GRB_FULL, o, s);
        a &= __shfl_xor_sync(GRB_FULL, a, s);
        bad |= __shfl_xor_sync(GRB_FULL, bad, s);
    }
    // one set of atomics per block, not per warp: hundreds of thousands of atomics on three words serialise
    if (lane_id() == 0) {
        s_or[threadIdx.x >> 5] = o;
        s_and[threadIdx.x >> 5] = a;
        s_bad[threadIdx.x >> 5] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            o |= s_or[w];
            a &= s_and[w];
            bad |= s_bad[w];
        }
        atomicOr((unsigned long long *)&st->or_all, (unsigned long long)o);
        atomicAnd((unsigned long long *)&st->and_all, (unsigned long long)a);
        if (bad) atomicOr(&st->unsorted, 1u);
    }
}

template <typename K>
__global__ void __launch_bounds__(SORT_TPB) k_sort_hist(const K *__restrict__ keys, size_t n, int shift, u32 ntiles,
                                                        u32 *__restrict__ hist /* [256][ntiles] */) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        size_t g = base + (size_t)i * SORT_TPB + threadIdx.x;
        if (g < n) atomicAdd(&h[(keys[g] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

template <typename K>
__global__ void k_copy_pairs(const K *__restrict__ keys, const u32 *__restrict__ vals, size_t n, K *__restrict__ keys_out, u32 *__restrict__ vals_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys_out[i] = keys[i];
    vals_out[i] = vals ? vals[i] : (u32)i;
}

// vals == nullptr: the values are the positions 0..n-1 (first pass of an argsort: nothing to read)
template <typename K>
__global__ void __launch_bounds__(SORT_TPB) k_scatter(const K *__restrict__ keys, const u32 *__restrict__ vals, size_t n,
                                                      int shift, u32 ntiles, const u32 *__restrict__ offs /* [256][ntiles] */,
                                                      K *__restrict__ keys_out, u32 *__restrict__ vals_out) {
    __shared__ u32 wcnt[SORT_WARPS][256];
    for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_TPB) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1;
    // warp w owns keys [base + w*256, base + (w+1)*256); item i of lane l is element i*32 + l of that chunk
    const size_t wbase = (size_t)blockIdx.x * SORT_TILE + (size_t)warp * (32 * SORT_ITEMS);
    K k[SORT_ITEMS];
    u32 v[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        size_t g = wbase + (size_t)i * 32 + lane;
        bool ok = g < n;
        k[i] = ok ? keys[g] : 0;
        v[i] = ok ? (vals ? vals[g] : (u32)g) : 0;
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        size_t g = wbase + (size_t)i * 32 + lane;
        bool ok = g < n;
        u32 d = (u32)(k[i] >> shift) & 0xff;
        u32 okmask = __ballot_sync(GRB_FULL, ok);
        u32 peers = __match_any_sync(GRB_FULL, d) & okmask;
        int leader = ok ? __ffs(peers) - 1 : 0;
        u32 old = 0;
        if (ok && (int)lane == leader) {
            old = wcnt[warp][d];
            wcnt[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(GRB_FULL, old, leader);
        rank[i] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread d turns the per-warp counts of digit d into global bases
        u32 d = threadIdx.x;
        u32 run = offs[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; w++) {
            u32 c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        size_t g = wbase + (size_t)i * 32 + lane;
        if (g < n) {
            u32 d = (u32)(k[i] >> shift) & 0xff;
            u32 pos = wcnt[warp][d] + rank[i];
            keys_out[pos] = k[i];
            vals_out[pos] = v[i];
        }
    }
}

}  // namespace

namespace {

// Stable LSD radix sort of (key, value) pairs.  `in_keys` is only read (vals_in == nullptr: values 0..n-1).  The passes
// ping-pong between (dst_k, dst_v) and (tmp_k, tmp_v) in such an order that the last pass lands in dst; when the keys are
// already sorted, or do not vary, one copy pass puts them there.  *was_sorted (may be NULL) reports sorted input.
template <typename K>
int sort_into(grb_ctx *ctx, const K *in_keys, const u32 *vals_in, K *dst_k, u32 *dst_v, K *tmp_k, u32 *tmp_v, size_t n, int key_bits,
              bool *was_sorted, bool copy_if_sorted, bool *ended_in_dst = nullptr) {
    if (was_sorted) *was_sorted = true;
    if (n == 0) return GRB_OK;
    if (n >= 0xfffffff0ull) return grb_fail(ctx, GRB_ERR_UNSUPPORTED, "sort: more than 2^32-16 ranges on one seqname");
    TempBuf stb;
    GRB_TRY(stb.alloc(ctx, sizeof(KeyStats)));
    KeyStats init = {0ull, ~0ull, 0u, 0u};
    GRB_CUDA(ctx, cudaMemcpyAsync(stb.p, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    int grid = (int)((n + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    k_key_stats<K><<<grid, 256, 0, ctx->stream>>>(in_keys, n, stb.as<KeyStats>());
    GRB_LAUNCHED(ctx);
    KeyStats st;
    GRB_TRY(grb_read_small(ctx, &st, stb.p, sizeof(st)));
    const u64 varying = st.or_all ^ st.and_all;
    int shifts[8], npass = 0;
    if (st.unsorted)
        for (int pass = 0; pass * 8 < key_bits; pass++)
            if ((varying >> (pass * 8)) & 0xff) shifts[npass++] = pass * 8;  // (the other digits are the same in every key)
    if (npass == 0) {
        // already in order
        if (copy_if_sorted) {
            k_copy_pairs<K><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(in_keys, vals_in, n, dst_k, dst_v);
            GRB_LAUNCHED(ctx);
        }
        return GRB_OK;
    }
    if (was_sorted) *was_sorted = false;
    const u32 ntiles = (u32)((n + SORT_TILE - 1) / SORT_TILE);
    TempBuf hist, offs;
    GRB_TRY(hist.alloc(ctx, (size_t)256 * ntiles * sizeof(u32)));
    GRB_TRY(offs.alloc(ctx, (size_t)256 * ntiles * sizeof(u32)));
    const K *kin = in_keys;
    const u32 *vin = vals_in;
    // the last pass must write dst: an odd number of passes starts there.  (ended_in_dst != NULL: the scratch pair IS the
    // input -- always start in dst, and tell the caller where the passes ended.)
    bool to_dst = ended_in_dst ? true : (npass & 1) != 0;
    if (ended_in_dst) *ended_in_dst = (npass & 1) != 0;
    for (int p = 0; p < npass; p++) {
        K *kout = to_dst ? dst_k : tmp_k;
        u32 *vout = to_dst ? dst_v : tmp_v;
        k_sort_hist<K><<<ntiles, SORT_TPB, 0, ctx->stream>>>(kin, n, shifts[p], ntiles, hist.as<u32>());
        GRB_LAUNCHED(ctx);
        GRB_TRY(grb_scan_u32_to_u32(ctx, hist.as<u32>(), offs.as<u32>(), (size_t)256 * ntiles, false));
        k_scatter<K><<<ntiles, SORT_TPB, 0, ctx->stream>>>(kin, vin, n, shifts[p], ntiles, offs.as<u32>(), kout, vout);
        GRB_LAUNCHED(ctx);
        kin = kout;
        vin = vout;
        to_dst = !to_dst;
    }
    return GRB_OK;
}

}  // namespace

int grb_sort_pairs(grb_ctx *ctx, u64 *keys, u32 *vals, u64 *keys_alt, u32 *vals_alt, size_t n, u64 **keys_out,
                   u32 **vals_out, bool *was_sorted) {
    // in place when sorted; otherwise the result lands in the alternate buffers (keys / vals serve as the scratch pair)
    *keys_out = keys;
    *vals_out = vals;
    bool sorted = true, in_alt = false;
    if (n > 1) GRB_TRY(sort_into<u64>(ctx, keys, vals, keys_alt, vals_alt, keys, vals, n, 64, &sorted, false, &in_alt));
    if (was_sorted) *was_sorted = sorted;
    if (!sorted && in_alt) {
        *keys_out = keys_alt;
        *vals_out = vals_alt;
    }
    return GRB_OK;
}

int grb_argsort_u32(grb_ctx *ctx, const u32 *keys, u32 *keys_sorted, u32 *perm, u32 *tmp_keys, u32 *tmp_vals, size_t n) {
    return sort_into<u32>(ctx, keys, nullptr, keys_sorted, perm, tmp_keys, tmp_vals, n, 32, nullptr, true);
}
