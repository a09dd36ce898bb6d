// sort_scan.cuh -- hand-written device-wide primitives (no CUB / Thrust):
//   * exclusive_scan_u32 : 3-phase tile scan (reduce, scan of tile sums (recursive), apply)
//   * radix_sort_pairs   : stable LSD radix sort, 8-bit digits, (key, uint32 payload), keys 32 or 64 bit
// Both are HBM-streaming kernels: grid = one CTA per 4096-element tile, 256 threads, fully coalesced
// loads (striped arrangement); the only on-chip state is the per-warp digit histogram in shared memory.
#pragma once
#include "common.cuh"

namespace nbk {

constexpr int PRIM_THREADS = 256;
constexpr int PRIM_ITEMS = 16;
constexpr int PRIM_TILE = PRIM_THREADS * PRIM_ITEMS;  // 4096

// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan of one value per thread (256 threads); returns exclusive prefix, total in *tot
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* swarp /*[8]*/, uint32_t* tot) {
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) swarp[w] = inc;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t c = swarp[i];
        if (i < (int)w) woff += c;
        total += c;
    }
    __syncthreads();
    *tot = total;
    return woff + inc - v;
}

// ---------------------------------------------------------------------------------------------
// generic u32 exclusive scan.  Blocked arrangement through shared memory so global access is coalesced.
static __global__ void __launch_bounds__(PRIM_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in, int64_t n, uint32_t* __restrict__ tsum) {
    __shared__ uint32_t swarp[8];
    int64_t base = (int64_t)blockIdx.x * PRIM_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int64_t i = base + r * PRIM_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    block_excl_scan_256(s, swarp, &tot);
    if (threadIdx.x == 0) tsum[blockIdx.x] = tot;
}

static __global__ void __launch_bounds__(PRIM_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in, int64_t n, const uint32_t* __restrict__ toff,
                                                                  uint32_t* __restrict__ out) {
    __shared__ uint32_t sdata[PRIM_TILE + PRIM_TILE / 32];  // padded: index i -> i + i/32
    __shared__ uint32_t swarp[8];
    int64_t base = (int64_t)blockIdx.x * PRIM_TILE;
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int li = r * PRIM_THREADS + threadIdx.x;
        int64_t i = base + li;
        sdata[li + (li >> 5)] = (i < n) ? in[i] : 0u;
    }
    __syncthreads();
    uint32_t v[PRIM_ITEMS], s = 0;
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int li = threadIdx.x * PRIM_ITEMS + r;
        v[r] = sdata[li + (li >> 5)];
        s += v[r];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan_256(s, swarp, &tot) + (toff ? toff[blockIdx.x] : 0u);
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int li = threadIdx.x * PRIM_ITEMS + r;
        sdata[li + (li >> 5)] = ex;
        ex += v[r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int li = r * PRIM_THREADS + threadIdx.x;
        int64_t i = base + li;
        if (i < n) out[i] = sdata[li + (li >> 5)];
    }
}

// scratch must hold scan_scratch_elems(n) uint32.  in may alias out.
static inline size_t scan_scratch_elems(int64_t n) {
    size_t tot = 0;
    while (n > PRIM_TILE) { n = (n + PRIM_TILE - 1) / PRIM_TILE; tot += (size_t)n + 1; }
    return tot + 2;
}
static void exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* scratch, cudaStream_t st, int64_t* launches = nullptr) {
    if (n <= 0) return;
    int nt = div_up(n, PRIM_TILE);
    if (nt == 1) {
        scan_apply_kernel<<<1, PRIM_THREADS, 0, st>>>(in, n, nullptr, out);
        if (launches) *launches += 1;
        return;
    }
    uint32_t* tsum = scratch;
    scan_reduce_kernel<<<nt, PRIM_THREADS, 0, st>>>(in, n, tsum);
    exclusive_scan_u32(tsum, tsum, nt, scratch + nt + 1, st, launches);
    scan_apply_kernel<<<nt, PRIM_THREADS, 0, st>>>(in, n, tsum, out);
    if (launches) *launches += 2;
}

// ---------------------------------------------------------------------------------------------
// radix sort
template <class K>
__global__ void __launch_bounds__(PRIM_THREADS) rs_hist_kernel(const K* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ table, int ntiles) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * PRIM_TILE;
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        int64_t i = base + r * PRIM_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    table[(size_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

// Stable scatter.  Warp w owns tile items [w*512, w*512+512); item j of lane l is w*512 + j*32 + l, so
// (warp, round, lane) order == index order and ranks computed round by round are stable.  The tile is first
// sorted by digit inside shared memory and then written out in sorted order, so consecutive threads write
// consecutive addresses of a digit's run (a direct scatter would touch 32 sectors per store instruction).
template <class K>
__global__ void __launch_bounds__(PRIM_THREADS) rs_scatter_kernel(const K* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, int shift,
                                                                  const uint32_t* __restrict__ table, int ntiles, K* __restrict__ okeys,
                                                                  uint32_t* __restrict__ ovals) {
    __shared__ uint32_t whist[8][256];
    __shared__ uint32_t gofs[256];       // global position of sorted-tile slot t of digit d = gofs[d] + t
    __shared__ uint32_t dstart[256];     // first sorted-tile slot of digit d
    __shared__ uint32_t swarp[8];
    __shared__ uint64_t stage64[PRIM_TILE];
    K* stage_k = reinterpret_cast<K*>(stage64);
    uint32_t* stage_v = reinterpret_cast<uint32_t*>(stage64);
    const unsigned lane = lane_id(), w = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int i = 0; i < 8; i++) whist[i][threadIdx.x] = 0;
    __syncthreads();
    const int64_t tile_base = (int64_t)blockIdx.x * PRIM_TILE;
    const int tile_count = (int)((n - tile_base) < (int64_t)PRIM_TILE ? (n - tile_base) : (int64_t)PRIM_TILE);
    int64_t base = tile_base + w * 512;
    K key[PRIM_ITEMS];
    uint32_t val[PRIM_ITEMS];
    uint32_t rank[PRIM_ITEMS];
#pragma unroll
    for (int j = 0; j < PRIM_ITEMS; j++) {
        int64_t i = base + j * 32 + lane;
        key[j] = (i < n) ? keys[i] : (K)0;
        val[j] = (i < n) ? (vals ? vals[i] : (uint32_t)i) : 0u;
    }
#pragma unroll
    for (int j = 0; j < PRIM_ITEMS; j++) {
        int64_t i = base + j * 32 + lane;
        bool valid = i < n;
        uint32_t d = valid ? ((uint32_t)(key[j] >> shift) & 255u) : 0xffffffffu;
        unsigned mask = __match_any_sync(0xffffffffu, d);
        uint32_t pre = valid ? whist[w][d] : 0u;
        __syncwarp();
        rank[j] = pre + __popc(mask & lt);
        if (valid && lane == (unsigned)(__ffs(mask) - 1)) whist[w][d] = pre + __popc(mask);
        __syncwarp();
    }
    __syncthreads();
    {
        uint32_t run = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t c = whist[i][threadIdx.x];
            whist[i][threadIdx.x] = run;
            run += c;
        }
        uint32_t tot;
        const uint32_t ds = block_excl_scan_256(run, swarp, &tot);
        dstart[threadIdx.x] = ds;
        gofs[threadIdx.x] = table[(size_t)threadIdx.x * ntiles + blockIdx.x] - ds;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PRIM_ITEMS; j++) {
        int64_t i = base + j * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(key[j] >> shift) & 255u;
            rank[j] = dstart[d] + whist[w][d] + rank[j];       // slot in the sorted tile
            stage_k[rank[j]] = key[j];
        }
    }
    __syncthreads();
    uint32_t gpos[PRIM_ITEMS];
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        const int tpos = r * PRIM_THREADS + threadIdx.x;
        gpos[r] = 0;
        if (tpos < tile_count) {
            const K k = stage_k[tpos];
            gpos[r] = gofs[(uint32_t)(k >> shift) & 255u] + (uint32_t)tpos;
            okeys[gpos[r]] = k;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PRIM_ITEMS; j++) {
        int64_t i = base + j * 32 + lane;
        if (i < n) stage_v[rank[j]] = val[j];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PRIM_ITEMS; r++) {
        const int tpos = r * PRIM_THREADS + threadIdx.x;
        if (tpos < tile_count) ovals[gpos[r]] = stage_v[tpos];
    }
}

// Sorts (keys, vals) ascending by key; vals==nullptr on entry means payload = index (iota).
// keys_a/vals_a hold the input and are clobbered; result pointers returned through *rk / *rv
// (they alias one of the two buffer sets).  table: 256*ntiles + scan scratch.
template <class K>
struct RadixSortPlan {
    int ntiles;
    size_t table_elems, scratch_elems;
    RadixSortPlan(int64_t n) {
        ntiles = div_up(n, PRIM_TILE);
        table_elems = (size_t)256 * ntiles;
        scratch_elems = scan_scratch_elems((int64_t)table_elems);
    }
    size_t temp_u32() const { return table_elems + scratch_elems; }
};

template <class K>
static void radix_sort_pairs(K* keys_a, uint32_t* vals_a, K* keys_b, uint32_t* vals_b, int64_t n, int key_bits, bool iota_vals,
                             uint32_t* temp_u32, cudaStream_t st, K** rk, uint32_t** rv, int64_t* launches = nullptr) {
    RadixSortPlan<K> plan(n);
    uint32_t* table = temp_u32;
    uint32_t* scratch = temp_u32 + plan.table_elems;
    K* kin = keys_a; K* kout = keys_b;
    uint32_t* vin = vals_a; uint32_t* vout = vals_b;
    bool first = true;
    for (int shift = 0; shift < key_bits; shift += 8) {
        rs_hist_kernel<K><<<plan.ntiles, PRIM_THREADS, 0, st>>>(kin, n, shift, table, plan.ntiles);
        exclusive_scan_u32(table, table, (int64_t)plan.table_elems, scratch, st, launches);
        rs_scatter_kernel<K><<<plan.ntiles, PRIM_THREADS, 0, st>>>(kin, (first && iota_vals) ? nullptr : vin, n, shift, table, plan.ntiles, kout, vout);
        if (launches) *launches += 2;
        first = false;
        K* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    *rk = kin;
    *rv = vin;
}

}  // namespace nbk
