// rz_sort.cu — puts the staged K1's queue entries in the order of their 16-bit sort keys: a one-pass counting sort
// ("binning") written for this job.  Round 1 called cub::DeviceRadixSort here (two 8-bit onesweep passes over key + index
// pairs, ~26 B of HBM traffic per entry, wrapped in a CUDA graph with a SWITCH node because cub wants its item count on the
// host); that was 12 % of the render step.
//
// What the consumer (rz_second_kernel) needs is weaker than a sort: entries with EQUAL KEYS must be contiguous and the keys
// ascending — the order inside a key's range is irrelevant (every entry of it yields the same cull bounds, and radiance is
// accumulated in integers, so the image does not depend on which 512 entries share a work unit).  Without the stability a
// multi-pass radix sort needs, one pass over the 65,536 possible keys is enough:
//   rz_bin_kernel<false>  count   keys -> bins[key]                              (reads 2 B per entry)
//   rz_bin_scan_kernel    scan    bins -> exclusive prefix = first slot of each key's range     (256 KB)
//   rz_bin_kernel<true>   scatter entry i -> slot atomicAdd(bins[key]) : idx_out[slot] = i, keys_out[slot] = key
//                                                                                (reads 2 B, writes 6 B per entry)
// i.e. ~10 B of traffic per entry and no temporary buffers.  A global atomic per entry would serialise on the popular keys
// (measured in round 1: +7 ms), so each CTA first aggregates a tile of 2048 keys in a shared-memory hash table (4096 slots,
// linear probing, warp-level __match_any_sync pre-aggregation so that a hot key costs one shared atomic per warp row): one
// global atomic per DISTINCT key of the tile.  The count pass keeps its table across tiles and flushes it only when it
// fills up.  The scattered 2- and 4-byte stores land on ~65 k slowly advancing frontiers (4 MB) that stay in the 126 MB L2
// until their sectors are full.  Every kernel takes the live entry count from device memory: no host round trip, no
// conditional graph, no 0xffff padding keys.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int RZ_BINS = 65536;
constexpr int RZ_BIN_THREADS = 256;
constexpr int RZ_BIN_ITEMS = 8;                              // keys per thread and tile
constexpr int RZ_BIN_TILE = RZ_BIN_THREADS * RZ_BIN_ITEMS;   // 2048
constexpr int RZ_BIN_SLOTS = 4096;                           // hash slots: load <= 50 % (scatter), <= 75 % (count)

struct RzBinArgs {
    const unsigned short *keys_in;   // [n] in producer order
    const unsigned int *count;       // live entries (device counter of the producing kernel)
    uint32_t cap;                    // slots of the buffers (the count is clamped to it)
    unsigned int *bins;              // [65536] counts -> (after the scan) next free slot of each key
    unsigned short *keys_out;        // [n] keys in slot order
    uint32_t *idx_out;               // [n] entry index in slot order
};

__device__ __forceinline__ uint32_t rz_bin_hash(uint32_t key) { return ((key * 40503u) >> 4) & (uint32_t)(RZ_BIN_SLOTS - 1); }

template <bool SCATTER>
__global__ void __launch_bounds__(RZ_BIN_THREADS) rz_bin_kernel(const RzBinArgs a) {
    __shared__ uint32_t t_key[RZ_BIN_SLOTS];   // key + 1; 0 = empty
    __shared__ uint32_t t_cnt[RZ_BIN_SLOTS];   // entries of the key in this tile; after the flush: their first global slot
    __shared__ uint32_t s_occ;                 // occupied slots (count pass: when to flush)
    const uint32_t tid = threadIdx.x, lane = tid & 31u, lt_mask = (1u << lane) - 1u;
    const uint32_t n = min(*a.count, a.cap);
    const uint32_t n_tiles = (n + RZ_BIN_TILE - 1) / RZ_BIN_TILE;
    for (uint32_t s = tid; s < RZ_BIN_SLOTS; s += RZ_BIN_THREADS) { t_key[s] = 0u; t_cnt[s] = 0u; }
    if (tid == 0) s_occ = 0u;
    __syncthreads();

    // one global atomic per occupied slot; the slot then holds the first global slot of its entries (scatter) and is
    // cleared for the next tile by `clear`
    auto flush = [&](bool clear) {
        for (uint32_t s = tid; s < RZ_BIN_SLOTS; s += RZ_BIN_THREADS) {
            const uint32_t k1 = t_key[s];
            if (k1) {
                const uint32_t base = atomicAdd(a.bins + (k1 - 1u), t_cnt[s]);
                if (clear) { t_key[s] = 0u; t_cnt[s] = 0u; } else t_cnt[s] = base;
            }
        }
    };

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t i0 = tile * RZ_BIN_TILE + tid;
        uint32_t key[RZ_BIN_ITEMS], slot[RZ_BIN_ITEMS], rank[RZ_BIN_ITEMS];
#pragma unroll
        for (int j = 0; j < RZ_BIN_ITEMS; j++) {
            const uint32_t i = i0 + (uint32_t)j * RZ_BIN_THREADS;
            key[j] = i < n ? (uint32_t)a.keys_in[i] : 0x10000u;   // past the end: a value no key has
        }
#pragma unroll
        for (int j = 0; j < RZ_BIN_ITEMS; j++) {
            const unsigned peers = __match_any_sync(0xffffffffu, key[j]);
            const int leader = __ffs((int)peers) - 1;
            uint32_t s = 0u, r0 = 0u;
            if ((int)lane == leader && key[j] < 0x10000u) {
                uint32_t h = rz_bin_hash(key[j]);
                for (int probe = 0; probe < RZ_BIN_SLOTS; probe++) {   // the table is never full: ends at the key's slot or an empty one
                    const uint32_t prev = atomicCAS(&t_key[h], 0u, key[j] + 1u);
                    if (prev == 0u) { if (!SCATTER) atomicAdd(&s_occ, 1u); break; }
                    if (prev == key[j] + 1u) break;
                    h = (h + 1u) & (uint32_t)(RZ_BIN_SLOTS - 1);
                }
                r0 = atomicAdd(&t_cnt[h], (uint32_t)__popc(peers));
                s = h;
            }
            slot[j] = __shfl_sync(0xffffffffu, s, leader);
            rank[j] = __shfl_sync(0xffffffffu, r0, leader) + (uint32_t)__popc(peers & lt_mask);
        }
        __syncthreads();
        if (SCATTER) {
            flush(false);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < RZ_BIN_ITEMS; j++) {
                if (key[j] < 0x10000u) {
                    const uint32_t pos = t_cnt[slot[j]] + rank[j];
                    if (pos < a.cap) {   // always true: the ranges partition [0, n)
                        a.idx_out[pos] = i0 + (uint32_t)j * RZ_BIN_THREADS;
                        a.keys_out[pos] = (unsigned short)key[j];
                    }
                }
            }
            __syncthreads();
            for (uint32_t s = tid; s < RZ_BIN_SLOTS; s += RZ_BIN_THREADS) { t_key[s] = 0u; t_cnt[s] = 0u; }
            __syncthreads();
        } else if (s_occ > (uint32_t)(RZ_BIN_SLOTS / 4)) {   // the next tile may add 2048 keys: keep the load below 75 %
            flush(true);
            __syncthreads();
            if (tid == 0) s_occ = 0u;
            __syncthreads();
        }
    }
    if (!SCATTER) flush(true);
}

// exclusive prefix over the 65,536 bins, in place: one CTA, 1024 threads x 64 consecutive bins
__global__ void __launch_bounds__(1024) rz_bin_scan_kernel(unsigned int *bins) {
    __shared__ uint32_t s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    uint4 *p = reinterpret_cast<uint4 *>(bins) + (size_t)tid * 16u;
    uint4 v[16];
    uint32_t sum = 0u;
#pragma unroll
    for (int i = 0; i < 16; i++) { v[i] = p[i]; sum += v[i].x + v[i].y + v[i].z + v[i].w; }
    uint32_t inc = sum;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((int)lane >= o) inc += t; }
    if (lane == 31u) s_warp[w] = inc;
    __syncthreads();
    if (w == 0u) {
        uint32_t x = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, x, o); if ((int)lane >= o) x += t; }
        s_warp[lane] = x;
    }
    __syncthreads();
    uint32_t run = inc - sum + (w ? s_warp[w - 1u] : 0u);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        uint4 o;
        o.x = run; run += v[i].x;
        o.y = run; run += v[i].y;
        o.z = run; run += v[i].z;
        o.w = run; run += v[i].w;
        p[i] = o;
    }
}

}  // namespace

extern "C" size_t rz_bin_scratch_bytes(void) { return (size_t)RZ_BINS * sizeof(unsigned int); }

// keys_in[0, *count) -> idx_out / keys_out: entry indices and keys grouped by ascending key.  bins: rz_bin_scratch_bytes().
extern "C" cudaError_t rz_bin_sort(const unsigned short *keys_in, const unsigned int *count, uint32_t cap, unsigned int *bins,
                                   unsigned short *keys_out, uint32_t *idx_out, int sm_count, cudaStream_t stream) {
    static int per_sm = 0;
    if (per_sm == 0) {
        int a = 0, b = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, rz_bin_kernel<false>, RZ_BIN_THREADS, 0);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, rz_bin_kernel<true>, RZ_BIN_THREADS, 0);
        if (e != cudaSuccess) return e;
        per_sm = a < b ? a : b;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
    }
    RzBinArgs a;
    a.keys_in = keys_in; a.count = count; a.cap = cap; a.bins = bins; a.keys_out = keys_out; a.idx_out = idx_out;
    const unsigned tiles = (cap + RZ_BIN_TILE - 1) / RZ_BIN_TILE;
    const unsigned grid = tiles < (unsigned)(sm_count * per_sm) ? (tiles ? tiles : 1u) : (unsigned)(sm_count * per_sm);
    cudaError_t e = cudaMemsetAsync(bins, 0, (size_t)RZ_BINS * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    rz_bin_kernel<false><<<grid, RZ_BIN_THREADS, 0, stream>>>(a);
    rz_bin_scan_kernel<<<1, 1024, 0, stream>>>(bins);
    rz_bin_kernel<true><<<grid, RZ_BIN_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

extern "C" cudaError_t rz_sort_warm(void) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, rz_bin_kernel<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, rz_bin_kernel<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, rz_bin_scan_kernel);
    return e;
}
