// rz_path.cu — K1: the persistent FP32 brute-force path-tracing megakernel for sm_100a.
//
// Replaces the body of Tracer.render's pixel loop and everything under it
// (reference src/renderer.zig:85-96, bounceRay :103-126, BVH.findHit hit.zig:181-216,
// Sphere.hitInner geom.zig:38-66, Material.scatter material.zig:167-176).
//
// Execution model
//   * persistent CTAs, grid = SMs x resident CTAs; each WARP pulls work units
//     (32-pixel tile x `chunk` samples) from one global atomic counter;
//   * every lane carries R independent paths ("streams").  When a path ends its stream is
//     REGENERATED at once with the next (pixel, sample) of the warp's unit (ballot + popc
//     rank), so the warp-uniform closest-hit loop always runs with full lanes: there is no
//     tail of long paths holding 31 idle lanes;
//   * radiance is accumulated with 64-bit fixed-point (2^-32) integer atomics, so the sum is
//     exact and independent of execution order => bit-identical images for any scheduling and
//     any multi-GPU row sharding;
//   * K1: the whole sphere set is staged once per CTA into shared memory with a 1-D bulk
//     async copy (cp.async.bulk + mbarrier, SASS UBLKCP) and searched by brute force with
//     warp-uniform (broadcast) LDS.128 operands and Blackwell packed FP32x2 arithmetic
//     (FFMA2/FADD2/FMUL2, two spheres per instruction): 11 issue slots per stationary
//     sphere PAIR and ray, 14 per moving pair (rz_search_brute2);
//   * large scenes use the BVH kernel of rz_bvh_trace.cu (K3) instead.
#include <cstdlib>

#include "rz_search.cuh"

// ------------------------------------------------------------------------------ the kernel
struct RzStream {
    RzRay ray;
    float3 thr;       // product of attenuations so far (bounceRay's vmul on unwind, renderer.zig:118)
    uint32_t lp;      // local (compact) pixel index -> accumulator slot
    uint32_t gpix;    // global pixel index j*W+i    -> RNG counter
    uint32_t sample;  // global sample index         -> RNG counter
    uint32_t seg;     // segments traced so far (max_bounces - depth of bounceRay)
    bool alive;
};

// G = sphere PAIRS per search-loop iteration
template <int R, int G, bool STATS, int MB>
__global__ void __launch_bounds__(128, MB) rz_path_kernel(const RzPathArgs a) {
    extern __shared__ __align__(16) unsigned char rz_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    float4 *s_pk = reinterpret_cast<float4 *>(rz_smem);

    rz_stage_scene_pk(a.set, s_pk, &s_bar);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;

    // warp-uniform unit state
    bool have_unit = true;
    uint32_t unit_lp0 = 0, unit_s0 = 0, unit_paths = 0, k_next = 0;

    RzStream st[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        st[r].alive = false;
        st[r].ray.o = f3(0.f, 0.f, 0.f);
        st[r].ray.d = f3(0.f, 1.f, 0.f);
        st[r].ray.time = 0.f;
        st[r].ray.self_k = -1;
        st[r].thr = f3(0.f, 0.f, 0.f);
        st[r].lp = 0; st[r].gpix = 0; st[r].sample = 0; st[r].seg = 0;
    }

    unsigned long long c_paths = 0, c_segs = 0, c_nodes = 0, c_sph = 0, c_hit[3] = {0, 0, 0}, c_sky = 0, c_abs = 0, c_depth = 0;

    while (true) {
        // ---------------------------------------------------------------- regenerate
#pragma unroll
        for (int r = 0; r < R; r++) {
            bool need = !st[r].alive;
            while (true) {
                const unsigned mask = __ballot_sync(0xffffffffu, need);
                if (mask == 0u || !have_unit) break;
                const uint32_t avail = unit_paths - k_next;
                if (avail == 0u) {
                    unsigned u = 0;
                    if (lane == 0) u = atomicAdd(a.unit_counter, 1u);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= a.n_units) { have_unit = false; break; }
                    const uint32_t tile = u / a.n_chunks, chunk = u - tile * a.n_chunks;
                    unit_lp0 = tile * 32u;
                    unit_s0 = chunk * a.chunk;
                    unit_paths = 32u * min(a.chunk, a.spp - unit_s0);
                    k_next = 0;
                    continue;
                }
                const uint32_t rank = __popc(mask & lt_mask);
                if (need && rank < avail) {
                    const uint32_t k = k_next + rank;
                    const uint32_t lp = unit_lp0 + (k & 31u);
                    if (lp < a.n_local_px) {
                        uint32_t pi, pj;
                        rz_local_to_global(lp, a.width, a.shard_index, a.shard_count, a.band_rows, pi, pj);
                        st[r].lp = lp;
                        st[r].gpix = pj * a.width + pi;
                        st[r].sample = a.sample_offset + unit_s0 + (k >> 5);
                        st[r].ray = rz_camera_ray(a.cam, pi, pj, st[r].gpix, st[r].sample, a.seed_lo, a.seed_hi);
                        st[r].thr = f3(1.f, 1.f, 1.f);
                        st[r].seg = 0;
                        st[r].alive = a.max_depth > 0u;
                        if (STATS) { c_paths++; if (!st[r].alive) c_depth++; }
                        need = !st[r].alive;
                    }
                }
                k_next += min((uint32_t)__popc(mask), avail);
            }
        }
        bool any_alive = false;
#pragma unroll
        for (int r = 0; r < R; r++) any_alive |= st[r].alive;
        if (!__any_sync(0xffffffffu, any_alive)) break;

        // ---------------------------------------------------------------- closest hit
        float bt[R];
        int bk[R];
        RzRay rays[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            bt[r] = 3.0e38f;
            bk[r] = -1;
            rays[r] = st[r].ray;
        }
        rz_search_brute2<R, G>(s_pk, (int)a.set.n_static_pad, (int)a.set.n_pad, rays, a.t_min, bt, bk);

        // ---------------------------------------------------------------- shade
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (!st[r].alive) continue;
            if (STATS) c_segs++;
            uint32_t kind;
            const int res = rz_shade_segment(a, st[r].ray, st[r].thr, st[r].seg, st[r].lp, st[r].gpix, st[r].sample, bk[r], kind);
            if (STATS) {
                if (kind < 3u) c_hit[kind]++;
                if (res == RZ_END_SKY) c_sky++;
                if (res == RZ_END_ABSORBED) c_abs++;
                if (res == RZ_END_DEPTH) c_depth++;
            }
            if (res != RZ_CONT) st[r].alive = false;
        }
    }

    if (STATS) {
        unsigned long long v[10] = {c_paths, c_segs, c_sph, c_nodes, c_hit[0], c_hit[1], c_hit[2], c_sky, c_abs, c_depth};
#pragma unroll
        for (int i = 0; i < 10; i++) {
            unsigned long long s = v[i];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && s) atomicAdd(&a.stats->v[i], s);
        }
    }
}

// ------------------------------------------------------------------------------ launcher
template <int R, int G, bool STATS, int MB = 5>
static cudaError_t rz_launch_one(const RzPathArgs &a, int sm_count, size_t smem, cudaStream_t stream, int *grid_out) {
    auto kern = rz_path_kernel<R, G, STATS, MB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const int grid = sm_count * per_sm;
    if (grid_out) *grid_out = grid;
    kern<<<grid, 128, smem, stream>>>(a);
    return cudaGetLastError();
}

// Forces the lazily loaded path kernels into the context (called once from rayz_cuda_create so that
// the first render does not pay module loading).
extern "C" cudaError_t rz_path_warm(void) {
    cudaFuncAttributes fa;
    return cudaFuncGetAttributes(&fa, rz_path_kernel<2, 2, false, 8>);
}

// rays_per_thread in {1,2}.
extern "C" cudaError_t rz_launch_path(const RzPathArgs *a, int rays_per_thread, int collect_stats, int sm_count, cudaStream_t stream,
                                      int *grid_out) {
    const bool stats = collect_stats != 0;
    const size_t smem = (size_t)(a->set.n_pad + (a->set.n_pad - a->set.n_static_pad)) * 16u;
    if (rays_per_thread == 1) {
        return stats ? rz_launch_one<1, 2, true>(*a, sm_count, smem, stream, grid_out)
                     : rz_launch_one<1, 2, false>(*a, sm_count, smem, stream, grid_out);
    }
    // 8 resident CTAs per SM (64 registers, 232 B of spills outside the search loop) measured 3 % faster than
    // the 5 CTAs the unconstrained 88-register build gets: 1522 vs 1472 Mpaths/s at config 2 / 100 spp.
    return stats ? rz_launch_one<2, 2, true>(*a, sm_count, smem, stream, grid_out)
                 : rz_launch_one<2, 2, false, 8>(*a, sm_count, smem, stream, grid_out);
}
