// rz_path.cu — K1: the brute-force FP32 path tracer for sm_100a, three kernels.
//
// Replaces the body of Tracer.render's pixel loop and everything under it
// (reference src/renderer.zig:85-96, bounceRay :103-126, BVH.findHit hit.zig:181-216,
// Sphere.hitInner geom.zig:38-66, Material.scatter material.zig:167-176).
//
//   rz_primary_kernel  (K1a)  camera segments; sphere set culled against each 32-pixel tile's frustum
//   rz_second_kernel   (K1c)  the next few segments, one per launch, over queue entries sorted by
//                             (origin cell, octant, reach); sphere set culled from each unit's actual rays
//   rz_path_kernel     (K1b)  the whole path loop in one persistent kernel (RZ_VARIANT_MEGA_SINGLE), or (QUEUE) every
//                             segment after the sorted stages when no host-built BVH is there for the tail
// The host (rz_context.cu) runs them in passes sized by the HBM queues between them; by default the tail of the paths —
// what survives the sorted stages — goes to the BVH kernel of rz_bvh_trace.cu instead of K1b.  All use the same packed
// arithmetic per sphere (rz_search.cuh), the same shading and RNG keys: the image does not depend on the staging.
//
// Execution model of rz_path_kernel
//   * persistent CTAs, grid = SMs x resident CTAs; each WARP pulls work units
//     (32-pixel tile x `chunk` samples, or 512 queue entries) from one global atomic counter;
//   * every lane carries R independent paths ("streams").  When a path ends its stream is
//     REGENERATED at once with the next path of the warp's unit (ballot + popc rank), so the
//     warp-uniform closest-hit loop always runs with full lanes: there is no tail of long
//     paths holding 31 idle lanes;
//   * radiance is accumulated with 64-bit fixed-point (2^-32) integer atomics, so the sum is
//     exact and independent of execution order => bit-identical images for any scheduling,
//     staging and multi-GPU row sharding;
//   * the whole sphere set is staged once per CTA into shared memory with a 1-D bulk
//     async copy (cp.async.bulk + mbarrier, SASS UBLKCP) and searched by brute force with
//     warp-uniform (broadcast) LDS.128 operands and Blackwell packed FP32x2 arithmetic
//     (FFMA2/FADD2/FMUL2, two spheres per instruction): 11 issue slots per stationary
//     sphere PAIR and ray, 14 per moving pair (rz_search_brute2);
//   * large scenes use the BVH kernels of rz_bvh_trace.cu (K3) instead.
#include <algorithm>

#include "rz_search.cuh"

// Resident CTAs per SM the staged kernels are compiled for (register cap = 65536 / (128 * N)); undefined = ptxas decides.
#ifndef RZ_SECOND_MINB
#define RZ_SECOND_MINB 7   // 72 registers: the sorted-stage kernel waits on its gathers, a seventh CTA hides more of them (45.6 -> 44.3 ms)
#endif
#define RZ_SECOND_BOUNDS __launch_bounds__(128, RZ_SECOND_MINB)
#ifdef RZ_PRIMARY_MINB
#define RZ_PRIMARY_BOUNDS __launch_bounds__(128, RZ_PRIMARY_MINB)
#else
#define RZ_PRIMARY_BOUNDS __launch_bounds__(128)
#endif

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
template <int R, int G, bool STATS, int MB, bool QUEUE>
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
    // QUEUE: paths start from the entries the primary kernel appended (512 per work unit)
    const uint32_t n_entries = QUEUE ? min(*a.q_in_count, a.queue_cap) : 0u;
    const uint32_t n_units = QUEUE ? (n_entries + 511u) / 512u : a.n_units;

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
                    if (u >= n_units) { have_unit = false; break; }
                    if (QUEUE) {
                        unit_lp0 = u * 512u;                       // first queue entry of the unit
                        unit_paths = min(512u, n_entries - unit_lp0);
                    } else {
                        const uint32_t tile = u / a.n_chunks, chunk = u - tile * a.n_chunks;
                        unit_lp0 = tile * 32u;
                        unit_s0 = chunk * a.chunk;
                        unit_paths = 32u * min(a.chunk, a.spp - unit_s0);
                    }
                    k_next = 0;
                    continue;
                }
                const uint32_t rank = __popc(mask & lt_mask);
                if (need && rank < avail) {
                    const uint32_t k = k_next + rank;
                    if (QUEUE) {
                        const float4 *e = a.q_in + (size_t)(unit_lp0 + k) * 4u;
                        const float4 qa = __ldcs(e), qb = __ldcs(e + 1), qc = __ldcs(e + 2), qd = __ldcs(e + 3);
                        st[r].ray.o = f3(qa.x, qa.y, qa.z); st[r].ray.time = qa.w;
                        st[r].ray.d = f3(qb.x, qb.y, qb.z); st[r].ray.self_k = __float_as_int(qb.w);
                        st[r].thr = f3(qc.x, qc.y, qc.z); st[r].seg = __float_as_uint(qc.w);
                        st[r].lp = __float_as_uint(qd.x); st[r].gpix = __float_as_uint(qd.y); st[r].sample = __float_as_uint(qd.z);
                        st[r].alive = true;
                        need = false;
                    } else {
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
            if (STATS) { c_segs++; c_sph += a.set.n; }   // brute force: every sphere of the set
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

// ------------------------------------------------------------------------------ primary kernel
// Stage 1 of the staged K1: the camera segment of every path (Camera.getRay camera.zig:59-77 + the first
// bounceRay level renderer.zig:103-126).  Camera rays of a 32-pixel tile are coherent, so the warp first
// culls the sphere set against the tile's frustum — a cone around the tile's mean direction, widened by the
// pixel footprint, the thin-lens blur and each sphere's motion — and the packed search then runs over the
// surviving handful of sphere pairs instead of all of them (rz_search_list2: same arithmetic, same (t, k)).
// Paths that scatter are appended, ballot-compacted, to the HBM queue the sorted stages start from; paths
// that leave the scene or are absorbed accumulate here.  36 % of all segments are camera segments.  The cull
// (rz_tile_cone / rz_tile_keep, rz_device.cuh) is host + device and property-tested on the CPU.
template <bool STATS>
__global__ void RZ_PRIMARY_BOUNDS rz_primary_kernel(const RzPathArgs a) {
    extern __shared__ __align__(16) unsigned char rz_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    float4 *s_pk = reinterpret_cast<float4 *>(rz_smem);
    const uint32_t n_sp = a.set.n_static_pad >> 1, n_mp = (a.set.n_pad - a.set.n_static_pad) >> 1;   // sphere pairs
    const uint32_t pk_f4 = a.set.n_static_pad + 2u * (a.set.n_pad - a.set.n_static_pad);
    unsigned short *ls = reinterpret_cast<unsigned short *>(s_pk + pk_f4) + (threadIdx.x >> 5) * (n_sp + n_mp);
    unsigned short *lm = ls + n_sp;

    rz_stage_scene_pk(a.set, s_pk, &s_bar);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned long long c_paths = 0, c_segs = 0, c_sph = 0, c_hit[3] = {0, 0, 0}, c_sky = 0, c_abs = 0, c_depth = 0;

    while (true) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(a.unit_counter, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= a.n_units) break;
        u += a.unit_base;
        const uint32_t tile = u / a.n_chunks, chunk = u - tile * a.n_chunks;
        const uint32_t lp = tile * 32u + lane;
        const bool valid = lp < a.n_local_px;
        uint32_t pi = 0, pj = 0;
        if (valid) rz_local_to_global(lp, a.width, a.shard_index, a.shard_count, a.band_rows, pi, pj);
        const uint32_t gpix = pj * a.width + pi;

        // ---- cone around the tile's camera rays: axis = mean pixel direction, half-angle from the pixel corners
        const float3 pc = rz_tile_pixel_dir(a.cam, pi, pj);
        float3 ax = valid ? normalize3(pc) : f3(0.f, 0.f, 0.f);
        for (int o = 16; o > 0; o >>= 1) {
            ax.x += __shfl_xor_sync(0xffffffffu, ax.x, o); ax.y += __shfl_xor_sync(0xffffffffu, ax.y, o); ax.z += __shfl_xor_sync(0xffffffffu, ax.z, o);
        }
        const bool has_axis = rz_tile_axis(ax);
        float cmin = valid ? rz_tile_corner_cos(a.cam, pc, ax) : 1.0f;
        for (int o = 16; o > 0; o >>= 1) cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        const RzTileCone cone = rz_tile_cone(a.cam, ax, has_axis, cmin, a.focus_dist, a.lens_radius);
        auto keep = [&](float cx, float cy, float cz, float vx, float vy, float vz, float w) -> bool {
            return rz_tile_keep(cone, cx, cy, cz, vx, vy, vz, w);
        };
        int n_ls = 0, n_lm = 0;
        for (uint32_t p0 = 0; p0 < n_sp; p0 += 32u) {
            const uint32_t p = p0 + lane;
            bool k = false;
            if (p < n_sp) {
                const float4 A = s_pk[2 * p], B = s_pk[2 * p + 1];
                k = keep(A.x, A.z, B.x, 0.f, 0.f, 0.f, B.z) || keep(A.y, A.w, B.y, 0.f, 0.f, 0.f, B.w);
            }
            const unsigned m = __ballot_sync(0xffffffffu, k);
            if (k) ls[n_ls + __popc(m & lt_mask)] = (unsigned short)p;
            n_ls += __popc(m);
        }
        const float4 *mv = s_pk + a.set.n_static_pad;
        for (uint32_t p0 = 0; p0 < n_mp; p0 += 32u) {
            const uint32_t p = p0 + lane;
            bool k = false;
            if (p < n_mp) {
                const float4 A = mv[4 * p], B = mv[4 * p + 1], VA = mv[4 * p + 2], VB = mv[4 * p + 3];
                k = keep(A.x, A.z, B.x, VA.x, VA.z, VB.x, B.z) || keep(A.y, A.w, B.y, VA.y, VA.w, VB.y, B.w);
            }
            const unsigned m = __ballot_sync(0xffffffffu, k);
            if (k) lm[n_lm + __popc(m & lt_mask)] = (unsigned short)p;
            n_lm += __popc(m);
        }
        __syncwarp();

        // ---- the tile's samples of this chunk, two per lane and iteration
        const uint32_t s0 = chunk * a.chunk, ns = min(a.chunk, a.spp - s0);
        for (uint32_t s = 0; s < ns; s += 2u) {
            RzRay rays[2];
            bool live[2];
            uint32_t smp[2];
            float bt[2];
            int bk[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                smp[r] = a.sample_offset + s0 + s + (uint32_t)r;
                const bool have = valid && (s + (uint32_t)r < ns);
                live[r] = have && a.max_depth > 0u;
                if (STATS && have) { c_paths++; if (!live[r]) c_depth++; }
                if (live[r]) rays[r] = rz_camera_ray(a.cam, pi, pj, gpix, smp[r], a.seed_lo, a.seed_hi);
                else { rays[r].o = f3(0.f, 0.f, 0.f); rays[r].d = f3(0.f, 1.f, 0.f); rays[r].time = 0.f; rays[r].self_k = -1; }
                bt[r] = 3.0e38f; bk[r] = -1;
            }
            rz_search_list2<2>(s_pk, ls, n_ls, lm, n_lm, (int)a.set.n_static_pad, rays, a.t_min, bt, bk);
            if (STATS) c_sph += (unsigned long long)(2 * (n_ls + n_lm)) * (live[0] ? 1u : 0u) + (unsigned long long)(2 * (n_ls + n_lm)) * (live[1] ? 1u : 0u);
            bool cont[2] = {false, false};
            float3 thr[2] = {f3(1.f, 1.f, 1.f), f3(1.f, 1.f, 1.f)};
            uint32_t seg[2] = {0u, 0u};
#pragma unroll
            for (int r = 0; r < 2; r++) {
                uint32_t kind = 3u;
                if (live[r]) {
                    if (STATS) c_segs++;
                    const int res = rz_shade_segment(a, rays[r], thr[r], seg[r], lp, gpix, smp[r], bk[r], kind);
                    if (STATS) {
                        if (kind < 3u) c_hit[kind]++;
                        if (res == RZ_END_SKY) c_sky++;
                        if (res == RZ_END_ABSORBED) c_abs++;
                        if (res == RZ_END_DEPTH) c_depth++;
                    }
                    cont[r] = res == RZ_CONT;
                }
            }
            const uint32_t lp2[2] = {lp, lp}, gpix2[2] = {gpix, gpix};
            rz_queue_push2(a, cont, lane, lt_mask, rays, thr, seg, lp2, gpix2, smp);
        }
        __syncwarp();   // the lists are rewritten for the next unit
    }

    if (STATS) {
        unsigned long long v[10] = {c_paths, c_segs, c_sph, 0ull, c_hit[0], c_hit[1], c_hit[2], c_sky, c_abs, c_depth};
#pragma unroll
        for (int i = 0; i < 10; i++) {
            unsigned long long sum = v[i];
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0 && sum) atomicAdd(&a.stats->v[i], sum);
        }
    }
}

// ------------------------------------------------------------------------------ sorted-segment kernel
// The stages after the camera segment, one launch per segment (segments 2..6 by default).  Scattered rays are incoherent,
// but their queue entries have been SORTED by (origin cell, direction octant, reach class) (rz_sort_key + cub radix sort),
// so 512 consecutive entries start close together, head the same way and leave the sphere layer after a similar distance.
// Per unit the warp merges the bounds its sorted KEYS stand for (rz_key_bounds: the cells of the origins, the axes on which
// every ray moves the same way, the longest stay T inside the box around the non-huge spheres); a sphere can then only be
// hit if it is not behind the cell box on such an axis and lies within T + r of it (rz_unit_keep).  The packed search runs
// over that list (12 % of the spheres on the RTOW scene at 500 spp) with the same arithmetic per sphere, so (t, k) is
// unchanged.  Both functions are host + device and property-tested on the CPU (tests/test_hostsim_cpu.py).
template <bool STATS>
__global__ void RZ_SECOND_BOUNDS rz_second_kernel(const RzPathArgs a) {
    extern __shared__ __align__(16) unsigned char rz_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    float4 *s_pk = reinterpret_cast<float4 *>(rz_smem);
    const uint32_t n_sp = a.set.n_static_pad >> 1, n_mp = (a.set.n_pad - a.set.n_static_pad) >> 1;
    const uint32_t pk_f4 = a.set.n_static_pad + 2u * (a.set.n_pad - a.set.n_static_pad);
    unsigned short *ls = reinterpret_cast<unsigned short *>(s_pk + pk_f4) + (threadIdx.x >> 5) * (n_sp + n_mp);
    unsigned short *lm = ls + n_sp;

    rz_stage_scene_pk(a.set, s_pk, &s_bar);

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t n_in = min(*a.q_in_count, a.queue_cap);
    const uint32_t ue = a.unit_entries;
    const uint32_t n_units = (n_in + ue - 1u) / ue;
    unsigned long long c_segs = 0, c_sph = 0, c_hit[3] = {0, 0, 0}, c_sky = 0, c_abs = 0, c_depth = 0;

    while (true) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(a.unit_counter, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const uint32_t e0 = u * ue, ne = min(ue, n_in - e0);

        // ---- what the unit's rays have in common
        RzUnitBounds U;
        rz_unit_bounds_init(U);
        // Bounds from the sorted KEYS, not from the entries: decoding 512 16-bit keys (1 KB, coalesced) replaces a gather of
        // 512 x 32 B through the index that was 18 % of this kernel's warp-state samples.  The key gives conservative bounds:
        // the origin lies in its cell (open-ended for the outermost cells, where out-of-box origins are clamped), the octant
        // is exact, and the reach is below the upper edge of its class (the top class is unbounded).
        uint32_t prev_key = 0xffffffffu;   // the keys are sorted: a lane mostly meets the key it has just decoded
        for (uint32_t i = lane; i < ne; i += 32u) {
            const uint32_t key = a.q_in_keys[e0 + i];
            if (key == prev_key) continue;
            prev_key = key;
            rz_unit_bounds_add_key(U, a, key);
        }
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int ax = 0; ax < 3; ax++) {
                U.lo[ax] = fminf(U.lo[ax], __shfl_xor_sync(0xffffffffu, U.lo[ax], o));
                U.hi[ax] = fmaxf(U.hi[ax], __shfl_xor_sync(0xffffffffu, U.hi[ax], o));
            }
            U.T = fmaxf(U.T, __shfl_xor_sync(0xffffffffu, U.T, o));
            U.all_pos &= __shfl_xor_sync(0xffffffffu, U.all_pos, o);
            U.all_neg &= __shfl_xor_sync(0xffffffffu, U.all_neg, o);
        }
        rz_unit_bounds_finish(U);
        auto keep = [&](float cx, float cy, float cz, float vx, float vy, float vz, float w) -> bool {
            return rz_unit_keep(U, a.huge_radius, cx, cy, cz, vx, vy, vz, w);
        };
        int n_ls = 0, n_lm = 0;
        for (uint32_t p0 = 0; p0 < n_sp; p0 += 32u) {
            const uint32_t p = p0 + lane;
            bool k = false;
            if (p < n_sp) {
                const float4 A = s_pk[2 * p], B = s_pk[2 * p + 1];
                k = keep(A.x, A.z, B.x, 0.f, 0.f, 0.f, B.z) || keep(A.y, A.w, B.y, 0.f, 0.f, 0.f, B.w);
            }
            const unsigned m = __ballot_sync(0xffffffffu, k);
            if (k) ls[n_ls + __popc(m & lt_mask)] = (unsigned short)p;
            n_ls += __popc(m);
        }
        const float4 *mv = s_pk + a.set.n_static_pad;
        for (uint32_t p0 = 0; p0 < n_mp; p0 += 32u) {
            const uint32_t p = p0 + lane;
            bool k = false;
            if (p < n_mp) {
                const float4 A = mv[4 * p], B = mv[4 * p + 1], VA = mv[4 * p + 2], VB = mv[4 * p + 3];
                k = keep(A.x, A.z, B.x, VA.x, VA.z, VB.x, B.z) || keep(A.y, A.w, B.y, VA.y, VA.w, VB.y, B.w);
            }
            const unsigned m = __ballot_sync(0xffffffffu, k);
            if (k) lm[n_lm + __popc(m & lt_mask)] = (unsigned short)p;
            n_lm += __popc(m);
        }
        __syncwarp();

        // ---- the unit's rays, two per lane and iteration
        for (uint32_t b0 = 0; b0 < ne; b0 += 64u) {
            RzRay rays[2];
            float3 thr[2];
            uint32_t seg[2], lp[2], gpix[2], smp[2];
            bool live[2];
            float bt[2];
            int bk[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const uint32_t i = b0 + lane + 32u * (uint32_t)r;
                live[r] = i < ne;
                if (live[r]) {
                    const float4 *e = a.q_in + (size_t)a.q_in_idx[e0 + i] * 4u;
                    const float4 qa = __ldcs(e), qb = __ldcs(e + 1), qc = __ldcs(e + 2), qd = __ldcs(e + 3);
                    rays[r].o = f3(qa.x, qa.y, qa.z); rays[r].time = qa.w;
                    rays[r].d = f3(qb.x, qb.y, qb.z); rays[r].self_k = __float_as_int(qb.w);
                    thr[r] = f3(qc.x, qc.y, qc.z); seg[r] = __float_as_uint(qc.w);
                    lp[r] = __float_as_uint(qd.x); gpix[r] = __float_as_uint(qd.y); smp[r] = __float_as_uint(qd.z);
                } else {
                    rays[r].o = f3(0.f, 0.f, 0.f); rays[r].d = f3(0.f, 1.f, 0.f); rays[r].time = 0.f; rays[r].self_k = -1;
                    thr[r] = f3(0.f, 0.f, 0.f); seg[r] = 0; lp[r] = 0; gpix[r] = 0; smp[r] = 0;
                }
                bt[r] = 3.0e38f; bk[r] = -1;
            }
            rz_search_list2<2>(s_pk, ls, n_ls, lm, n_lm, (int)a.set.n_static_pad, rays, a.t_min, bt, bk);
            if (STATS) c_sph += (unsigned long long)(2 * (n_ls + n_lm)) * ((live[0] ? 1u : 0u) + (live[1] ? 1u : 0u));
            bool cont[2] = {false, false};
#pragma unroll
            for (int r = 0; r < 2; r++) {
                if (live[r]) {
                    if (STATS) c_segs++;
                    uint32_t kind;
                    const int res = rz_shade_segment(a, rays[r], thr[r], seg[r], lp[r], gpix[r], smp[r], bk[r], kind);
                    if (STATS) {
                        if (kind < 3u) c_hit[kind]++;
                        if (res == RZ_END_SKY) c_sky++;
                        if (res == RZ_END_ABSORBED) c_abs++;
                        if (res == RZ_END_DEPTH) c_depth++;
                    }
                    cont[r] = res == RZ_CONT;
                }
            }
            rz_queue_push2(a, cont, lane, lt_mask, rays, thr, seg, lp, gpix, smp);
        }
        __syncwarp();
    }

    if (STATS) {
        unsigned long long v[10] = {0ull, c_segs, c_sph, 0ull, c_hit[0], c_hit[1], c_hit[2], c_sky, c_abs, c_depth};
#pragma unroll
        for (int i = 0; i < 10; i++) {
            unsigned long long sum = v[i];
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0 && sum) atomicAdd(&a.stats->v[i], sum);
        }
    }
}

// ------------------------------------------------------------------------------ launchers
template <int R, int G, bool STATS, int MB, bool QUEUE>
static cudaError_t rz_launch_one(const RzPathArgs &a, int sm_count, size_t smem, cudaStream_t stream, int *grid_out) {
    auto kern = rz_path_kernel<R, G, STATS, MB, QUEUE>;
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

static size_t rz_pk_bytes(const RzPathArgs &a) { return (size_t)(a.set.n_pad + (a.set.n_pad - a.set.n_static_pad)) * 16u; }

// Forces the lazily loaded path kernels into the context (called once from rayz_cuda_create so that
// the first render does not pay module loading).
extern "C" cudaError_t rz_path_warm(void) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, rz_path_kernel<2, 2, false, 8, true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, rz_primary_kernel<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, rz_second_kernel<false>);
    return e;
}

// Shared memory the primary kernel needs: the pair-interleaved set + one pair list per warp.
extern "C" size_t rz_primary_smem_bytes(const RzPathArgs *a) {
    const size_t pairs = a->set.n_pad / 2u;
    return rz_pk_bytes(*a) + ((4u * pairs * sizeof(unsigned short) + 15u) & ~size_t(15));
}

// Stage 1: camera segments of the work units [unit_base, unit_base + n_units) -> queue.
extern "C" cudaError_t rz_launch_primary(const RzPathArgs *a, int collect_stats, int sm_count, cudaStream_t stream) {
    const size_t smem = rz_primary_smem_bytes(a);
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        const unsigned want = (a->n_units + 3u) / 4u;   // 4 warps per CTA, one unit per warp at a time
        const unsigned grid = (unsigned)std::max(1u, std::min((unsigned)(sm_count * per_sm), want));
        kern<<<grid, 128, smem, stream>>>(*a);
        return cudaGetLastError();
    };
    return collect_stats ? launch(rz_primary_kernel<true>) : launch(rz_primary_kernel<false>);
}

// Stage 2: second segments of the sorted queue (q_in through q_in_idx) -> q_out.
extern "C" cudaError_t rz_launch_second(const RzPathArgs *a, int collect_stats, int sm_count, cudaStream_t stream) {
    const size_t smem = rz_primary_smem_bytes(a);
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorInvalidConfiguration;
        kern<<<sm_count * per_sm, 128, smem, stream>>>(*a);
        return cudaGetLastError();
    };
    return collect_stats ? launch(rz_second_kernel<true>) : launch(rz_second_kernel<false>);
}

// Last stage (q_in != nullptr): the persistent megakernel started from a queue; or the single-stage form
// (q_in == nullptr) that generates its own camera rays.  rays_per_thread in {1,2}.
extern "C" cudaError_t rz_launch_path(const RzPathArgs *a, int rays_per_thread, int collect_stats, int sm_count, cudaStream_t stream,
                                      int *grid_out) {
    const bool stats = collect_stats != 0;
    const size_t smem = rz_pk_bytes(*a);
    if (a->q_in) {
        if (rays_per_thread == 1)
            return stats ? rz_launch_one<1, 2, true, 5, true>(*a, sm_count, smem, stream, grid_out)
                         : rz_launch_one<1, 2, false, 5, true>(*a, sm_count, smem, stream, grid_out);
        return stats ? rz_launch_one<2, 2, true, 5, true>(*a, sm_count, smem, stream, grid_out)
                     : rz_launch_one<2, 2, false, 8, true>(*a, sm_count, smem, stream, grid_out);
    }
    if (rays_per_thread == 1)
        return stats ? rz_launch_one<1, 2, true, 5, false>(*a, sm_count, smem, stream, grid_out)
                     : rz_launch_one<1, 2, false, 5, false>(*a, sm_count, smem, stream, grid_out);
    // 8 resident CTAs per SM (64 registers, 232 B of spills outside the search loop) measured 3 % faster than
    // the 5 CTAs the unconstrained 88-register build gets: 1522 vs 1472 Mpaths/s at config 2 / 100 spp.
    return stats ? rz_launch_one<2, 2, true, 5, false>(*a, sm_count, smem, stream, grid_out)
                 : rz_launch_one<2, 2, false, 8, false>(*a, sm_count, smem, stream, grid_out);
}
