// rz_path.cu — K1/K3: the persistent FP32 path-tracing megakernel for sm_100a.
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
//     warp-uniform (broadcast) LDS.128 operands: 10 FP32 instructions per stationary
//     sphere-ray pair, 13 per moving one, FFMA dominated;
//   * K3: large scenes traverse a BVH2 through the read-only path (ld.global.nc).
#include "rz_device.cuh"

#define RZ_FAR_BIT 0x40000000

// ------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t rz_smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void rz_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rz_smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void rz_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rz_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rz_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rz_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(rz_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void rz_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(rz_smem_addr(bar)), "r"(parity)
            : "memory");
    }
}

// ------------------------------------------------------------------------------ candidates
// Rare path of the search: a sphere whose discriminant is positive.  Root rule of
// Sphere.hitInner (geom.zig:52-58): near root if inside (t_min, best), else far root.  For the
// sphere the ray starts on, the t~0 root is excluded analytically (see RzRay::self_k).
__device__ __forceinline__ void rz_consider(int k, float b, float disc, int self_k, float t_min, float &bt, int &bk) {
    const float sq = sqrtf(disc);
    float t = b - sq;
    int tag = k;
    if (k == self_k) {
        t = (b > 0.0f) ? b + sq : -1.0f;
        tag = k | RZ_FAR_BIT;
    } else if (t < t_min) {
        t = b + sq;
        tag = k | RZ_FAR_BIT;
    }
    if (t > t_min && t < bt) {
        bt = t;
        bk = tag;
    }
}

// ------------------------------------------------------------------------------ K1 search
// Brute force over the shared-memory sphere set for R rays at once.  Unit-length directions:
//   oc = C - o; b = d.oc; c = oc.oc - r^2; disc = b^2 - c      (geom.zig:40-48 with a = 1)
// cr.w holds -r^2 so c is three FFMAs.  G spheres per iteration share one max/branch.
template <int R, int G>
__device__ __forceinline__ void rz_search_brute(const float4 *__restrict__ s_cr, const float4 *__restrict__ s_vel,
                                                int n_static_pad, int n_pad, const RzRay (&ray)[R], float t_min,
                                                float (&bt)[R], int (&bk)[R]) {
    int i = 0;
#pragma unroll 1
    for (; i < n_static_pad; i += G) {
        float4 s[G];
#pragma unroll
        for (int j = 0; j < G; j++) s[j] = s_cr[i + j];
        float b[R][G], disc[R][G];
        float m = -1.0f;
#pragma unroll
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int j = 0; j < G; j++) {
                const float ocx = s[j].x - ray[r].o.x, ocy = s[j].y - ray[r].o.y, ocz = s[j].z - ray[r].o.z;
                b[r][j] = fmaf(ocz, ray[r].d.z, fmaf(ocy, ray[r].d.y, ocx * ray[r].d.x));
                const float c = fmaf(ocz, ocz, fmaf(ocy, ocy, fmaf(ocx, ocx, s[j].w)));
                disc[r][j] = fmaf(b[r][j], b[r][j], -c);
                m = fmaxf(m, disc[r][j]);
            }
        }
        if (m > 0.0f) {
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < G; j++)
                    if (disc[r][j] > 0.0f) rz_consider(i + j, b[r][j], disc[r][j], ray[r].self_k, t_min, bt[r], bk[r]);
        }
    }
#pragma unroll 1
    for (; i < n_pad; i += G) {
        float4 s[G], v[G];
#pragma unroll
        for (int j = 0; j < G; j++) {
            s[j] = s_cr[i + j];
            v[j] = s_vel[i - n_static_pad + j];
        }
        float b[R][G], disc[R][G];
        float m = -1.0f;
#pragma unroll
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int j = 0; j < G; j++) {
                // centre(t) = center.origin + center.dir * ray.time (geom.zig:40)
                const float ocx = fmaf(v[j].x, ray[r].time, s[j].x) - ray[r].o.x;
                const float ocy = fmaf(v[j].y, ray[r].time, s[j].y) - ray[r].o.y;
                const float ocz = fmaf(v[j].z, ray[r].time, s[j].z) - ray[r].o.z;
                b[r][j] = fmaf(ocz, ray[r].d.z, fmaf(ocy, ray[r].d.y, ocx * ray[r].d.x));
                const float c = fmaf(ocz, ocz, fmaf(ocy, ocy, fmaf(ocx, ocx, s[j].w)));
                disc[r][j] = fmaf(b[r][j], b[r][j], -c);
                m = fmaxf(m, disc[r][j]);
            }
        }
        if (m > 0.0f) {
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int j = 0; j < G; j++)
                    if (disc[r][j] > 0.0f) rz_consider(i + j, b[r][j], disc[r][j], ray[r].self_k, t_min, bt[r], bk[r]);
        }
    }
}

// ------------------------------------------------------------------------------ K3 search
// BVH2 traversal, ordered (near child first), per-thread stack.  Restates the role of
// BVH.findHit + AABB.hit (hit.zig:70-98,181-216) on a flattened, SAH-built tree; closest-hit
// results do not depend on tree shape.
__device__ __forceinline__ void rz_search_bvh(const RzPathArgs &a, const RzRay &ray, float t_min, float &bt, int &bk,
                                              unsigned long long &n_nodes, unsigned long long &n_sph) {
    const float ix = 1.0f / ray.d.x, iy = 1.0f / ray.d.y, iz = 1.0f / ray.d.z;
    const float ox = ray.o.x, oy = ray.o.y, oz = ray.o.z;
    int stack[48];
    int sp = 0;
    int node = 0;
    const float4 *nodes = reinterpret_cast<const float4 *>(a.bvh);
    while (true) {
        const float4 q0 = __ldg(nodes + node * 4 + 0);  // lox0 lox1 hix0 hix1
        const float4 q1 = __ldg(nodes + node * 4 + 1);  // loy0 loy1 hiy0 hiy1
        const float4 q2 = __ldg(nodes + node * 4 + 2);  // loz0 loz1 hiz0 hiz1
        const int4 q3 = __ldg(reinterpret_cast<const int4 *>(nodes + node * 4 + 3));
        n_nodes += 2;
        float tn[2], tf[2];
        {
            const float lox[2] = {q0.x, q0.y}, hix[2] = {q0.z, q0.w};
            const float loy[2] = {q1.x, q1.y}, hiy[2] = {q1.z, q1.w};
            const float loz[2] = {q2.x, q2.y}, hiz[2] = {q2.z, q2.w};
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const float ax = (lox[c] - ox) * ix, bx = (hix[c] - ox) * ix;
                const float ay = (loy[c] - oy) * iy, by = (hiy[c] - oy) * iy;
                const float az = (loz[c] - oz) * iz, bz = (hiz[c] - oz) * iz;
                tn[c] = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t_min));
                tf[c] = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), bt));
            }
        }
        const int child[2] = {q3.x, q3.y};
        const int cnt[2] = {q3.z, q3.w};
        int next[2];
        float nt[2];
        int nn = 0;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            // slack of 2 ulp-ish on the box test; boxes are already padded outward at build time
            if (tn[c] <= tf[c] * 1.0000004f) {
                if (child[c] < 0) {
                    const int first = ~child[c];
                    for (int e = 0; e < cnt[c]; e++) {
                        const int k = first + e;
                        const float4 s = __ldg(a.set.cr + k);
                        const float4 v = __ldg(a.set.vel + k);
                        n_sph++;
                        const float ocx = fmaf(v.x, ray.time, s.x) - ox;
                        const float ocy = fmaf(v.y, ray.time, s.y) - oy;
                        const float ocz = fmaf(v.z, ray.time, s.z) - oz;
                        const float b = fmaf(ocz, ray.d.z, fmaf(ocy, ray.d.y, ocx * ray.d.x));
                        const float cc = fmaf(ocz, ocz, fmaf(ocy, ocy, fmaf(ocx, ocx, s.w)));
                        const float disc = fmaf(b, b, -cc);
                        if (disc > 0.0f) rz_consider(k, b, disc, ray.self_k, t_min, bt, bk);
                    }
                } else {
                    next[nn] = child[c];
                    nt[nn] = tn[c];
                    nn++;
                }
            }
        }
        if (nn == 2) {
            const bool swap = nt[1] < nt[0];
            const int nearc = swap ? next[1] : next[0];
            const int farc = swap ? next[0] : next[1];
            if (sp < 48) stack[sp++] = farc;
            node = nearc;
        } else if (nn == 1) {
            node = next[0];
        } else {
            if (sp == 0) break;
            node = stack[--sp];
        }
    }
}

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

template <int R, int G, bool STATS, bool BVH>
__global__ void __launch_bounds__(128) rz_path_kernel(const RzPathArgs a) {
    extern __shared__ __align__(16) unsigned char rz_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    float4 *s_cr = reinterpret_cast<float4 *>(rz_smem);
    float4 *s_vel = s_cr + a.set.n_pad;

    if (!BVH) {
        // Stage the sphere set global -> shared with the bulk async-copy engine.
        const uint32_t bytes_cr = a.set.n_pad * 16u;
        const uint32_t bytes_vel = (a.set.n_pad - a.set.n_static_pad) * 16u;
        if (threadIdx.x == 0) rz_mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            rz_mbar_expect_tx(&s_bar, bytes_cr + bytes_vel);
            rz_bulk_g2s(s_cr, a.set.cr, bytes_cr, &s_bar);
            if (bytes_vel) rz_bulk_g2s(s_vel, a.set.vel + a.set.n_static_pad, bytes_vel, &s_bar);
        }
        rz_mbar_wait(&s_bar, 0);
    }

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
        if (BVH) {
#pragma unroll
            for (int r = 0; r < R; r++)
                if (st[r].alive) rz_search_bvh(a, rays[r], a.t_min, bt[r], bk[r], c_nodes, c_sph);
        } else {
            rz_search_brute<R, G>(s_cr, s_vel, (int)a.set.n_static_pad, (int)a.set.n_pad, rays, a.t_min, bt, bk);
        }

        // ---------------------------------------------------------------- shade
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (!st[r].alive) continue;
            if (STATS) c_segs++;
            if (bk[r] < 0) {
                // miss: background (renderer.zig:124-125), path ends
                const float3 L = st[r].thr * rz_sky(st[r].ray.d);
                unsigned long long *acc = a.accum + (size_t)st[r].lp * 4u;
                atomicAdd(acc + 0, __float2ull_rn(fminf(fmaxf(L.x, 0.f), 1048576.f) * 4294967296.f));
                atomicAdd(acc + 1, __float2ull_rn(fminf(fmaxf(L.y, 0.f), 1048576.f) * 4294967296.f));
                atomicAdd(acc + 2, __float2ull_rn(fminf(fmaxf(L.z, 0.f), 1048576.f) * 4294967296.f));
                st[r].alive = false;
                if (STATS) c_sky++;
                continue;
            }
            const int k = bk[r] & ~RZ_FAR_BIT;
            const RzHit h = rz_refine_hit(a.set, st[r].ray, k, (bk[r] & RZ_FAR_BIT) != 0);
            const uint32_t mat = a.set.mat[k];
            const uint32_t kind = a.mats.kind[mat];
            if (STATS) c_hit[kind < 3u ? kind : 0u]++;
            st[r].seg++;
            const uint4 rb = rz_philox(st[r].gpix, st[r].sample, st[r].seg, 0u, a.seed_lo, a.seed_hi);
            const float4 u = make_float4(rz_u01(rb.x >> 8), rz_u01(rb.y >> 8), rz_u01(rb.z >> 8), rz_u01(rb.w >> 8));
            float3 att;
            if (!rz_scatter(a.mats, a.texs, mat, kind, h, k, u, st[r].ray, att)) {
                st[r].alive = false;  // absorbed: contributes black (renderer.zig:109,120)
                if (STATS) c_abs++;
                continue;
            }
            st[r].thr = st[r].thr * att;
            if (st[r].seg >= a.max_depth) {
                st[r].alive = false;  // depth == 0 => black (renderer.zig:104-105)
                if (STATS) c_depth++;
            }
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
template <int R, int G, bool STATS, bool BVH>
static cudaError_t rz_launch_one(const RzPathArgs &a, int sm_count, size_t smem, cudaStream_t stream, int *grid_out) {
    auto kern = rz_path_kernel<R, G, STATS, BVH>;
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

// variant: 1 = brute (K1), 3 = BVH (K3).  rays_per_thread in {1,2}.
extern "C" cudaError_t rz_launch_path(const RzPathArgs *a, int variant, int rays_per_thread, int collect_stats, int sm_count,
                                      cudaStream_t stream, int *grid_out) {
    const bool stats = collect_stats != 0;
    if (variant == 3) {
        return stats ? rz_launch_one<1, 1, true, true>(*a, sm_count, 0, stream, grid_out)
                     : rz_launch_one<1, 1, false, true>(*a, sm_count, 0, stream, grid_out);
    }
    const size_t smem = (size_t)(a->set.n_pad + (a->set.n_pad - a->set.n_static_pad)) * 16u;
    if (rays_per_thread == 1) {
        return stats ? rz_launch_one<1, 4, true, false>(*a, sm_count, smem, stream, grid_out)
                     : rz_launch_one<1, 4, false, false>(*a, sm_count, smem, stream, grid_out);
    }
    return stats ? rz_launch_one<2, 4, true, false>(*a, sm_count, smem, stream, grid_out)
                 : rz_launch_one<2, 4, false, false>(*a, sm_count, smem, stream, grid_out);
}
