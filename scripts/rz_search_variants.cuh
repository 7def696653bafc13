// rz_search_variants.cuh — (scripts only; round-1 forms with the textbook discriminant b^2 - c: since round 2 the product
// uses the cancellation-free test rz_sphere_test, so these are no longer bit-identical to it, only rate comparisons)
// forms of the K1 search that were measured and NOT adopted; only scripts/searchbench.cu
// includes this header (profiles/r01_search_microbench.md holds the numbers).  All of them return bit-identical (t, k).
//   rz_search_brute     scalar FFMA form of the first round-1 kernel (47 % of FP32 peak)
//   RzSrcConst          packed search with the operands from the constant bank (LDCU -> UR.F32x2): LDCU-rate bound
//   rz_search_brute_rp  packed by ray pairs instead of sphere pairs: same rate as the production form
#pragma once
#include "../rayz_b200/csrc/rz_search.cuh"

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
                    if (disc[r][j] > 0.0f) rz_consider(i + j, -b[r][j], -disc[r][j], ray[r].self_k, t_min, bt[r], bk[r]);
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
                // oc = (c0 - o) + v * time; centre(t) = center.origin + center.dir * ray.time (geom.zig:40)
                const float ocx = fmaf(v[j].x, ray[r].time, s[j].x - ray[r].o.x);
                const float ocy = fmaf(v[j].y, ray[r].time, s[j].y - ray[r].o.y);
                const float ocz = fmaf(v[j].z, ray[r].time, s[j].z - ray[r].o.z);
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
                    if (disc[r][j] > 0.0f) rz_consider(i + j, -b[r][j], -disc[r][j], ray[r].self_k, t_min, bt[r], bk[r]);
        }
    }
}

// Operand sources for the packed search.  The sphere operands are WARP-UNIFORM (every lane tests
// the same spheres), so besides shared memory (LDS.128 broadcasts into vector registers) they can
// come from the constant bank through the uniform datapath: LDCU.64 into uniform registers, which
// FFMA2/FADD2 take directly as `UR.F32x2` operands.  That removes the LDS register-file write-back
// and the "cold" vector-register operand reads that hold the FMA pipe at ~67 % in the LDS form
// (measured: scripts/searchbench.cu mix_kernel, 67 % -> 80 % FMA-pipe utilisation).
#ifndef RZ_CONST_PK_FLOAT4
#define RZ_CONST_PK_FLOAT4 4000   // 64,000 B of the 64 KB constant bank: 2000 moving or 4000 stationary spheres
#endif
static __constant__ float4 rz_c_pk[RZ_CONST_PK_FLOAT4];   // one instance per translation unit that includes this header

struct RzSrcConst {
    __device__ __forceinline__ float4 operator[](int i) const { return rz_c_pk[i]; }
};

// Entry into the rare path.  A warp-uniform form (__any_sync, no BSSY/BSYNC pair per iteration) was
// measured with scripts/searchbench.cu: no difference (53.2 % vs 53.1 % of FP32 peak), so the plain
// per-lane branch stays.
// ------------------------------------------------------------------------------ K1 search, ray-paired
// Transposed packing: one packed instruction works on TWO RAYS (.x/.y halves) against one sphere,
// whose numbers enter as 32-bit broadcast operands straight from the LDS destination registers.
// The 64-bit operands are then the long-lived ray registers (-o, d, time as pairs), which G
// consecutive instructions share through the operand-reuse cache.  Sphere record in shared memory
// = 2 x float4: (cx, cy, cz, vx) (vy, vz, w, w), w = -r^2 (stationary: v = 0 and not read).
// Same operation order per half as rz_search_brute2 => bit-identical results.
template <int R, int G>
__device__ __forceinline__ void rz_search_brute_rp(const float4 *__restrict__ s_rp, int n_static_pad, int n_pad,
                                                   const RzRay (&ray)[R], float t_min, float (&bt)[R], int (&bk)[R]) {
    static_assert(R % 2 == 0, "ray-paired search needs an even number of rays per thread");
    constexpr int P = R / 2;
    float2 nox[P], noy[P], noz[P], dx[P], dy[P], dz[P], tm[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        nox[p] = rz_f2(-ray[2 * p].o.x, -ray[2 * p + 1].o.x); noy[p] = rz_f2(-ray[2 * p].o.y, -ray[2 * p + 1].o.y);
        noz[p] = rz_f2(-ray[2 * p].o.z, -ray[2 * p + 1].o.z);
        dx[p] = rz_f2(ray[2 * p].d.x, ray[2 * p + 1].d.x); dy[p] = rz_f2(ray[2 * p].d.y, ray[2 * p + 1].d.y);
        dz[p] = rz_f2(ray[2 * p].d.z, ray[2 * p + 1].d.z); tm[p] = rz_f2(ray[2 * p].time, ray[2 * p + 1].time);
    }
    const float4 *q = s_rp;
    int k = 0;
#pragma unroll 1
    for (; k < n_static_pad; k += G, q += 2 * G) {
        float4 A[G];
        float2 W[G];
#pragma unroll
        for (int j = 0; j < G; j++) { A[j] = q[2 * j]; W[j] = *reinterpret_cast<const float2 *>(&q[2 * j + 1].z); }
        float2 b[P][G], disc[P][G];
        float m = -1.0f;
#pragma unroll
        for (int p = 0; p < P; p++) {
#pragma unroll
            for (int j = 0; j < G; j++) {
                const float2 ocx = __fadd2_rn(nox[p], rz_f2(A[j].x, A[j].x));
                const float2 ocy = __fadd2_rn(noy[p], rz_f2(A[j].y, A[j].y));
                const float2 ocz = __fadd2_rn(noz[p], rz_f2(A[j].z, A[j].z));
                b[p][j] = __ffma2_rn(ocz, dz[p], __ffma2_rn(ocy, dy[p], __fmul2_rn(ocx, dx[p])));
                const float2 c = __ffma2_rn(ocz, ocz, __ffma2_rn(ocy, ocy, __ffma2_rn(ocx, ocx, W[j])));
                disc[p][j] = rz_f2(fmaf(b[p][j].x, b[p][j].x, -c.x), fmaf(b[p][j].y, b[p][j].y, -c.y));
                m = fmaxf(m, fmaxf(disc[p][j].x, disc[p][j].y));
            }
        }
        if (m > 0.0f) {
#pragma unroll
            for (int p = 0; p < P; p++)
#pragma unroll
                for (int j = 0; j < G; j++) {
                    if (disc[p][j].x > 0.0f) rz_consider(k + j, -b[p][j].x, -disc[p][j].x, ray[2 * p].self_k, t_min, bt[2 * p], bk[2 * p]);
                    if (disc[p][j].y > 0.0f) rz_consider(k + j, -b[p][j].y, -disc[p][j].y, ray[2 * p + 1].self_k, t_min, bt[2 * p + 1], bk[2 * p + 1]);
                }
        }
    }
#pragma unroll 1
    for (; k < n_pad; k += G, q += 2 * G) {
        float4 A[G], B[G];
#pragma unroll
        for (int j = 0; j < G; j++) { A[j] = q[2 * j]; B[j] = q[2 * j + 1]; }
        float2 b[P][G], disc[P][G];
        float m = -1.0f;
#pragma unroll
        for (int p = 0; p < P; p++) {
#pragma unroll
            for (int j = 0; j < G; j++) {
                // oc = (c0 - o) + v * time   (centre(t) = center.origin + center.dir * ray.time, geom.zig:40)
                const float2 ocx = __ffma2_rn(tm[p], rz_f2(A[j].w, A[j].w), __fadd2_rn(nox[p], rz_f2(A[j].x, A[j].x)));
                const float2 ocy = __ffma2_rn(tm[p], rz_f2(B[j].x, B[j].x), __fadd2_rn(noy[p], rz_f2(A[j].y, A[j].y)));
                const float2 ocz = __ffma2_rn(tm[p], rz_f2(B[j].y, B[j].y), __fadd2_rn(noz[p], rz_f2(A[j].z, A[j].z)));
                b[p][j] = __ffma2_rn(ocz, dz[p], __ffma2_rn(ocy, dy[p], __fmul2_rn(ocx, dx[p])));
                const float2 c = __ffma2_rn(ocz, ocz, __ffma2_rn(ocy, ocy, __ffma2_rn(ocx, ocx, rz_f2(B[j].z, B[j].w))));
                disc[p][j] = rz_f2(fmaf(b[p][j].x, b[p][j].x, -c.x), fmaf(b[p][j].y, b[p][j].y, -c.y));
                m = fmaxf(m, fmaxf(disc[p][j].x, disc[p][j].y));
            }
        }
        if (m > 0.0f) {
#pragma unroll
            for (int p = 0; p < P; p++)
#pragma unroll
                for (int j = 0; j < G; j++) {
                    if (disc[p][j].x > 0.0f) rz_consider(k + j, -b[p][j].x, -disc[p][j].x, ray[2 * p].self_k, t_min, bt[2 * p], bk[2 * p]);
                    if (disc[p][j].y > 0.0f) rz_consider(k + j, -b[p][j].y, -disc[p][j].y, ray[2 * p + 1].self_k, t_min, bt[2 * p + 1], bk[2 * p + 1]);
                }
        }
    }
}

// ------------------------------------------------------------------------------ stage scene (cr / vel layout)
// Stage the brute-force sphere set global -> shared with the bulk async-copy engine
// (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP).  All threads of the CTA must call.
__device__ __forceinline__ void rz_stage_scene(const RzSphereSet &set, float4 *s_cr, float4 *s_vel, uint64_t *bar) {
    const uint32_t bytes_cr = set.n_pad * 16u;
    const uint32_t bytes_vel = (set.n_pad - set.n_static_pad) * 16u;
    if (threadIdx.x == 0) rz_mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        rz_mbar_expect_tx(bar, bytes_cr + bytes_vel);
        rz_bulk_g2s(s_cr, set.cr, bytes_cr, bar);
        if (bytes_vel) rz_bulk_g2s(s_vel, set.vel + set.n_static_pad, bytes_vel, bar);
    }
    rz_mbar_wait(bar, 0);
}

