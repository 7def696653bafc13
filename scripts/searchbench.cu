// searchbench.cu — standalone microbenchmark of the K1 closest-hit search loop (bench tooling,
// not part of the library).  Times the production search functions of rz_search.cuh on an
// RTOW-shaped synthetic sphere set with realistic ray mixes, next to the FFMA / FFMA2 peak
// kernels of rz_misc.cu, so that one gpurun call compares many (variant, R, G, occupancy) points.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -lineinfo \
//        scripts/searchbench.cu rayz_b200/csrc/rz_misc.cu -o scripts/_build/searchbench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <random>
#include <algorithm>

#include "rz_search_variants.cuh"

extern "C" cudaError_t rz_launch_ffma_peak(float *sink, int grid, int iters, int mode, cudaStream_t stream);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

struct SbArgs {
    RzSphereSet set;
    const float4 *rp;      // ray-paired layout, 2 float4 per sphere
    const float4 *ray_o;   // xyz, w = time
    const float4 *ray_d;   // xyz unit, w = self_k bits
    int iters;
    float t_min;
    int *out_k;
    float *out_t;
};

template <int MODE, int R, int G>
__global__ void __launch_bounds__(128) sb_kernel(const SbArgs a) {
    extern __shared__ __align__(16) unsigned char rz_smem[];
    __shared__ __align__(8) uint64_t s_bar;
    float4 *s0 = reinterpret_cast<float4 *>(rz_smem);
    float4 *s_vel = s0 + a.set.n_pad;
    if (MODE == 0) rz_stage_scene(a.set, s0, s_vel, &s_bar);
    else if (MODE == 1) rz_stage_scene_pk(a.set, s0, &s_bar);
    else if (MODE == 2) {
        for (unsigned i = threadIdx.x; i < 2u * a.set.n_pad; i += blockDim.x) s0[i] = a.rp[i];
        __syncthreads();
    }

    const unsigned gt = blockIdx.x * blockDim.x + threadIdx.x;
    RzRay ray[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const float4 o = a.ray_o[gt * R + r], d = a.ray_d[gt * R + r];
        ray[r].o = f3(o.x, o.y, o.z);
        ray[r].d = f3(d.x, d.y, d.z);
        ray[r].time = o.w;
        ray[r].self_k = __float_as_int(d.w);
    }
    int acc_k = 0;
    float acc_t = 0.f;
#pragma unroll 1
    for (int it = 0; it < a.iters; it++) {
        float bt[R];
        int bk[R];
#pragma unroll
        for (int r = 0; r < R; r++) { bt[r] = 3.0e38f; bk[r] = -1; }
        if (MODE == 0) rz_search_brute<R, G>(s0, s_vel, (int)a.set.n_static_pad, (int)a.set.n_pad, ray, a.t_min, bt, bk);
        else if (MODE == 1) rz_search_brute2<R, G>(s0, (int)a.set.n_static_pad, (int)a.set.n_pad, ray, a.t_min, bt, bk);
        else if (MODE == 3) rz_search_brute2<R, G, RzSrcConst>(RzSrcConst{}, (int)a.set.n_static_pad, (int)a.set.n_pad, ray, a.t_min, bt, bk);
        else rz_search_brute_rp<(R % 2 ? R + 1 : R), G>(s0, (int)a.set.n_static_pad, (int)a.set.n_pad, reinterpret_cast<const RzRay (&)[(R % 2 ? R + 1 : R)]>(ray), a.t_min, reinterpret_cast<float (&)[(R % 2 ? R + 1 : R)]>(bt), reinterpret_cast<int (&)[(R % 2 ? R + 1 : R)]>(bk));
#pragma unroll
        for (int r = 0; r < R; r++) {
            acc_k += bk[r];
            acc_t += bt[r] < 1e30f ? bt[r] : 0.f;
            // data dependence on the result so the loop cannot be hoisted; keeps the ray put
            ray[r].time = fminf(0.999f, ray[r].time + (float)(bk[r] & 1) * 1e-7f);
        }
    }
    a.out_k[gt] = acc_k;
    a.out_t[gt] = acc_t;
}


// ------------------------------------------------------------------ instruction-mix microbenchmarks
// OP: 0 FFMA2 r,r,r   1 FFMA2 with a 32-bit broadcast operand   2 FADD2 (broadcast)   3 FMUL2 (broadcast)
//     4 the search's own mix (per pair: 3 FFMA2 centre, 3 FADD2, FMUL2 + 2 FFMA2, 3 FFMA2, 2 FFMA, 1 FMNMX3), data in registers
//     5 = 4 with the operands re-read from shared memory every iteration (LDS.128 broadcast)
__constant__ float4 c_sph[64];

template <int OP>
__global__ void __launch_bounds__(128) mix_kernel(float *sink, float a, float b, int iters, const float4 *gsrc) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = gsrc[threadIdx.x];
    __syncthreads();
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = make_float2(1.0f + 0.001f * (threadIdx.x + i), 0.5f + 0.002f * i);
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
    float ox = -13.f - 1e-3f * threadIdx.x, oy = -2.f + a, oz = -3.f + b, dx = 0.6f * a, dy = -0.1f * a, dz = 0.79f * a, tm = 0.3f, m = -1.f;
    float2 m2 = make_float2(0.f, 0.f);
    float4 rg[16];
#pragma unroll
    for (int i = 0; i < 16; i++) rg[i] = gsrc[(i * 5 + (threadIdx.x >> 5)) & 63];
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (OP <= 3) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (OP == 0) p[i] = __ffma2_rn(p[i], a2, b2);
                    if (OP == 1) p[i] = __ffma2_rn(p[i], make_float2(a, a), b2);
                    if (OP == 2) p[i] = __fadd2_rn(p[i], make_float2(b, b));
                    if (OP == 3) p[i] = __fmul2_rn(p[i], make_float2(a, a));
                }
            }
        } else {
            // OP >= 4: feature bits of (OP - 4): 1 LDS operands, 2 packed disc (no scalar FFMA), 4 packed max (no FMNMX3), 8 no centre FFMA2
            constexpr int F = OP - 4;
            const float4 *q = sm + ((it & 3) << 3);
#pragma unroll
            for (int j = 0; j < 4; j++) {   // 4 sphere pairs = 8 tests for this one ray
                float4 A, B, VA, VB;
                if (F & 32) { const float4 *cq = c_sph + ((it & 3) << 3); A = cq[2 * j]; B = cq[2 * j + 1]; VA = cq[32 + 2 * j]; VB = cq[32 + 2 * j + 1]; }
                else if (F & 1) { A = q[2 * j]; B = q[2 * j + 1]; VA = q[32 + 2 * j]; VB = q[32 + 2 * j + 1]; }
                else { A = rg[4 * j]; B = rg[4 * j + 1]; VA = rg[4 * j + 2]; VB = rg[4 * j + 3]; }
                const float2 t2 = make_float2(tm, tm);
                float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w), cz = make_float2(B.x, B.y);
                if (!(F & 8) && !(F & 16)) { cx = __ffma2_rn(make_float2(VA.x, VA.y), t2, cx); cy = __ffma2_rn(make_float2(VA.z, VA.w), t2, cy); cz = __ffma2_rn(make_float2(VB.x, VB.y), t2, cz); }
                float2 ocx = __fadd2_rn(cx, make_float2(ox, ox));
                float2 ocy = __fadd2_rn(cy, make_float2(oy, oy));
                float2 ocz = __fadd2_rn(cz, make_float2(oz, oz));
                if (!(F & 8) && (F & 16)) { ocx = __ffma2_rn(make_float2(VA.x, VA.y), t2, ocx); ocy = __ffma2_rn(make_float2(VA.z, VA.w), t2, ocy); ocz = __ffma2_rn(make_float2(VB.x, VB.y), t2, ocz); }
                const float2 bb = __ffma2_rn(ocz, make_float2(dz, dz), __ffma2_rn(ocy, make_float2(dy, dy), __fmul2_rn(ocx, make_float2(dx, dx))));
                const float2 c = __ffma2_rn(ocz, ocz, __ffma2_rn(ocy, ocy, __ffma2_rn(ocx, ocx, make_float2(B.z, B.w))));
                float2 disc;
                if (F & 64) disc = __ffma2_rn(bb, bb, make_float2(__int_as_float(__float_as_int(c.x) ^ 0x80000000), __int_as_float(__float_as_int(c.y) ^ 0x80000000)));
                else if (F & 2) disc = __ffma2_rn(bb, bb, c);
                else disc = make_float2(fmaf(bb.x, bb.x, -c.x), fmaf(bb.y, bb.y, -c.y));
                if (F & 4) m2 = __fadd2_rn(m2, disc);
                else m = fmaxf(m, fmaxf(disc.x, disc.y));
            }
            ox += 1e-6f; tm += 1e-7f;
        }
    }
    float r = m + m2.x + m2.y;
#pragma unroll
    for (int i = 0; i < 8; i++) r += p[i].x + p[i].y;
    if (r == 12345.678f) sink[0] = r;
}

template <int OP>
static void run_mix(float *sink, const float4 *gsrc, int sms, int ctas, double peak) {
    const int grid = sms * ctas, iters = OP <= 3 ? 20000 : 40000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    mix_kernel<OP><<<grid, 128>>>(sink, 0.9999f, 1e-3f, 1000, gsrc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    mix_kernel<OP><<<grid, 128>>>(sink, 0.9999f, 1e-3f, iters, gsrc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // per iteration and thread: OP<=3: 64 packed instructions; OP>=4: 4 pairs x (13 packed + 2 FFMA)
    const int F = OP - 4;
    const double packed = OP <= 3 ? 64.0 : 4.0 * (9 + ((F & 8) ? 0 : 3) + ((F & 66) ? 1 : 0) + ((F & 4) ? 1 : 0)), scalar = OP <= 3 ? 0.0 : ((F & 66) ? 0.0 : 8.0);
    const double pipe_cycles = (packed * 2 + scalar) * iters;                 // FMA-pipe cycles per warp if FFMA2 = 2 cycles
    const double cyc = ms * 1e-3 * 1.965e9;                                   // at 1965 MHz
    const double warps_per_smsp = ctas * 4 / 4.0;
    printf("{\"mix_op\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"fma_pipe_util_if_2cyc\": %.4f, \"issue_per_cycle\": %.4f}\n", OP, ctas, ms,
           pipe_cycles * warps_per_smsp / cyc, (packed + scalar + (OP >= 4 ? ((F & 4) ? 0 : 4) + ((F & 1) ? 16 : 0) + 6 : 3)) * iters * warps_per_smsp / cyc);
    (void)peak;
    fflush(stdout);
}

// cold-operand forms: 8 (V, C) register pairs per thread stay live; every iteration reads them again.
//   FORM 0: acc += fma2(V, t, C)              (two cold 64-bit sources + one 32-bit)
//   FORM 1: acc += fma2(V, t, C + no)         (one cold 64-bit source per instruction)
//   FORM 2: acc  = fma2(V, t, acc)            (one cold)
//   FORM 3: acc += fma2(V, V2, C)             (three cold 64-bit sources)
//   FORM 4: acc += fadd2(V, C)                (two cold, FADD2)
template <int FORM>
__global__ void __launch_bounds__(128) cold_kernel(float *sink, float t, float no, int iters, const float4 *gsrc) {
    float2 V[8], C[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float4 g = gsrc[(i * 7 + (threadIdx.x >> 5)) & 63];
        V[i] = make_float2(g.x, g.y); C[i] = make_float2(g.z, g.w); acc[i] = make_float2(0.f, 0.f);
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (FORM == 0) acc[i] = __fadd2_rn(acc[i], __ffma2_rn(V[i], make_float2(t, t), C[i]));
                if (FORM == 1) acc[i] = __fadd2_rn(acc[i], __ffma2_rn(V[i], make_float2(t, t), __fadd2_rn(C[i], make_float2(no, no))));
                if (FORM == 2) acc[i] = __ffma2_rn(V[i], make_float2(t, t), acc[i]);
                if (FORM == 3) acc[i] = __fadd2_rn(acc[i], __ffma2_rn(V[i], V[(i + 1) & 7], C[i]));
                if (FORM == 4) acc[i] = __fadd2_rn(acc[i], __fadd2_rn(V[i], C[i]));
            }
            t += 1e-7f;
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) r += acc[i].x + acc[i].y;
    if (r == 12345.678f) sink[0] = r;
}

template <int FORM>
static void run_cold(float *sink, const float4 *gsrc, int sms, int ctas) {
    const int grid = sms * ctas, iters = 20000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    cold_kernel<FORM><<<grid, 128>>>(sink, 0.3f, -2.f, 1000, gsrc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    cold_kernel<FORM><<<grid, 128>>>(sink, 0.3f, -2.f, iters, gsrc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double packed = 32.0 * (FORM == 1 ? 3 : FORM == 2 ? 1 : 2);
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("{\"cold_form\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"cycles_per_packed_instr\": %.3f}\n", FORM, ctas, ms, cyc / (packed * iters * ctas));
    fflush(stdout);
}

struct Result { float ms; long long checksum; double tsum; int grid; int per_sm; std::vector<int> hk; std::vector<float> ht; };

template <int MODE, int R, int G>
static Result run(const SbArgs &a0, int sms, int ctas_per_sm_cap, int iters, size_t smem) {
    auto kern = sb_kernel<MODE, R, G>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
    if (ctas_per_sm_cap > 0) per_sm = std::min(per_sm, ctas_per_sm_cap);
    const int grid = sms * per_sm;
    SbArgs a = a0;
    a.iters = iters;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<grid, 128, smem>>>(a);  // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        kern<<<grid, 128, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    std::vector<int> hk((size_t)grid * 128);
    std::vector<float> ht((size_t)grid * 128);
    CK(cudaMemcpy(hk.data(), a.out_k, hk.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ht.data(), a.out_t, ht.size() * 4, cudaMemcpyDeviceToHost));
    Result r;
    r.ms = best; r.grid = grid; r.per_sm = per_sm; r.checksum = 0; r.tsum = 0;
    // checksum over the first 148*128 threads' rays only would depend on R; use per-ray normalisation:
    for (size_t i = 0; i < hk.size(); i++) { r.checksum += hk[i]; r.tsum += ht[i]; }
    r.hk = hk; r.ht = ht;
    return r;
}

int main(int argc, char **argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 400;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;

    // ---- FP32 peak, per mode (0 scalar 2-reg chains, 1 SGEMM-like 3-reg, 2 packed FFMA2)
    float *sink;
    CK(cudaMalloc(&sink, 64));
    double peak = 0;
    for (int mode = 0; mode < 3; mode++) {
        const int grid = sms * 8, it = 40000;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(rz_launch_ffma_peak(sink, grid, 2000, mode, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        CK(rz_launch_ffma_peak(sink, grid, it, mode, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 128.0 * it * 256.0 * grid / (ms * 1e-3) / 1e12;
        printf("{\"ffma_peak_mode\": %d, \"tflops\": %.3f, \"ms\": %.3f}\n", mode, tf, ms);
        peak = std::max(peak, tf);
    }

    {
        std::vector<float4> src(64);
        for (int i = 0; i < 64; i++) src[i] = i < 32 ? make_float4(i * 0.37f - 5, i * 0.37f - 4.5f, 0.2f, 0.2f) : make_float4(0.f, 0.f, 0.01f * i, 0.02f * i);
        for (int i = 1; i < 32; i += 2) src[i] = make_float4(i * 0.21f - 3, i * 0.21f - 2.7f, -0.04f, -0.04f);
        float4 *gsrc;
        CK(cudaMalloc(&gsrc, 64 * 16));
        CK(cudaMemcpy(gsrc, src.data(), 64 * 16, cudaMemcpyHostToDevice));
        CK(cudaMemcpyToSymbol(c_sph, src.data(), 64 * 16));
        for (int ctas : {4, 8}) {
            run_mix<0>(sink, gsrc, sms, ctas, peak); run_mix<1>(sink, gsrc, sms, ctas, peak); run_mix<2>(sink, gsrc, sms, ctas, peak);
            run_mix<3>(sink, gsrc, sms, ctas, peak); run_mix<4>(sink, gsrc, sms, ctas, peak); run_mix<5>(sink, gsrc, sms, ctas, peak);
            run_mix<4 + 16>(sink, gsrc, sms, ctas, peak); run_mix<4 + 17>(sink, gsrc, sms, ctas, peak); run_mix<4 + 16 + 64>(sink, gsrc, sms, ctas, peak); run_mix<4 + 17 + 64>(sink, gsrc, sms, ctas, peak); run_mix<4 + 16 + 2>(sink, gsrc, sms, ctas, peak); run_mix<4 + 16 + 6>(sink, gsrc, sms, ctas, peak);
            run_mix<4 + 8>(sink, gsrc, sms, ctas, peak); run_mix<4 + 9>(sink, gsrc, sms, ctas, peak);
        }
    }
    if (argc > 2 && atoi(argv[2]) == 2) {
        float4 *g2;
        CK(cudaMalloc(&g2, 64 * 16));
        std::vector<float4> src(64);
        for (int i = 0; i < 64; i++) src[i] = make_float4(0.01f * i, 0.02f * i, i * 0.37f - 5, i * 0.37f - 4.5f);
        CK(cudaMemcpy(g2, src.data(), 64 * 16, cudaMemcpyHostToDevice));
        for (int ctas : {4, 8}) { run_cold<0>(sink, g2, sms, ctas); run_cold<1>(sink, g2, sms, ctas); run_cold<2>(sink, g2, sms, ctas); run_cold<3>(sink, g2, sms, ctas); run_cold<4>(sink, g2, sms, ctas); }
        return 0;
    }
    if (argc > 2 && atoi(argv[2]) == 1) return 0;

    // ---- RTOW-shaped sphere set: ground + 3 big (static) + 22x22 small, 80 % moving in y
    std::mt19937 rng(42);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    struct S { float c[3], v[3], r; };
    std::vector<S> st, mv;
    st.push_back({{0, -1000, 0}, {0, 0, 0}, 1000});
    st.push_back({{0, 1, 0}, {0, 0, 0}, 1});
    st.push_back({{-4, 1, 0}, {0, 0, 0}, 1});
    st.push_back({{4, 1, 0}, {0, 0, 0}, 1});
    for (int aa = -11; aa < 11; aa++)
        for (int bb = -11; bb < 11; bb++) {
            const float m = U(rng);
            S s = {{aa + 0.9f * U(rng), 0.2f, bb + 0.9f * U(rng)}, {0, 0, 0}, 0.2f};
            const float dx = s.c[0] - 4, dz = s.c[2];
            if (sqrtf(dx * dx + dz * dz) <= 0.9f) continue;
            if (m < 0.8f) { s.v[1] = 0.5f * U(rng); mv.push_back(s); } else st.push_back(s);
        }
    const uint32_t n_static = (uint32_t)st.size(), n_moving = (uint32_t)mv.size();
    const uint32_t n_static_pad = (n_static + 3u) & ~3u;
    const uint32_t n_pad = n_static_pad + ((n_moving + 3u) & ~3u);
    std::vector<float4> cr(n_pad, make_float4(0, 0, 0, 1.0f)), vel(n_pad, make_float4(0, 0, 0, 0));
    for (uint32_t i = 0; i < n_static; i++) { cr[i] = make_float4(st[i].c[0], st[i].c[1], st[i].c[2], -st[i].r * st[i].r); vel[i] = make_float4(0, 0, 0, st[i].r); }
    for (uint32_t i = 0; i < n_moving; i++) {
        cr[n_static_pad + i] = make_float4(mv[i].c[0], mv[i].c[1], mv[i].c[2], -mv[i].r * mv[i].r);
        vel[n_static_pad + i] = make_float4(mv[i].v[0], mv[i].v[1], mv[i].v[2], mv[i].r);
    }
    std::vector<float4> pk;
    for (uint32_t k = 0; k < n_pad; k += 2) {
        const float4 a = cr[k], b = cr[k + 1];
        pk.push_back(make_float4(a.x, b.x, a.y, b.y));
        pk.push_back(make_float4(a.z, b.z, a.w, b.w));
        if (k >= n_static_pad) {
            const float4 va = vel[k], vb = vel[k + 1];
            pk.push_back(make_float4(va.x, vb.x, va.y, vb.y));
            pk.push_back(make_float4(va.z, vb.z, 0.f, 0.f));
        }
    }
    // ---- rays: 40 % camera rays, 60 % bounce rays leaving the ground (self_k = 0) or a small sphere
    const size_t max_threads = (size_t)sms * 16 * 128;
    const size_t n_rays = max_threads * 4;
    std::vector<float4> ro(n_rays), rd(n_rays);
    for (size_t i = 0; i < n_rays; i++) {
        float o[3], d[3];
        int self = -1;
        if (U(rng) < 0.4f) {
            o[0] = 13 + 0.05f * (U(rng) - 0.5f); o[1] = 2 + 0.05f * (U(rng) - 0.5f); o[2] = 3;
            const float tx = -11 + 22 * U(rng), tz = -11 + 22 * U(rng), ty = 3.0f * U(rng) * U(rng);
            d[0] = tx - o[0]; d[1] = ty - o[1]; d[2] = tz - o[2];
        } else {
            o[0] = -11 + 22 * U(rng); o[1] = 0.0f; o[2] = -11 + 22 * U(rng);
            self = 0;
            const float z = U(rng), ph = 6.2831853f * U(rng), rr = sqrtf(1 - z * z);
            d[0] = rr * cosf(ph); d[1] = z; d[2] = rr * sinf(ph);
        }
        if (argc > 3 && atoi(argv[3]) == 1) { o[0] = 30 * U(rng); o[1] = 100; o[2] = 30 * U(rng); d[0] = U(rng) - 0.5f; d[1] = 1; d[2] = U(rng) - 0.5f; self = -1; }
        const float il = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        ro[i] = make_float4(o[0], o[1], o[2], U(rng));
        float selff; memcpy(&selff, &self, 4);
        rd[i] = make_float4(d[0] * il, d[1] * il, d[2] * il, selff);
    }
    std::vector<float4> rp;
    for (uint32_t k = 0; k < n_pad; k++) {
        rp.push_back(make_float4(cr[k].x, cr[k].y, cr[k].z, vel[k].x));
        rp.push_back(make_float4(vel[k].y, vel[k].z, cr[k].w, cr[k].w));
    }
    float4 *d_rp;
    CK(cudaMalloc(&d_rp, rp.size() * 16));
    CK(cudaMemcpy(d_rp, rp.data(), rp.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(rz_c_pk, pk.data(), pk.size() * 16));
    float4 *d_cr, *d_vel, *d_pk, *d_ro, *d_rd;
    int *d_k; float *d_t;
    CK(cudaMalloc(&d_cr, cr.size() * 16)); CK(cudaMalloc(&d_vel, vel.size() * 16)); CK(cudaMalloc(&d_pk, pk.size() * 16));
    CK(cudaMalloc(&d_ro, ro.size() * 16)); CK(cudaMalloc(&d_rd, rd.size() * 16));
    CK(cudaMalloc(&d_k, max_threads * 4)); CK(cudaMalloc(&d_t, max_threads * 4));
    CK(cudaMemcpy(d_cr, cr.data(), cr.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_vel, vel.data(), vel.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pk, pk.data(), pk.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ro, ro.data(), ro.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_rd, rd.data(), rd.size() * 16, cudaMemcpyHostToDevice));
    SbArgs a;
    memset(&a, 0, sizeof a);
    a.set.cr = d_cr; a.set.vel = d_vel; a.set.pk = d_pk;
    a.set.n = n_static + n_moving; a.set.n_static = n_static; a.set.n_static_pad = n_static_pad; a.set.n_pad = n_pad;
    a.rp = d_rp; a.ray_o = d_ro; a.ray_d = d_rd; a.t_min = 1e-4f; a.out_k = d_k; a.out_t = d_t;
    const size_t smem = (size_t)(2 * n_pad) * 16u;
    const double flop_per_search = 16.0 * n_static + 22.0 * n_moving;
    printf("{\"n_static\": %u, \"n_moving\": %u, \"smem\": %zu, \"flop_per_search\": %.0f, \"peak_tflops\": %.3f}\n", n_static, n_moving, smem, flop_per_search, peak);

#define RUN(MODE, R, G, CAP)                                                                                          \
    {                                                                                                                 \
        Result r = run<MODE, R, G>(a, sms, CAP, iters, smem);                                                         \
        const double searches = (double)r.grid * 128.0 * R * iters;                                                   \
        const double tf = searches * flop_per_search / (r.ms * 1e-3) / 1e12;                                          \
        printf("{\"mode\": %d, \"R\": %d, \"G\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"Gsearch_s\": %.4f, \"tflops\": %.2f, \"frac\": %.4f, \"k_per_search\": %.4f, \"t_per_search\": %.5f}\n", \
               MODE, R, G, r.per_sm, r.ms, searches / (r.ms * 1e-3) / 1e9, tf, tf / peak, (double)r.checksum / searches, r.tsum / searches); \
        fflush(stdout);                                                                                               \
    }
    RUN(0, 2, 4, 0)
    RUN(1, 2, 2, 0)
    RUN(3, 2, 2, 0)
    RUN(3, 2, 1, 0)
    RUN(3, 1, 2, 0)
    RUN(3, 3, 2, 0)
    RUN(3, 4, 2, 0)
    RUN(3, 4, 1, 0)
    RUN(3, 2, 2, 5)
    RUN(3, 2, 2, 4)
    RUN(3, 2, 2, 3)
    RUN(3, 4, 2, 3)
    RUN(3, 4, 2, 2)
    {   // the packed search must return bit-identical (t, k) to the scalar one
        Result r0 = run<3, 2, 2>(a, sms, 2, 50, smem), r1 = run<1, 2, 2>(a, sms, 2, 50, smem), r2 = run<1, 4, 2>(a, sms, 2, 50, smem);
        Result r3 = run<0, 1, 4>(a, sms, 2, 50, smem), r4 = run<1, 1, 4>(a, sms, 2, 50, smem);
        size_t bad = 0, bad2 = 0;
        for (size_t i = 0; i < r0.hk.size(); i++) bad += (r0.hk[i] != r1.hk[i]) || (r0.ht[i] != r1.ht[i]);
        for (size_t i = 0; i < r3.hk.size(); i++) bad2 += (r3.hk[i] != r4.hk[i]) || (r3.ht[i] != r4.ht[i]);
        printf("{\"packed_vs_scalar_mismatching_threads\": %zu, \"r1\": %zu, \"of\": %zu, \"r4_checksum\": %lld}\n", bad, bad2, r0.hk.size(), r2.checksum);
    }
    return 0;
}
