// Stand-alone probe for the brute-force scan kernel (the FP32-issue roofline kernel, wf_scan_*).
// Answers, on the GPU it runs on:
//   1. the lane-instruction rate of scalar FFMA, of packed FFMA2 (fma.rn.f32x2) and of FFMA2 with
//      FMNMX3 / LDS.128 mixed in at the ratio the filter loop needs;
//   2. the rate of the filter loop itself (10 FP32-pipe lane-instructions per (ray, sphere), the
//      accounting of DESIGN.md) over a shared-memory tile, for 1/2/4 rays per thread, scalar and packed.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scan_probe tools/scan_probe.cu
// Not part of the library; the numbers it printed are recorded in DESIGN.md.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 neg2(u64 a) { float lo, hi; upk(a, lo, hi); return pk(-lo, -hi); }

// ---- 1. raw pipes
__global__ void __launch_bounds__(256) k_ffma(float *out, int iters, const float *in)
{
    float a[8];
    for (int k = 0; k < 8; k++) a[k] = in[0] + threadIdx.x + k;
    const float m0 = in[1], m1 = in[2], c0 = in[3], c1 = in[4];
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = __fmaf_rn(a[j], (j & 1) ? m1 : m0, (j & 2) ? c1 : c0);
        }
    }
    float s = 0;
    for (int k = 0; k < 8; k++) s += a[k];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float *out, int iters, const float *in)
{
    u64 a[8];
    for (int k = 0; k < 8; k++) a[k] = pk(in[0] + threadIdx.x + k, in[0] + k);
    const u64 m0 = pk(in[1], in[2]), m1 = pk(in[2], in[1]), c0 = pk(in[3], in[4]), c1 = pk(in[4], in[3]);
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int j = 0; j < 8; j++) a[j] = fma2(a[j], (j & 1) ? m1 : m0, (j & 2) ? c1 : c0);
        }
    }
    float s = 0;
    for (int k = 0; k < 8; k++) { float lo, hi; upk(a[k], lo, hi); s += lo + hi; }
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 10 FFMA2 : 1 FMNMX3 : 1 LDS.128 (the mix of the packed filter loop, two rays per thread)
__global__ void __launch_bounds__(256) k_mix(float *out, int iters, const float *in)
{
    __shared__ float4 sm[256];
    sm[threadIdx.x] = make_float4(in[1], in[2], in[3], in[4]);
    __syncthreads();
    u64 a[10];
    for (int k = 0; k < 10; k++) a[k] = pk(in[0] + threadIdx.x + k, in[0] + k);
    float mx = -1e38f;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float4 s = sm[(i + k) & 255];                    // broadcast LDS.128
            const u64 m0 = pk(s.x, s.y), c0 = pk(s.z, s.w);
#pragma unroll
            for (int j = 0; j < 10; j++) a[j] = fma2(a[j], m0, c0);
            float lo, hi;
            upk(a[k % 10], lo, hi);
            mx = fmaxf(fmaxf(mx, lo), hi);
        }
    }
    float s = mx;
    for (int k = 0; k < 10; k++) { float lo, hi; upk(a[k], lo, hi); s += lo + hi; }
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 with scalar FFMA / FMNMX mixed in from independent chains: does anything issue in the shadow of a packed op?
template <int N2, int N1, int NM>
__global__ void __launch_bounds__(256) k_mix2(float *out, int iters, const float *in)
{
    u64 a[N2 > 0 ? N2 : 1];
    float b[N1 > 0 ? N1 : 1], c[NM > 0 ? NM : 1];
    for (int k = 0; k < N2; k++) a[k] = pk(in[0] + threadIdx.x + k, in[0] + k);
    for (int k = 0; k < N1; k++) b[k] = in[0] + threadIdx.x + 0.5f * k;
    for (int k = 0; k < NM; k++) c[k] = in[0] + threadIdx.x + 0.25f * k;
    const u64 m2 = pk(in[1], in[2]), c2 = pk(in[3], in[4]);
    const float m1 = in[1], c1 = in[3], lim = in[2];
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
#pragma unroll
            for (int j = 0; j < N2; j++) a[j] = fma2(a[j], m2, c2);
#pragma unroll
            for (int j = 0; j < N1; j++) b[j] = __fmaf_rn(b[j], m1, c1);
#pragma unroll
            for (int j = 0; j < NM; j++) c[j] = fmaxf(fminf(c[j], lim), c[(j + 1) % NM] * 1.0f);
        }
    }
    float s = 0;
    for (int k = 0; k < N2; k++) { float lo, hi; upk(a[k], lo, hi); s += lo + hi; }
    for (int k = 0; k < N1; k++) s += b[k];
    for (int k = 0; k < NM; k++) s += c[k];
    if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 2. the filter loop over a resident tile
// scalar: tile of float4 {cx, cy, cz, R}; NR rays per thread
template <int NR>
__global__ void __launch_bounds__(128) k_scan_scalar(const float4 *__restrict__ g, int n, int reps, const float *__restrict__ rays, float *out)
{
    extern __shared__ float4 tile[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = g[i];
    __syncthreads();
    float ox[NR], oy[NR], oz[NR], dx[NR], dy[NR], dz[NR], m[NR];
    for (int r = 0; r < NR; r++) {
        const float *p = rays + 6 * ((blockIdx.x * blockDim.x + threadIdx.x) * NR + r);
        ox[r] = p[0]; oy[r] = p[1]; oz[r] = p[2]; dx[r] = p[3]; dy[r] = p[4]; dz[r] = p[5];
        m[r] = -3e38f;
    }
    int hits = 0;
#pragma unroll 1
    for (int rep = 0; rep < reps; rep++) {
#pragma unroll 1
        for (int k0 = 0; k0 < n; k0 += 16) {
            float mg[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) mg[r] = -3e38f;
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
                const float4 s0 = tile[k0 + u], s1 = tile[k0 + u + 1];
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    const float cx0 = s0.x - ox[r], cy0 = s0.y - oy[r], cz0 = s0.z - oz[r];
                    const float cx1 = s1.x - ox[r], cy1 = s1.y - oy[r], cz1 = s1.z - oz[r];
                    const float b0 = __fmaf_rn(dz[r], cz0, __fmaf_rn(dy[r], cy0, dx[r] * cx0));
                    const float b1 = __fmaf_rn(dz[r], cz1, __fmaf_rn(dy[r], cy1, dx[r] * cx1));
                    const float w0 = __fmaf_rn(cx0, cx0, __fmaf_rn(cy0, cy0, __fmaf_rn(cz0, cz0, -s0.w)));
                    const float w1 = __fmaf_rn(cx1, cx1, __fmaf_rn(cy1, cy1, __fmaf_rn(cz1, cz1, -s1.w)));
                    const float v0 = __fmaf_rn(b0, b0, -w0), v1 = __fmaf_rn(b1, b1, -w1);
                    mg[r] = fmaxf(fmaxf(mg[r], v0), v1);
                }
            }
            bool any = false;
#pragma unroll
            for (int r = 0; r < NR; r++) any |= !(mg[r] < 0.f);
            if (any) {
#pragma unroll
                for (int r = 0; r < NR; r++) { if (!(mg[r] < 0.f)) { hits++; m[r] = fmaxf(m[r], mg[r]); } }
            }
        }
    }
    float s = (float)hits;
    for (int r = 0; r < NR; r++) s += m[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed: tile of pairs {cx0,cx1,cy0,cy1 | cz0,cz1,R0,R1}; NR rays per thread
template <int NR>
__global__ void __launch_bounds__(128) k_scan_packed(const float4 *__restrict__ g, int n, int reps, const float *__restrict__ rays, float *out)
{
    extern __shared__ float4 tile[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = g[i];      // n float4 = n/2 pairs... caller passes packed layout
    __syncthreads();
    u64 ox[NR], oy[NR], oz[NR], dx[NR], dy[NR], dz[NR];
    float m[NR];
    for (int r = 0; r < NR; r++) {
        const float *p = rays + 6 * ((blockIdx.x * blockDim.x + threadIdx.x) * NR + r);
        ox[r] = pk(p[0], p[0]); oy[r] = pk(p[1], p[1]); oz[r] = pk(p[2], p[2]);
        dx[r] = pk(p[3], p[3]); dy[r] = pk(p[4], p[4]); dz[r] = pk(p[5], p[5]);
        m[r] = -3e38f;
    }
    int hits = 0;
#pragma unroll 1
    for (int rep = 0; rep < reps; rep++) {
#pragma unroll 1
        for (int k0 = 0; k0 < n; k0 += 16) {                              // 16 spheres = 8 pairs = 16 float4
            float mg[NR];
#pragma unroll
            for (int r = 0; r < NR; r++) mg[r] = -3e38f;
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
                const float4 a = tile[k0 + u], b4 = tile[k0 + u + 1];
                const u64 CX = pk(a.x, a.y), CY = pk(a.z, a.w), CZ = pk(b4.x, b4.y), R = pk(b4.z, b4.w);
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    const u64 cx = sub2(CX, ox[r]), cy = sub2(CY, oy[r]), cz = sub2(CZ, oz[r]);
                    const u64 b = fma2(dz[r], cz, fma2(dy[r], cy, mul2(dx[r], cx)));
                    const u64 nw = fma2(neg2(cx), cx, fma2(neg2(cy), cy, fma2(neg2(cz), cz, R)));
                    const u64 v = fma2(b, b, nw);
                    float v0, v1;
                    upk(v, v0, v1);
                    mg[r] = fmaxf(fmaxf(mg[r], v0), v1);
                }
            }
            bool any = false;
#pragma unroll
            for (int r = 0; r < NR; r++) any |= !(mg[r] < 0.f);
            if (any) {
#pragma unroll
                for (int r = 0; r < NR; r++) { if (!(mg[r] < 0.f)) { hits++; m[r] = fmaxf(m[r], mg[r]); } }
            }
        }
    }
    float s = (float)hits;
    for (int r = 0; r < NR; r++) s += m[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, %d kHz\n", prop.name, sms, prop.clockRate);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float *out, *in;
    CK(cudaMalloc(&out, (size_t)sms * 64 * 256 * sizeof(float)));
    CK(cudaMalloc(&in, 8 * sizeof(float)));
    const float host_in[8] = {1.0f, 0.9999f, 0.99991f, 0.0001f, 0.00011f, 0, 0, 0};
    CK(cudaMemcpy(in, host_in, sizeof host_in, cudaMemcpyHostToDevice));
    const int iters = 4096;
    auto report = [&](const char *name, double lane_ops, double ms) {
        printf("%-44s %8.3f ms  %9.1f Glane-instr/s\n", name, ms, lane_ops / (ms * 1e-3) / 1e9);
    };
    for (int blocks_per_sm : {2, 4, 8}) {
        const int blocks = sms * blocks_per_sm;
        double best[3] = {1e30, 1e30, 1e30};
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0)); k_ffma<<<blocks, 256>>>(out, iters, in); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best[0] = std::min(best[0], time_ms(e0, e1));
            CK(cudaEventRecord(e0)); k_ffma2<<<blocks, 256>>>(out, iters, in); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best[1] = std::min(best[1], time_ms(e0, e1));
            CK(cudaEventRecord(e0)); k_mix<<<blocks, 256>>>(out, iters, in); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best[2] = std::min(best[2], time_ms(e0, e1));
        }
        char nm[96];
        snprintf(nm, sizeof nm, "FFMA  rrr, %d blocks/SM", blocks_per_sm);
        report(nm, (double)blocks * 256 * iters * 128, best[0]);
        snprintf(nm, sizeof nm, "FFMA2 (lanes x2), %d blocks/SM", blocks_per_sm);
        report(nm, (double)blocks * 256 * iters * 128 * 2, best[1]);
        snprintf(nm, sizeof nm, "10 FFMA2 : 1 FMNMX3 : 1 LDS.128, %d blocks/SM", blocks_per_sm);
        report(nm, (double)blocks * 256 * iters * 80 * 2, best[2]);
    }
    {
        const int blocks = sms * 8;
        auto mix = [&](const char *name, auto kern, int n2, int n1, int nm) {
            double best = 1e30;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaEventRecord(e0)); kern<<<blocks, 256>>>(out, iters, in); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                best = std::min(best, time_ms(e0, e1));
            }
            const double fp32 = (double)blocks * 256 * iters * 16 * (2.0 * n2 + n1);
            printf("%-44s %8.3f ms  %9.1f G FP32 lane-instr/s  (+%d min/max per round)\n", name, best, fp32 / (best * 1e-3) / 1e9, 2 * nm);
        };
        mix("8 FFMA2", k_mix2<8, 0, 0>, 8, 0, 0);
        mix("8 FFMA2 + 4 FFMA", k_mix2<8, 4, 0>, 8, 4, 0);
        mix("8 FFMA2 + 8 FFMA", k_mix2<8, 8, 0>, 8, 8, 0);
        mix("4 FFMA2 + 8 FFMA", k_mix2<4, 8, 0>, 4, 8, 0);
        mix("16 FFMA", k_mix2<0, 16, 0>, 0, 16, 0);
        mix("8 FFMA2 + 2x2 FMNMX", k_mix2<8, 0, 2>, 8, 0, 2);
        mix("8 FFMA2 + 4x2 FMNMX", k_mix2<8, 0, 4>, 8, 0, 4);
        mix("16 FFMA + 4x2 FMNMX", k_mix2<0, 16, 4>, 0, 16, 4);
    }
    CK(cudaGetLastError());

    // filter loop over a resident tile of 2048 spheres (32 KB)
    const int n = 2048;
    std::vector<float> sph(4 * n), pkd(4 * n), rays;
    srand(7);
    auto rnd = [] { return (float)rand() / RAND_MAX; };
    for (int i = 0; i < n; i++) {
        sph[4 * i] = 200 * rnd() - 100; sph[4 * i + 1] = 200 * rnd() - 100; sph[4 * i + 2] = 5 + 200 * rnd();
        float r = 0.2f + 0.8f * rnd();
        sph[4 * i + 3] = r * r;
    }
    for (int p = 0; p < n / 2; p++) {
        const float *a = &sph[8 * p], *b = &sph[8 * p + 4];
        float *o = &pkd[8 * p];
        o[0] = a[0]; o[1] = b[0]; o[2] = a[1]; o[3] = b[1]; o[4] = a[2]; o[5] = b[2]; o[6] = a[3]; o[7] = b[3];
    }
    const int max_rays = sms * 16 * 128 * 4;
    rays.resize((size_t)max_rays * 6);
    for (int i = 0; i < max_rays; i++) {
        float d[3] = {rnd() - 0.5f, rnd() - 0.5f, 1.0f}, l = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        rays[6 * i] = rnd(); rays[6 * i + 1] = rnd(); rays[6 * i + 2] = -2;
        rays[6 * i + 3] = d[0] / l; rays[6 * i + 4] = d[1] / l; rays[6 * i + 5] = d[2] / l;
    }
    float4 *g_s, *g_p;
    float *g_r, *o_s, *o_p;
    CK(cudaMalloc(&g_s, n * 16)); CK(cudaMalloc(&g_p, n * 16));
    CK(cudaMalloc(&g_r, rays.size() * 4));
    CK(cudaMalloc(&o_s, (size_t)max_rays * 4)); CK(cudaMalloc(&o_p, (size_t)max_rays * 4));
    CK(cudaMemcpy(g_s, sph.data(), n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g_p, pkd.data(), n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(g_r, rays.data(), rays.size() * 4, cudaMemcpyHostToDevice));
    const int reps = 64;
    auto run = [&](const char *name, auto kern, const float4 *g, int nr, int blocks_per_sm, float *o) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, n * 16));
        const int blocks = sms * blocks_per_sm;
        double best = 1e30;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            kern<<<blocks, 128, n * 16>>>(g, n, reps, g_r, o);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            best = std::min(best, time_ms(e0, e1));
        }
        CK(cudaGetLastError());
        char nm[96];
        snprintf(nm, sizeof nm, "%s, %d blocks/SM", name, blocks_per_sm);
        report(nm, (double)blocks * 128 * nr * (double)n * reps * 10, best);
    };
    for (int bps : {4, 6, 8}) {
        run("scan scalar 1 ray/thread", k_scan_scalar<1>, g_s, 1, bps, o_s);
        run("scan scalar 2 rays/thread", k_scan_scalar<2>, g_s, 2, bps, o_s);
        run("scan scalar 4 rays/thread", k_scan_scalar<4>, g_s, 4, bps, o_s);
        run("scan packed 1 ray/thread", k_scan_packed<1>, g_p, 1, bps, o_p);
        run("scan packed 2 rays/thread", k_scan_packed<2>, g_p, 2, bps, o_p);
        run("scan packed 4 rays/thread", k_scan_packed<4>, g_p, 4, bps, o_p);
    }
    // the two layouts must agree bit for bit (same operations per sphere)
    {
        std::vector<float> a((size_t)sms * 4 * 128), b(a.size());
        k_scan_scalar<2><<<sms * 4, 128, n * 16>>>(g_s, n, 1, g_r, o_s);
        k_scan_packed<2><<<sms * 4, 128, n * 16>>>(g_p, n, 1, g_r, o_p);
        CK(cudaMemcpy(a.data(), o_s, a.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), o_p, b.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < a.size(); i++) bad += a[i] != b[i];
        printf("packed vs scalar results: %zu of %zu differ\n", bad, a.size());
    }
    return 0;
}
