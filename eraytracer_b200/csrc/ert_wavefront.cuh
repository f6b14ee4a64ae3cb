// Wavefront form of the eraytracer hot path for large sphere counts (ERT_ACCEL_BVH).
//
// The per-pixel recursion of raytracer.erl:186-252 is cut into queues that live in HBM, so
// that every traversal warp is full of live rays of ONE kind:
//
//   bounce b:  wf_trace_path    path queue  -> nearest hit (erl:300-346) -> hit queue (compacted)
//              wf_trace_shadow  hit queue x lights -> shadow_factor (erl:256-267) -> lit flags
//              wf_shade         hit queue + lit flags -> colour accumulation (erl:209-252)
//                                                     -> path queue of bounce b+1 (compacted)
//   end:       wf_finalize      colour -> framebuffer (quantisation of erl:678-680 fused)
//
// The arithmetic is the megakernel's (ert_device.cuh): FP64 in the literal operation order for
// everything that feeds a decision or a colour, FP32 only in conservative filters.  Both forms
// produce bit-identical frames (tests/test_gpu_parity.py).
#pragma once

#include <limits.h>

#include "ert_device.cuh"

#ifdef ERT_WF_DEBUG
#include <cstdio>
#define WF_ASSERT(cond, ...)                                            \
    do {                                                                \
        if (!(cond)) {                                                  \
            printf("WF_ASSERT %s:%d " #cond " : ", __FILE__, __LINE__); \
            printf(__VA_ARGS__);                                        \
            printf("\n");                                               \
            __trap();                                                   \
        }                                                               \
    } while (0)
#else
#define WF_ASSERT(cond, ...) do { } while (0)
#endif

namespace ert {

struct WfBuf {
    int n_pad;                  // pixels of this part padded to whole 8x4 tiles: tiles * 32
    int tiles_x;                // tiles per row of tiles
    double *C;                  // [3][n_pad] colour so far, by pixel
    double *W;                  // [n_pad]    product of (lights * reflectivity) of earlier bounces
    int *q_pid;                 // path queue: pixel of each ray
    double *q_ray;              // [6][n_pad] origin xyz, direction xyz
    int *h_pid;                 // hit queue: pixel,
    double *h_geo;              // [9][n_pad] hit location, normal, incoming direction
    int *h_obj, *h_order;       //            object code and list position
    unsigned char *lit;         // [n_lights][n_pad] shadow factor of (light, hit)
    unsigned int *ctr;          // [depth][kWfCtr] queue lengths
    // hits of bounces >= 1 are binned by location before their shadow rays are traced
    int *r_pid;                 // unsorted hit queue, same layout as h_*
    double *r_geo;
    int *r_obj, *r_order;
    unsigned int *r_key;        // cell of the hit location (Morton order)
    unsigned int *hist;         // [kSortCells] cell histogram -> offsets -> scatter cursors
    unsigned int *sums;         // [kSortBlocks] per-block totals of the histogram scan
};
enum WfCtr : int { WF_NHITS = 0, WF_NNEXT = 1, kWfCtr = 4 };
constexpr int kWfThreads = 256;
constexpr int kSortBits = 7;                                  // per axis
constexpr int kSortCells = 1 << (3 * kSortBits);              // 2 Mi cells over the sphere bounds
constexpr int kSortScanBlock = 4096;                          // cells scanned by one block
constexpr int kSortBlocks = kSortCells / kSortScanBlock;      // 512

// ------------------------------------------------------------------ slim ray for the traversal
// Same bounds as FRay (DESIGN.md "Filter bounds"); the slab constants are written so that the
// near/far choice needs no select: with m the absolute margin,
//     ta = lo*inv - (o+m)*inv,  tb = hi*inv - (o-m)*inv,  t_enter = min(ta,tb), t_exit = max(ta,tb)
// pushes the entered plane outward by m and the left plane outward by m for either sign of inv.
struct SRay {
    float ox, oy, oz;
    float dx, dy, dz;
    float ix, iy, iz;
    float klx, kly, klz;
    float khx, khy, khz;
    float theta, bcull, pad2, m4;
    double inv_sqrt_a, a;
};
#define ERT_KAPPA 1.00006103515625f   /* 1 + 2^-14 >= (1+2^-16)/(1-2^-16): slab slack as one factor */

__device__ __forceinline__ void make_sray(const DevScene &sc, d3 O, d3 D, SRay &f)
{
    f.a = D.x * D.x + D.y * D.y + D.z * D.z;
    double inv = 1.0 / sqrt(f.a);
    f.inv_sqrt_a = inv;
    f.ox = (float)O.x; f.oy = (float)O.y; f.oz = (float)O.z;
    double k = inv * ERT_KD;
    f.dx = (float)(D.x * k); f.dy = (float)(D.y * k); f.dz = (float)(D.z * k);
    float eo = 2.0f * fmaxf(fmaxf(fabsf((float)(O.x - (double)f.ox)), fabsf((float)(O.y - (double)f.oy))),
                            fabsf((float)(O.z - (double)f.oz)));
    eo *= 1.0001f;
    f.theta = 16.0f * sc.r_max * eo + 8.0f * eo * (eo / ERT_U);
    f.bcull = 2.0f * (5.0f * ERT_U * sc.r_max + 2.0f * eo + 2.0f * sc.eta_c_max);
    f.pad2 = 2.0f * (f.theta + sc.pad_c_max) + 1e-30f;
    float oabs = fmaxf(fmaxf(fabsf(f.ox), fabsf(f.oy)), fabsf(f.oz));
    float m = eo + 32.0f * ERT_U * (oabs + sc.abs_max);
    f.m4 = 4.0f * m;
    f.ix = 1.0f / clamp_dir(f.dx); f.iy = 1.0f / clamp_dir(f.dy); f.iz = 1.0f / clamp_dir(f.dz);
    f.klx = -(f.ox + m) * f.ix; f.kly = -(f.oy + m) * f.iy; f.klz = -(f.oz + m) * f.iz;
    f.khx = -(f.ox - m) * f.ix; f.khy = -(f.oy - m) * f.iy; f.khz = -(f.oz - m) * f.iz;
}

// 6 FFMA + 6 FMNMX + 2 FMNMX3 + FMNMX + FMUL + FMNMX + compare.  `cullk` already carries kappa.
__device__ __forceinline__ bool slab_test(const SRay &f, float lox, float hix, float loy, float hiy, float loz,
                                          float hiz, float cullk, float &tnear)
{
    float ax = __fmaf_rn(lox, f.ix, f.klx), bx = __fmaf_rn(hix, f.ix, f.khx);
    float ay = __fmaf_rn(loy, f.iy, f.kly), by = __fmaf_rn(hiy, f.iy, f.khy);
    float az = __fmaf_rn(loz, f.iz, f.klz), bz = __fmaf_rn(hiz, f.iz, f.khz);
    float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    tnear = fmaxf(tn, 0.f);
    return tnear <= fminf(tf * ERT_KAPPA, cullk);
}

template <class R>
__device__ __forceinline__ float cullk_from(const R &f, const Hit &best)
{
    return cull_from(f, best) * ERT_KAPPA;
}

// One leaf sphere: FP32 filter, then the literal FP64 test on survivors (rare).
template <bool COUNT>
__device__ __forceinline__ void leaf_sphere(const DevScene &sc, const SRay &f, d3 O, d3 D, float4 fs, int slot,
                                            int skip_obj, Hit &best, float &cullk, Tally<COUNT> &tl)
{
    float b, v;
    TALLY(filter);
    if (!filter_stage1(f, fs, b, v)) return;
    if (!filter_stage2(f, fs, b, v, cullk)) return;
    int sph = __ldg(sc.leaf_sph + slot);
    int code = obj_code(OBJ_SPHERE, sph);
    if (code == skip_obj) return;
    double t;
    TALLY(exact_sph);
    if (sphere_exact(O, D, f.a, sc.sph_exact[sph], t)) {
        int ord = sc.sph_order[sph];
        if (better(t, ord, best)) {
            best.t = t; best.order = ord; best.obj = code;
            cullk = cullk_from(f, best);
        }
    }
}

// While-while traversal: all lanes walk inner nodes until each holds a leaf (or is done), then
// the leaves are processed together.  ANY: stop at the first improvement of the incumbent.
template <bool ANY, bool COUNT>
__device__ __forceinline__ void traverse_bvh(const DevScene &sc, d3 O, d3 D, const SRay &f, Hit &best, int skip_obj,
                                             int seed_obj, Tally<COUNT> &tl)
{
    constexpr int SENT = INT_MIN;
    int stack[kBvhStack];
    float tstack[ANY ? 1 : kBvhStack];
    int sp = 0;
    stack[0] = SENT;
    // -inf: a triangle incumbent can have t < 0 (erl:402-455 has no t >= 0 test), so cullk may be negative
    if constexpr (!ANY) tstack[0] = __int_as_float(0xff800000);
    sp = 1;
    int node = 0;
    float cullk = cullk_from(f, best);
    for (;;) {
        while (node >= 0) {
            WF_ASSERT(node < sc.n_nodes, "node %d of %d sp %d", node, sc.n_nodes, sp);
            const float4 *np = reinterpret_cast<const float4 *>(sc.nodes + node);
            float4 a0 = __ldg(np), a1 = __ldg(np + 1), a2 = __ldg(np + 2);
            int2 ch = __ldg(reinterpret_cast<const int2 *>(np + 3));
            float tn0, tn1;
            if constexpr (COUNT) tl.box += 2;
            bool h0 = slab_test(f, a0.x, a0.y, a0.z, a0.w, a2.x, a2.y, cullk, tn0);
            bool h1 = slab_test(f, a1.x, a1.y, a1.z, a1.w, a2.z, a2.w, cullk, tn1);
            if (h0 && h1) {
                bool swap = tn1 < tn0;
                node = swap ? ch.y : ch.x;
                if (sp < kBvhStack) {
                    stack[sp] = swap ? ch.x : ch.y;
                    if constexpr (!ANY) tstack[sp] = swap ? tn0 : tn1;
                    sp++;
                }
            } else if (h0) {
                node = ch.x;
            } else if (h1) {
                node = ch.y;
            } else {
                if constexpr (ANY) {
                    node = stack[--sp];
                } else {
                    do { --sp; WF_ASSERT(sp >= 0, "sp %d cullk %g", sp, cullk); node = stack[sp]; } while (tstack[sp] > cullk);
                }
            }
            WF_ASSERT(sp >= 0 && sp <= kBvhStack, "sp %d", sp);
        }
        if (node == SENT) break;
        {
            int code = ~node;
            int first = code >> 3, cnt = (code & 7) + 1;
            WF_ASSERT(first >= 0 && first + cnt <= sc.n_spheres, "leaf first %d cnt %d node %d sp %d", first, cnt, node, sp);
#pragma unroll 1
            for (int k = 0; k < cnt; k++) {
                float4 fs = __ldg(sc.leaf_filter + first + k);
                leaf_sphere<COUNT>(sc, f, O, D, fs, first + k, skip_obj, best, cullk, tl);
            }
        }
        if constexpr (ANY) {
            if (best.obj != seed_obj) break;
            node = stack[--sp];
        } else {
            do { --sp; WF_ASSERT(sp >= 0, "sp %d cullk %g (leaf)", sp, cullk); node = stack[sp]; } while (tstack[sp] > cullk);
        }
    }
}

__device__ void trace_ray_wavefront(const DevScene &sc, d3 O, d3 D, Hit &best)
{
    SRay f;
    Tally<false> tl;
    make_sray(sc, O, D, f);
    traverse_bvh<false, false>(sc, O, D, f, best, -1, -1, tl);
}

// ------------------------------------------------------------------ pixel <-> queue index
// Pixel i of the part: 8x4 tile (i >> 5), lane (i & 31) inside it, tiles row-major.
__device__ __forceinline__ void pixel_of_index(const FrameParams &fp, const WfBuf &wf, int i, int &X, int &local_y,
                                               int &Y, bool &inside)
{
    int tile = i >> 5, lane = i & 31;
    int ty = tile / wf.tiles_x, tx = tile - ty * wf.tiles_x;
    X = tx * 8 + (lane & 7);
    local_y = ty * 4 + (lane >> 3);
    if (fp.n_parts > 1 && fp.band_rows > 0) {
        int j = local_y / fp.band_rows;
        Y = (fp.part + j * fp.n_parts) * fp.band_rows + (local_y - j * fp.band_rows);
    } else {
        Y = local_y;
    }
    inside = X < fp.width && local_y < fp.local_rows && Y < fp.height;
}

template <bool COUNT>
__device__ __forceinline__ void wf_flush(const FrameParams &fp, unsigned int rays, const Tally<COUNT> &tl)
{
    flush_counters<COUNT>(fp, (int)rays, tl);
}

// ------------------------------------------------------------------ binning hits by location
// Rays that leave a curved surface scatter: the hits of bounce >= 1 arrive in an order that
// has little to do with where they are, and a warp of their shadow rays walks 32 unrelated
// parts of the tree.  A counting sort by the Morton cell of the hit location (a uniform grid
// over the sphere bounds) restores the coherence bounce 0 has for free.  The order of a queue
// never changes a pixel (every record carries its pixel), so this is purely a schedule.
__device__ __forceinline__ unsigned int part1by2(unsigned int x)
{
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x30000ffu;
    x = (x | (x << 8)) & 0x300f00fu;
    x = (x | (x << 4)) & 0x30c30c3u;
    x = (x | (x << 2)) & 0x9249249u;
    return x;
}
__device__ __forceinline__ unsigned int sort_cell(const DevScene &sc, d3 P)
{
    const float top = (float)((1 << kSortBits) - 1);
    float fx = fminf(fmaxf(((float)P.x - sc.grid_lo[0]) * sc.grid_scale[0], 0.f), top);
    float fy = fminf(fmaxf(((float)P.y - sc.grid_lo[1]) * sc.grid_scale[1], 0.f), top);
    float fz = fminf(fmaxf(((float)P.z - sc.grid_lo[2]) * sc.grid_scale[2], 0.f), top);
    return part1by2((unsigned int)fx) | (part1by2((unsigned int)fy) << 1) | (part1by2((unsigned int)fz) << 2);
}

// exclusive scan of the histogram, phase A: each block scans kSortScanBlock cells in place
__global__ void __launch_bounds__(1024) wf_bin_scan_a(const __grid_constant__ WfBuf wf)
{
    __shared__ unsigned int warp_sums[32];
    unsigned int *h = wf.hist + (size_t)blockIdx.x * kSortScanBlock + threadIdx.x * 4;
    uint4 v = *reinterpret_cast<uint4 *>(h);
    unsigned int t = v.x + v.y + v.z + v.w;
    unsigned int incl = t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = warp_sums[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        warp_sums[lane] = wi - w;
        if (lane == 31) wf.sums[blockIdx.x] = wi;
    }
    __syncthreads();
    unsigned int base = warp_sums[warp] + incl - t;
    uint4 o;
    o.x = base; o.y = base + v.x; o.z = o.y + v.y; o.w = o.z + v.z;
    *reinterpret_cast<uint4 *>(h) = o;
}
// phase B: exclusive scan of the kSortBlocks block totals (one block)
__global__ void __launch_bounds__(kSortBlocks) wf_bin_scan_b(const __grid_constant__ WfBuf wf)
{
    __shared__ unsigned int warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int t = wf.sums[threadIdx.x], incl = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = lane < kSortBlocks / 32 ? warp_sums[lane] : 0u, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        warp_sums[lane] = wi - w;
    }
    __syncthreads();
    wf.sums[threadIdx.x] = warp_sums[warp] + incl - t;
}
// scatter: every unsorted hit record moves to its cell's range of the sorted hit queue
__global__ void __launch_bounds__(kWfThreads) wf_bin_scatter(const __grid_constant__ WfBuf wf, int bounce)
{
    const unsigned int n_hits = wf.ctr[bounce * kWfCtr + WF_NHITS];
    const size_t np = (size_t)wf.n_pad;
    for (size_t h = (size_t)blockIdx.x * kWfThreads + threadIdx.x; h < n_hits; h += (size_t)gridDim.x * kWfThreads) {
        unsigned int key = wf.r_key[h];
        size_t s = (size_t)atomicAdd(wf.hist + key, 1u) + wf.sums[key / kSortScanBlock];
        wf.h_pid[s] = wf.r_pid[h];
        wf.h_obj[s] = wf.r_obj[h];
        wf.h_order[s] = wf.r_order[h];
#pragma unroll
        for (int k = 0; k < 9; k++) wf.h_geo[k * np + s] = wf.r_geo[k * np + h];
    }
}

// ------------------------------------------------------------------ kernels

// Path rays of one bounce.  FIRST: rays are generated from the pixel index (erl:486-511).
template <bool FIRST, bool SORT, bool COUNT>
__global__ void __launch_bounds__(kWfThreads)
wf_trace_path(const __grid_constant__ DevScene sc, const __grid_constant__ FrameParams fp,
              const __grid_constant__ WfBuf wf, int bounce)
{
    unsigned int *ctr = wf.ctr + bounce * kWfCtr;
    const unsigned int n = FIRST ? (unsigned int)wf.n_pad : wf.ctr[(bounce - 1) * kWfCtr + WF_NNEXT];
    const int lane = threadIdx.x & 31;
    const unsigned int warp = (blockIdx.x * kWfThreads + threadIdx.x) >> 5;
    const unsigned int n_warps = (gridDim.x * kWfThreads) >> 5;
    const size_t np = (size_t)wf.n_pad;
    Tally<COUNT> tl;
    unsigned int rays = 0;
    for (unsigned long long base = (unsigned long long)warp * 32; base < n; base += (unsigned long long)n_warps * 32) {
        unsigned int i = (unsigned int)base + lane;
        bool valid = i < n;
        d3 O = mk(0, 0, 0), D = mk(0, 0, 1);
        int pid = 0;
        if constexpr (FIRST) {
            int X, ly, Y;
            pixel_of_index(fp, wf, (int)i, X, ly, Y, valid);
            pid = (int)i;
            if (valid) primary_ray(fp, X, Y, O, D);
        } else if (valid) {
            WF_ASSERT(i < (unsigned)wf.n_pad, "queue index %u of %d", i, wf.n_pad);
            pid = wf.q_pid[i];
            O = mk(wf.q_ray[i], wf.q_ray[np + i], wf.q_ray[2 * np + i]);
            D = mk(wf.q_ray[3 * np + i], wf.q_ray[4 * np + i], wf.q_ray[5 * np + i]);
        }
        Hit best;
        best.obj = -1; best.t = 0.0; best.order = 0x7fffffff;
        if (valid) {
            rays++;
            scan_others<COUNT>(sc, O, D, best, -1, tl);
            if (sc.n_spheres > 0) {
                SRay f;
                make_sray(sc, O, D, f);
                traverse_bvh<false, COUNT>(sc, O, D, f, best, -1, -1, tl);
            }
        }
        bool hit = valid && best.obj >= 0;
        unsigned int m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            int leader = __ffs(m) - 1;
            unsigned int slot0 = 0;
            if (lane == leader) slot0 = atomicAdd(ctr + WF_NHITS, (unsigned int)__popc(m));
            slot0 = __shfl_sync(0xffffffffu, slot0, leader);
            if (hit) {
                size_t s = slot0 + __popc(m & ((1u << lane) - 1u));
                d3 P = vadd(O, vscale(D, best.t));            // erl:384-387 / 443-447 / 471-475
                d3 N = hit_normal(sc, best.obj, P);
                (SORT ? wf.r_pid : wf.h_pid)[s] = pid;
                (SORT ? wf.r_obj : wf.h_obj)[s] = best.obj;
                (SORT ? wf.r_order : wf.h_order)[s] = best.order;
                if constexpr (SORT) {
                    unsigned int key = sort_cell(sc, P);
                    wf.r_key[s] = key;
                    atomicAdd(wf.hist + key, 1u);
                }
                double *g = (SORT ? wf.r_geo : wf.h_geo) + s;
                g[0] = P.x; g[np] = P.y; g[2 * np] = P.z;
                g[3 * np] = N.x; g[4 * np] = N.y; g[5 * np] = N.z;
                g[6 * np] = D.x; g[7 * np] = D.y; g[8 * np] = D.z;
            }
        }
    }
    wf_flush<COUNT>(fp, rays, tl);
}

// shadow_factor/4 (erl:256-267) of every (light, hit) pair of one bounce, light-major so that a
// warp holds rays that leave one light towards neighbouring hit locations.
template <bool COUNT>
__global__ void __launch_bounds__(kWfThreads)
wf_trace_shadow(const __grid_constant__ DevScene sc, const __grid_constant__ FrameParams fp,
                const __grid_constant__ WfBuf wf, int bounce)
{
    const unsigned int n_hits = wf.ctr[bounce * kWfCtr + WF_NHITS];
    const unsigned long long total = (unsigned long long)n_hits * (unsigned long long)sc.n_lights;
    const int lane = threadIdx.x & 31;
    const unsigned int warp = (blockIdx.x * kWfThreads + threadIdx.x) >> 5;
    const unsigned int n_warps = (gridDim.x * kWfThreads) >> 5;
    const size_t np = (size_t)wf.n_pad;
    Tally<COUNT> tl;
    unsigned int rays = 0;
    for (unsigned long long base = (unsigned long long)warp * 32; base < total; base += (unsigned long long)n_warps * 32) {
        unsigned long long j = base + lane;
        if (j >= total) continue;
        unsigned int l = (unsigned int)(base / n_hits);
        unsigned long long h64 = base - (unsigned long long)l * n_hits + lane;
        while (h64 >= n_hits) { h64 -= n_hits; l++; }
        size_t h = (size_t)h64;
        rays++;
        d3 P = mk(wf.h_geo[h], wf.h_geo[np + h], wf.h_geo[2 * np + h]);
        int obj = wf.h_obj[h], order = wf.h_order[h];
        const double *lt = sc.lights + 9 * (size_t)l;
        d3 O = mk(lt[3], lt[4], lt[5]);
        d3 D = vnormalize(vsub(P, O));                        // erl:257-260
        double a = D.x * D.x + D.y * D.y + D.z * D.z;
        unsigned char lit = 0;
        double t;
        // "nearest == Object" (erl:263) <=> Object is hit and nothing beats its (t, order)
        if (object_exact(sc, obj, O, D, a, t)) {
            Hit best;
            best.t = t; best.order = order; best.obj = obj;
            scan_others<COUNT>(sc, O, D, best, obj, tl);
            if (best.obj == obj && sc.n_spheres > 0) {
                SRay f;
                make_sray(sc, O, D, f);
                traverse_bvh<true, COUNT>(sc, O, D, f, best, obj, obj, tl);
            }
            lit = best.obj == obj;
        }
        wf.lit[(size_t)l * np + h] = lit;
    }
    wf_flush<COUNT>(fp, rays, tl);
}

// Folds the lights of every hit of one bounce (erl:209-252 in forward form, see pix_consume)
// and emits the reflection rays of the next bounce.
__global__ void __launch_bounds__(kWfThreads)
wf_shade(const __grid_constant__ DevScene sc, const __grid_constant__ FrameParams fp,
         const __grid_constant__ WfBuf wf, int bounce)
{
    unsigned int *ctr = wf.ctr + bounce * kWfCtr;
    const unsigned int n_hits = ctr[WF_NHITS];
    const int lane = threadIdx.x & 31;
    const unsigned int warp = (blockIdx.x * kWfThreads + threadIdx.x) >> 5;
    const unsigned int n_warps = (gridDim.x * kWfThreads) >> 5;
    const size_t np = (size_t)wf.n_pad;
    const int L = sc.n_lights;
    for (unsigned long long base = (unsigned long long)warp * 32; base < n_hits; base += (unsigned long long)n_warps * 32) {
        size_t h = (size_t)base + lane;
        bool valid = h < n_hits;
        bool cont = false;
        int pid = 0;
        d3 P = mk(0, 0, 0), N = P, D = P;
        if (valid) {
            pid = wf.h_pid[h];
            const double *g = wf.h_geo + h;
            P = mk(g[0], g[np], g[2 * np]);
            N = mk(g[3 * np], g[4 * np], g[5 * np]);
            D = mk(g[6 * np], g[7 * np], g[8 * np]);
            const double *mat = material_ptr(sc, wf.h_obj[h]);
            d3 S = mk(0.0, 0.0, 0.0);
            for (int l = 0; l < L; l++) {
                if (wf.lit[(size_t)l * np + h]) S = vadd(S, light_term(sc.lights + 9 * (size_t)l, mat, P, N, D));
            }
            d3 Cold = mk(0.0, 0.0, 0.0);
            double Wold = 1.0;
            if (bounce > 0) {
                Cold = mk(wf.C[pid], wf.C[np + pid], wf.C[2 * np + pid]);
                Wold = wf.W[pid];
            }
            d3 Cnew = vadd(Cold, vscale(S, Wold));
            wf.C[pid] = Cnew.x; wf.C[np + pid] = Cnew.y; wf.C[2 * np + pid] = Cnew.z;
            double refl = mat[5];
            cont = !(bounce + 1 >= fp.depth || refl == 0.0 || Wold == 0.0);
            if (cont) wf.W[pid] = Wold * ((double)L * refl);
        }
        unsigned int m = __ballot_sync(0xffffffffu, cont);
        if (m) {
            int leader = __ffs(m) - 1;
            unsigned int slot0 = 0;
            if (lane == leader) slot0 = atomicAdd(ctr + WF_NNEXT, (unsigned int)__popc(m));
            slot0 = __shfl_sync(0xffffffffu, slot0, leader);
            if (cont) {
                size_t s = slot0 + __popc(m & ((1u << lane) - 1u));
                d3 nd = vbounce(D, N);                            // erl:219-221
                wf.q_pid[s] = pid;
                double *q = wf.q_ray + s;
                q[0] = P.x; q[np] = P.y; q[2 * np] = P.z;
                q[3 * np] = nd.x; q[4 * np] = nd.y; q[5 * np] = nd.z;
            }
        }
    }
}

__global__ void __launch_bounds__(kWfThreads)
wf_finalize(const __grid_constant__ FrameParams fp, const __grid_constant__ WfBuf wf)
{
    const size_t np = (size_t)wf.n_pad;
    for (size_t i = (size_t)blockIdx.x * kWfThreads + threadIdx.x; i < np; i += (size_t)gridDim.x * kWfThreads) {
        int X, ly, Y;
        bool inside;
        pixel_of_index(fp, wf, (int)i, X, ly, Y, inside);
        if (!inside) continue;
        Pix p;
        p.C = mk(wf.C[i], wf.C[np + i], wf.C[2 * np + i]);
        pix_store(p, fp, X, ly);
    }
}

}  // namespace ert
