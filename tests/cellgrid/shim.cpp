// Test-only C shim around the host-side cell-grid builder (eraytracer_b200/csrc/cell_grid.cpp) plus
// a restatement, in host C++ with the same FP32 operations (fmaf, correctly rounded 1/x), of the
// walk the device code makes (grid_start / grid_advance in ert_wavefront.cuh).  The CPU suite uses
// it to check the one property that matters: every sphere a ray touches is listed in a cell the
// walk visits no later than the ray reaches the sphere.
#include <cmath>
#include <limits>

#include "../../eraytracer_b200/csrc/cell_grid.h"

using namespace ert;

namespace {
constexpr float kU = 5.9604644775390625e-8f;
constexpr double kKD = 1.0000019073486328125;
constexpr float kKappa = 1.00006103515625f;
float clamp_dir(float d) { return std::fabs(d) < 1e-20f ? (d < 0.f ? -1e-20f : 1e-20f) : d; }
}  // namespace

extern "C" {

void *cg_build(const double *centers, const double *radii, const float *filter, long long n, float abs_max,
               double density)
{
    CellGrid *g = new CellGrid();
    build_cell_grid(centers, radii, filter, n, abs_max, density, *g);
    return g;
}
void cg_free(void *p) { delete (CellGrid *)p; }
int cg_enabled(void *p) { return ((CellGrid *)p)->enabled ? 1 : 0; }
void cg_geometry(void *p, int *res, float *lo, float *hi, float *cs_eps)
{
    const CellGrid &g = *(CellGrid *)p;
    for (int a = 0; a < 3; a++) { res[a] = g.res[a]; lo[a] = g.lo[a]; hi[a] = g.hi[a]; }
    cs_eps[0] = g.cs; cs_eps[1] = g.eps;
}
long long cg_n_cells(void *p) { return (long long)((CellGrid *)p)->cells.size(); }
long long cg_n_refs(void *p) { return (long long)((CellGrid *)p)->ref_sph.size(); }
long long cg_n_big(void *p) { return (long long)((CellGrid *)p)->big.size(); }
const unsigned int *cg_cells(void *p) { return ((CellGrid *)p)->cells.data(); }
const int *cg_ref_sph(void *p) { return ((CellGrid *)p)->ref_sph.data(); }
const int *cg_big(void *p) { return ((CellGrid *)p)->big.data(); }
// the layout the device walks (pack_cell_blocks): 128-byte blocks + overflow arrays
const void *cg_blocks(void *p) { return ((CellGrid *)p)->blocks.data(); }
long long cg_n_over(void *p) { return (long long)((CellGrid *)p)->over_sph.size(); }
const int *cg_over_sph(void *p) { return ((CellGrid *)p)->over_sph.data(); }
const float *cg_over_filter(void *p) { return ((CellGrid *)p)->over_filter.data(); }
const float *cg_ref_filter(void *p) { return ((CellGrid *)p)->ref_filter.data(); }

// The device walk for the ray (O, D), no incumbent.  Writes up to `cap` visited cell ids and the
// ray parameter (filter space: |d| = 1 + 2^-19) at which each was entered; returns the number of
// cells visited, or -1 when the ray's margin is too large for the grid (the device then walks the
// BVH), or -2 when `cap` was too small.
long long cg_walk(void *p, const double *O, const double *D, float abs_max, int *cells_out, float *t_enter_out,
                  long long cap)
{
    const CellGrid &g = *(CellGrid *)p;
    // make_sray
    const double a = D[0] * D[0] + D[1] * D[1] + D[2] * D[2];
    const double inv = (std::fabs(a - 1.0) <= 1e-9) ? (1.5 - 0.5 * a) : 1.0 / std::sqrt(a);
    const double k = inv * kKD;
    float o[3], d[3], iv[3], kl[3], kh[3];
    float eo = 0.f, oabs = 0.f;
    for (int x = 0; x < 3; x++) {
        o[x] = (float)O[x];
        d[x] = (float)(D[x] * k);
        eo = std::fmax(eo, std::fabs((float)(O[x] - (double)o[x])));
        oabs = std::fmax(oabs, std::fabs(o[x]));
    }
    eo = 2.0f * eo * 1.0001f;
    const float m = eo + 32.0f * kU * (oabs + abs_max);
    if (!(4.0f * m <= g.eps)) return -1;
    for (int x = 0; x < 3; x++) {
        iv[x] = 1.0f / clamp_dir(d[x]);
        kl[x] = -(o[x] + m) * iv[x];
        kh[x] = -(o[x] - m) * iv[x];
    }
    // slab_test against the grid bounds
    float tn = -std::numeric_limits<float>::infinity(), tf = std::numeric_limits<float>::infinity();
    for (int x = 0; x < 3; x++) {
        float ta = std::fmaf(g.lo[x], iv[x], kl[x]), tb = std::fmaf(g.hi[x], iv[x], kh[x]);
        tn = std::fmax(tn, std::fmin(ta, tb));
        tf = std::fmin(tf, std::fmax(ta, tb));
    }
    const float t0 = std::fmax(tn, 0.f);
    if (!(t0 <= tf * kKappa)) return 0;
    // grid_start
    const float inv_cs = 1.0f / g.cs;
    float A[3], B[3], nb[3], t[3];
    int step[3], idx[3];
    const int stride[3] = {1, g.res[0], g.res[0] * g.res[1]};
    long long id = 0;
    for (int x = 0; x < 3; x++) {
        A[x] = g.cs * iv[x];
        B[x] = std::fmaf(g.lo[x], iv[x], -(o[x] * iv[x]));
        float c = std::floor((std::fmaf(d[x], t0, o[x]) - g.lo[x]) * inv_cs);
        c = std::fmin(std::fmax(c, 0.f), (float)(g.res[x] - 1));
        const bool pos = iv[x] >= 0.f;
        nb[x] = c + (pos ? 1.f : 0.f);
        t[x] = std::fmaf(nb[x], A[x], B[x]);
        step[x] = pos ? 1 : -1;
        idx[x] = (int)c;
        id += (long long)stride[x] * idx[x];
    }
    long long n = 0;
    float t_in = t0;
    for (;;) {
        if (n >= cap) return -2;
        cells_out[n] = (int)id;
        t_enter_out[n] = t_in;
        n++;
        const float te = std::fmin(std::fmin(t[0], t[1]), t[2]);
        const int x = (t[0] == te) ? 0 : ((t[1] == te) ? 1 : 2);
        nb[x] += (float)step[x];
        t[x] = std::fmaf(nb[x], A[x], B[x]);
        id += (long long)step[x] * stride[x];
        if (nb[x] == (step[x] > 0 ? (float)(g.res[x] + 1) : -1.f)) break;
        t_in = te;
    }
    return n;
}

}
