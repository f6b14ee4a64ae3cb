// Builder of the uniform cell grid over the spheres.  See cell_grid.h.
#include "cell_grid.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace ert {
namespace {

float f_down(double x)
{
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -INFINITY);
    return f;
}
float f_up(double x)
{
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    return f;
}

}  // namespace

void cell_grid_range(const CellGrid &g, int axis, double a, double b, int &i0, int &i1)
{
    const double lo = (double)g.lo[axis], cs = (double)g.cs;
    double f0 = std::floor((a - (double)g.eps - lo) / cs), f1 = std::floor((b + (double)g.eps - lo) / cs);
    const double top = (double)(g.res[axis] - 1);
    i0 = (int)std::min(std::max(f0, 0.0), top);
    i1 = (int)std::min(std::max(f1, 0.0), top);
}

void build_cell_grid(const double *centers, const double *radii, const float *filter, int64_t n, float abs_max,
                     double density, CellGrid &out)
{
    out = CellGrid();
    if (n < kCellGridMinSpheres || !(density > 0) || !std::isfinite(abs_max) || abs_max > 1e6f) return;
    const double u = 5.9604644775390625e-8;
    // The walk follows the FP32 ray within m/2 of the true one (m: the per-ray margin of make_sray,
    // >= 32u(|o|+abs_max)); boxes are inflated by eps and a ray may use the grid iff 4m <= eps.
    const double eps = 1024.0 * u * std::max((double)abs_max, 1.0);
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int64_t k = 0; k < n; k++) {
        const double r = std::fabs(radii[k]);
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], centers[k * 3 + a] - r);
            hi[a] = std::max(hi[a], centers[k * 3 + a] + r);
        }
    }
    double ext[3], vol = 1.0;
    for (int a = 0; a < 3; a++) {
        out.lo[a] = f_down(lo[a] - 2.0 * eps);
        out.hi[a] = f_up(hi[a] + 2.0 * eps);
        ext[a] = (double)out.hi[a] - (double)out.lo[a];
        if (!(ext[a] > 0) || !std::isfinite(ext[a])) return;
        vol *= ext[a];
    }
    // cubic cells: edge from the wanted cell count, then never so small that an axis needs more
    // than kCellGridMaxRes cells
    double cs = std::cbrt(vol / (density * (double)n));
    for (int a = 0; a < 3; a++) cs = std::max(cs, ext[a] / (double)(kCellGridMaxRes - 1));
    if (!(cs > 64.0 * eps) || !std::isfinite(cs)) return;
    out.cs = f_up(cs);
    out.inv_cs = (float)(1.0 / (double)out.cs);
    out.eps = (float)eps;
    size_t n_cells = 1;
    for (int a = 0; a < 3; a++) {
        int r = (int)std::ceil(ext[a] / (double)out.cs) + 1;
        out.res[a] = std::min(std::max(r, 1), kCellGridMaxRes);
        if ((double)out.res[a] * (double)out.cs < ext[a]) return;
        n_cells *= (size_t)out.res[a];
    }
    if (n_cells > ((size_t)1 << 23)) return;           // 256 bytes per cell on the device

    // pass 1: counts
    std::vector<uint32_t> count(n_cells, 0);
    std::vector<uint8_t> is_big((size_t)n, 0);
    uint64_t refs = 0;
    auto range = [&](int64_t k, int (&i0)[3], int (&i1)[3]) {
        const double r = std::fabs(radii[k]);
        for (int a = 0; a < 3; a++) cell_grid_range(out, a, centers[k * 3 + a] - r, centers[k * 3 + a] + r, i0[a], i1[a]);
    };
    const size_t sx = 1, sy = (size_t)out.res[0], sz = (size_t)out.res[0] * (size_t)out.res[1];
    // A cell inside the box of a sphere may still be out of the ball's reach (the corners of the box):
    // the ball inflated by eps must come within the cell, also inflated by eps for the float planes.
    auto reaches = [&](int64_t k, int x, int y, int z) {
        const double r = std::fabs(radii[k]) + 2.0 * eps;
        const int c[3] = {x, y, z};
        double d2 = 0;
        for (int a = 0; a < 3; a++) {
            const double lo_a = (double)out.lo[a] + (double)c[a] * (double)out.cs, hi_a = lo_a + (double)out.cs;
            const double p = centers[k * 3 + a];
            const double d = p < lo_a ? lo_a - p : (p > hi_a ? p - hi_a : 0.0);
            d2 += d * d;
        }
        return d2 <= r * r * (1.0 + 1e-9);
    };
    for (int64_t k = 0; k < n; k++) {
        int i0[3], i1[3];
        range(k, i0, i1);
        uint64_t cells = (uint64_t)(i1[0] - i0[0] + 1) * (uint64_t)(i1[1] - i0[1] + 1) * (uint64_t)(i1[2] - i0[2] + 1);
        if (cells > (uint64_t)kCellGridBigCells) {
            is_big[(size_t)k] = 1;
            out.big.push_back((int32_t)k);
            if ((int)out.big.size() > kCellGridMaxBig) { out = CellGrid(); return; }
            continue;
        }
        for (int z = i0[2]; z <= i1[2]; z++)
            for (int y = i0[1]; y <= i1[1]; y++)
                for (int x = i0[0]; x <= i1[0]; x++)
                    if (reaches(k, x, y, z)) { count[x * sx + y * sy + z * sz]++; refs++; }
        if (refs >= kCellGridMaxRefs) { out = CellGrid(); return; }
    }
    // pass 2: offsets, packed cell words
    out.cells.resize(n_cells);
    std::vector<uint32_t> cursor(n_cells);
    uint32_t run = 0;
    for (size_t c = 0; c < n_cells; c++) {
        if (count[c] > (uint32_t)kCellGridMaxCount) { out = CellGrid(); return; }
        out.cells[c] = (run << 7) | count[c];
        cursor[c] = run;
        // every cell's list is stored in whole groups of kCellGridPad entries: the filter loop of the walk reads
        // a group at a time without looking at the count (the padding is a sphere that never passes)
        run += (count[c] + (uint32_t)kCellGridPad - 1) / (uint32_t)kCellGridPad * (uint32_t)kCellGridPad;
        if (run >= kCellGridMaxRefs) { out = CellGrid(); return; }
    }
    // pass 3: fill, spheres in list order inside a cell
    run += (uint32_t)kCellGridPad;                   // the loop may fetch one group past the last list
    out.ref_filter.assign((size_t)run * 4, 0.f);
    for (size_t k = 0; k < (size_t)run; k++) out.ref_filter[k * 4 + 3] = -3.0e38f;
    out.ref_sph.assign((size_t)run, -1);
    for (int64_t k = 0; k < n; k++) {
        if (is_big[(size_t)k]) continue;
        int i0[3], i1[3];
        range(k, i0, i1);
        for (int z = i0[2]; z <= i1[2]; z++)
            for (int y = i0[1]; y <= i1[1]; y++)
                for (int x = i0[0]; x <= i1[0]; x++) {
                    if (!reaches(k, x, y, z)) continue;
                    uint32_t slot = cursor[x * sx + y * sy + z * sz]++;
                    memcpy(&out.ref_filter[(size_t)slot * 4], &filter[(size_t)k * 4], 16);
                    out.ref_sph[slot] = (int32_t)k;
                }
    }
    out.enabled = true;
    pack_cell_blocks(out);
}

void pack_cell_blocks(CellGrid &g)
{
    const size_t n_cells = g.cells.size();
    constexpr uint32_t kIn = (uint32_t)(kCellGridInline * kCellGridBlocks);
    CellBlock empty;
    for (int k = 0; k < kCellGridInline; k++) {
        empty.f[k][0] = empty.f[k][1] = empty.f[k][2] = 0.f;
        empty.f[k][3] = -3.0e38f;
        empty.sph[k] = -1;
    }
    empty.cnt = 0; empty.more = 0;
    g.blocks.assign(n_cells * kCellGridBlocks, empty);
    auto padded = [](uint32_t cnt) { return cnt > kIn ? ((size_t)(cnt - kIn) + kCellGridPad - 1) / kCellGridPad * kCellGridPad : 0; };
    size_t n_over = 0;
    for (size_t c = 0; c < n_cells; c++) n_over += padded(g.cells[c] & 127u);
    n_over += kCellGridPad;                          // the walk's loop may fetch one group past the last list
    g.over_filter.assign(n_over * 4, 0.f);
    for (size_t k = 0; k < n_over; k++) g.over_filter[k * 4 + 3] = -3.0e38f;
    g.over_sph.assign(n_over, -1);
    size_t run = 0;
    for (size_t c = 0; c < n_cells; c++) {
        const uint32_t cnt = g.cells[c] & 127u, first = g.cells[c] >> 7;
        CellBlock *b = &g.blocks[c * kCellGridBlocks];
        b[0].cnt = cnt;
        b[0].more = (uint32_t)run;
        for (uint32_t k = 0; k < cnt; k++) {
            const size_t slot = (size_t)first + k;
            if (k < kIn) {
                CellBlock &bb = b[k / kCellGridInline];
                memcpy(bb.f[k % kCellGridInline], &g.ref_filter[slot * 4], 16);
                bb.sph[k % kCellGridInline] = g.ref_sph[slot];
            } else {
                const size_t o = run + (k - kIn);
                memcpy(&g.over_filter[o * 4], &g.ref_filter[slot * 4], 16);
                g.over_sph[o] = g.ref_sph[slot];
            }
        }
        run += padded(cnt);
    }
}

}  // namespace ert
