// Test-only C shim around the host-side light-grid builder (eraytracer_b200/csrc/light_grid.cpp),
// so the CPU suite can check the one property that matters: a sphere a shadow ray can touch is
// always listed in the ray's direction cell.
#include "../../eraytracer_b200/csrc/light_grid.h"

using namespace ert;

extern "C" {

void *lg_build(const double *centers, const double *radii, const float *filter, long long n, const double *light, int res)
{
    LightGrid *g = new LightGrid();
    if (!build_light_grid(centers, radii, filter, n, light, res, *g)) { delete g; return nullptr; }
    return g;
}
void lg_free(void *p) { delete (LightGrid *)p; }
long long lg_cell(const double *d, int res) { return light_grid_cell(d, res); }
long long lg_cell_f32(const float *d, int res) { return light_grid_cell_f32(d, res); }
long long lg_n_entries(void *p) { return (long long)((LightGrid *)p)->entries.size(); }
long long lg_n_always(void *p) { return (long long)((LightGrid *)p)->always.size(); }
const unsigned int *lg_offsets(void *p) { return ((LightGrid *)p)->cell_off.data(); }
const void *lg_entries(void *p) { return ((LightGrid *)p)->entries.data(); }
const int *lg_always(void *p) { return ((LightGrid *)p)->always.data(); }

}
