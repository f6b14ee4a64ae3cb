// Host-side sphere BVH builder (binned SAH, binary, two child boxes per node).
// The reference has no acceleration structure: nearest_object_intersecting_ray
// (raytracer.erl:300-346) is a linear scan.  The BVH only prunes; every sphere
// it lets through is still decided by the literal FP64 test, so the result is
// the linear scan's.
#pragma once
#include <cstdint>
#include <vector>

namespace ert {

// 64-byte node, children boxes stored in the parent so one node fetch decides both.
// child >= 0: inner node index.  child < 0: leaf, ~child = (first << 3) | (count - 1).
// An unused child has an inverted box (lo > hi) and can never be entered.
struct alignas(16) BvhNode {
    float c0x[2], c0y[2];   // child 0: lo.x hi.x lo.y hi.y
    float c1x[2], c1y[2];   // child 1
    float c0z[2], c1z[2];   // lo.z hi.z of child 0, then of child 1
    int32_t child[2];
    int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

constexpr int kBvhLeafMax = 4;      // default spheres per leaf (<= 8 by the encoding)
constexpr float kBvhTravCost = 0.f; // default SAH traversal cost (see build_sphere_bvh)
constexpr int kSahMaxDepth = 30;    // below this depth splits fall back to the median (log2 n more levels)
constexpr int kBvhStack = 64;       // traversal stack entries (depth <= 30 + log2(n) + 1)

struct Bvh {
    std::vector<BvhNode> nodes;     // nodes[0] is the root (always an inner node)
    std::vector<int32_t> leaf_prim; // leaf-ordered sphere indices
    int depth = 0;
};

// centers: n*3 doubles, radii: n doubles.  Boxes are [c-r, c+r] rounded outward to float.
// leaf_max: largest leaf (1..8).  trav_cost: cost of visiting one inner node in units of one
// ray/sphere filter test; a range of <= leaf_max spheres stays a leaf when the SAH says a
// split would not pay (trav_cost <= 0: always split down to leaf_max, the round-1 behaviour).
void build_sphere_bvh(const double *centers, const double *radii, int64_t n, Bvh &out, int leaf_max = kBvhLeafMax,
                      float trav_cost = 0.f);

}  // namespace ert
