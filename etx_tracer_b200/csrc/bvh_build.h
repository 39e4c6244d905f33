// bvh_build.h — host BVH container + builder entry point.
#pragma once
#include <vector>

#include "bvh.h"

namespace etxb {

struct Bvh {
  std::vector<BvhNode> nodes;       // nodes[0] is the root and always an inner node
  std::vector<F4> tri_pos;          // 3 entries per leaf slot
  std::vector<uint32_t> tri_index;  // leaf slot -> original triangle index
  uint32_t max_depth = 0;           // levels of inner nodes; build_bvh keeps it below kBvhStackSize (median splits once SAH would exceed it)
};

// positions: first 3 floats of each vertex record; indices: first 3 uint32 of each triangle record
// (strides in bytes — the reference's Vertex is 56 B, Triangle 32 B: sources/etx/render/shared/math.hxx:599,607)
struct WideBvh {
  std::vector<WideNode> nodes;  // nodes[0] is the root; the first 512 are the top levels, breadth-first
  uint32_t max_stack = 0;       // deepest the traversal stack can get (sum over a root-to-leaf path of children - 1)
};
// collapses the BVH2 into the 4-wide quantised form (bvh.h WideNode); leaves and tri_pos are shared with `bvh`
void build_wide_bvh(const Bvh& bvh, WideBvh& out);

void build_bvh(const float* positions, uint32_t position_stride_bytes, const uint32_t* indices, uint32_t index_stride_bytes, uint32_t tri_count, Bvh& out);

}  // namespace etxb
