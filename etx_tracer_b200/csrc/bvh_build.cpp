// bvh_build.cpp — host-side binned-SAH BVH2 builder (replaces rtcCommitScene, sources/etx/rt/rt.cxx:66-88).
// Deterministic: same input -> same nodes, so the oracle and the CUDA module traverse identical trees.
#include "bvh_build.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace etxb {

namespace {

struct Box {
  float lo[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float hi[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  void grow(const float* p) {
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(lo[k], p[k]);
      hi[k] = std::max(hi[k], p[k]);
    }
  }
  void grow(const Box& b) {
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(lo[k], b.lo[k]);
      hi[k] = std::max(hi[k], b.hi[k]);
    }
  }
  float half_area() const {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0.0f) return 0.0f;
    return dx * dy + dy * dz + dz * dx;
  }
};

struct Builder {
  const float* positions;
  uint32_t position_stride;  // in floats
  const uint32_t* indices;
  uint32_t index_stride;  // in uint32
  uint32_t tri_count;

  std::vector<Box> tri_box;
  std::vector<float> centroid;  // 3 * tri_count
  std::vector<uint32_t> order;

  std::vector<BvhNode> nodes;
  std::vector<uint32_t> leaf_tris;  // slot -> original triangle index
  uint32_t max_depth = 0;           // deepest inner node (root = 1): what the traversal stack must hold

  static constexpr int kBins = 16;

  const float* vertex(uint32_t tri, int k) const {
    return positions + size_t(indices[size_t(tri) * index_stride + k]) * position_stride;
  }

  int32_t make_leaf(uint32_t begin, uint32_t end) {
    uint32_t first = uint32_t(leaf_tris.size());
    uint32_t count = end - begin;
    for (uint32_t i = begin; i < end; ++i) leaf_tris.push_back(order[i]);
    uint32_t ref = (first << 2) | (count - 1u);
    return int32_t(~ref);
  }

  // returns child reference; writes bounds of the subtree into `bounds`
  int32_t build(uint32_t begin, uint32_t end, Box& bounds, bool force_inner, uint32_t depth = 1) {
    max_depth = std::max(max_depth, depth);
    bounds = Box();
    Box cbox;
    for (uint32_t i = begin; i < end; ++i) {
      bounds.grow(tri_box[order[i]]);
      cbox.grow(&centroid[size_t(order[i]) * 3]);
    }
    uint32_t count = end - begin;
    if ((count <= uint32_t(kBvhMaxLeafTris)) && !force_inner) {
      if (count <= 2u) return make_leaf(begin, end);
    }

    uint32_t mid = begin;
    bool split_found = false;
    // The traversal keeps at most one stack entry per level (kBvhStackSize).  SAH splits of clustered input can peel off a few triangles
    // per level without bound; once the levels left would not fit a median-split subtree of this range any more, the range is halved in
    // its current order instead (never reached by the scenes of the BASELINE configs: their trees are 30-40 levels deep).
    uint32_t levels_needed = 1;
    while ((1ull << levels_needed) < count) levels_needed++;
    const bool depth_limited = depth + levels_needed + 2u >= uint32_t(kBvhStackSize);
    if ((count > 1u) && depth_limited) {
      if ((count > uint32_t(kBvhMaxLeafTris)) || force_inner) {
        mid = begin + count / 2u;
        split_found = true;
      }
    } else if (count > 1u) {
      float best_cost = std::numeric_limits<float>::max();
      int best_axis = -1, best_bin = -1;
      for (int axis = 0; axis < 3; ++axis) {
        float cmin = cbox.lo[axis], cmax = cbox.hi[axis];
        if (!(cmax > cmin)) continue;
        Box bin_box[kBins];
        uint32_t bin_count[kBins] = {};
        float scale = float(kBins) / (cmax - cmin);
        for (uint32_t i = begin; i < end; ++i) {
          uint32_t t = order[i];
          int b = std::min(kBins - 1, std::max(0, int((centroid[size_t(t) * 3 + axis] - cmin) * scale)));
          bin_box[b].grow(tri_box[t]);
          bin_count[b]++;
        }
        float right_area[kBins];
        uint32_t right_count[kBins];
        Box acc;
        uint32_t cnt = 0;
        for (int b = kBins - 1; b > 0; --b) {
          acc.grow(bin_box[b]);
          cnt += bin_count[b];
          right_area[b] = acc.half_area();
          right_count[b] = cnt;
        }
        acc = Box();
        cnt = 0;
        for (int b = 0; b < kBins - 1; ++b) {
          acc.grow(bin_box[b]);
          cnt += bin_count[b];
          if (cnt == 0 || right_count[b + 1] == 0) continue;
          float cost = acc.half_area() * float(cnt) + right_area[b + 1] * float(right_count[b + 1]);
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = axis;
            best_bin = b;
          }
        }
      }
      if (best_axis >= 0) {
        float leaf_cost = bounds.half_area() * float(count);
        bool must_split = (count > uint32_t(kBvhMaxLeafTris)) || force_inner;
        if (must_split || (best_cost + bounds.half_area() * 1.0f < leaf_cost)) {
          float cmin = cbox.lo[best_axis], cmax = cbox.hi[best_axis];
          float scale = float(kBins) / (cmax - cmin);
          auto it = std::stable_partition(order.begin() + begin, order.begin() + end, [&](uint32_t t) {
            int b = std::min(kBins - 1, std::max(0, int((centroid[size_t(t) * 3 + best_axis] - cmin) * scale)));
            return b <= best_bin;
          });
          mid = uint32_t(it - order.begin());
          split_found = (mid > begin) && (mid < end);
        }
      }
      if (!split_found && ((count > uint32_t(kBvhMaxLeafTris)) || force_inner)) {
        // degenerate centroids: median split in current order
        mid = begin + count / 2u;
        split_found = (mid > begin) && (mid < end);
      }
    }

    if (!split_found) {
      if (force_inner) {
        // single triangle at the root: inner node with the leaf in child0 and an empty box in child1
        uint32_t node_index = uint32_t(nodes.size());
        nodes.emplace_back();
        BvhNode n = {};
        std::memcpy(n.lo0, bounds.lo, 12);
        std::memcpy(n.hi0, bounds.hi, 12);
        for (int k = 0; k < 3; ++k) {
          n.lo1[k] = std::numeric_limits<float>::max();
          n.hi1[k] = -std::numeric_limits<float>::max();
        }
        n.child0 = make_leaf(begin, end);
        n.child1 = n.child0;
        nodes[node_index] = n;
        return int32_t(node_index);
      }
      return make_leaf(begin, end);
    }

    uint32_t node_index = uint32_t(nodes.size());
    nodes.emplace_back();
    Box b0, b1;
    int32_t c0 = build(begin, mid, b0, false, depth + 1u);
    int32_t c1 = build(mid, end, b1, false, depth + 1u);
    BvhNode n = {};
    std::memcpy(n.lo0, b0.lo, 12);
    std::memcpy(n.hi0, b0.hi, 12);
    std::memcpy(n.lo1, b1.lo, 12);
    std::memcpy(n.hi1, b1.hi, 12);
    n.child0 = c0;
    n.child1 = c1;
    nodes[node_index] = n;
    return int32_t(node_index);
  }
};

}  // namespace

void build_bvh(const float* positions, uint32_t position_stride_bytes, const uint32_t* indices, uint32_t index_stride_bytes, uint32_t tri_count, Bvh& out) {
  out.nodes.clear();
  out.tri_pos.clear();
  out.tri_index.clear();

  Builder b;
  b.positions = positions;
  b.position_stride = position_stride_bytes / 4u;
  b.indices = indices;
  b.index_stride = index_stride_bytes / 4u;
  b.tri_count = tri_count;

  if (tri_count == 0) {
    BvhNode n = {};
    for (int k = 0; k < 3; ++k) {
      n.lo0[k] = n.lo1[k] = std::numeric_limits<float>::max();
      n.hi0[k] = n.hi1[k] = -std::numeric_limits<float>::max();
    }
    // both children point at an (empty-box) leaf that is never reached
    n.child0 = n.child1 = int32_t(~0u);
    out.nodes.push_back(n);
    out.tri_pos.resize(3, F4{0, 0, 0, 0});
    out.tri_index.push_back(0xffffffffu);
    out.max_depth = 1;
    return;
  }

  b.tri_box.resize(tri_count);
  b.centroid.resize(size_t(tri_count) * 3);
  b.order.resize(tri_count);
  for (uint32_t t = 0; t < tri_count; ++t) {
    Box bx;
    for (int k = 0; k < 3; ++k) bx.grow(b.vertex(t, k));
    b.tri_box[t] = bx;
    for (int k = 0; k < 3; ++k) b.centroid[size_t(t) * 3 + k] = 0.5f * (bx.lo[k] + bx.hi[k]);
    b.order[t] = t;
  }
  b.nodes.reserve(tri_count);
  b.leaf_tris.reserve(tri_count);

  Box root_box;
  b.build(0, tri_count, root_box, true);

  out.max_depth = b.max_depth;
  // node order: the top of the tree breadth-first (nodes[0 .. kTopNodes) are the levels every ray walks through: the traversal kernels copy
  // them into shared memory in one bulk transfer, dtrav.cuh), everything below in the builder's depth-first order (a subtree stays contiguous,
  // which is what the caches want once rays have spread out).  Only indices move — boxes, child order and leaves stay, so every traversal visits
  // the same candidates in the same order.
  {
    constexpr size_t kTopNodes = 512;
    const size_t n = b.nodes.size();
    std::vector<uint32_t> order_new;
    order_new.reserve(n);
    std::vector<uint32_t> new_index(n, 0xffffffffu);
    order_new.push_back(0u);
    new_index[0] = 0u;
    for (size_t head = 0; (head < order_new.size()) && (order_new.size() < kTopNodes); ++head) {
      const BvhNode& nd = b.nodes[order_new[head]];
      const int32_t kids[2] = {nd.child0, nd.child1};
      for (int32_t c : kids) {
        if ((c >= 0) && (new_index[uint32_t(c)] == 0xffffffffu) && (order_new.size() < kTopNodes)) {
          new_index[uint32_t(c)] = uint32_t(order_new.size());
          order_new.push_back(uint32_t(c));
        }
      }
    }
    for (size_t k = 0; k < n; ++k) {
      if (new_index[k] == 0xffffffffu) {
        new_index[k] = uint32_t(order_new.size());
        order_new.push_back(uint32_t(k));
      }
    }
    std::vector<BvhNode> reordered(n);
    for (size_t k = 0; k < n; ++k) {
      BvhNode nd = b.nodes[order_new[k]];
      if (nd.child0 >= 0) nd.child0 = int32_t(new_index[uint32_t(nd.child0)]);
      if (nd.child1 >= 0) nd.child1 = int32_t(new_index[uint32_t(nd.child1)]);
      reordered[k] = nd;
    }
    b.nodes = std::move(reordered);
  }
  out.nodes = std::move(b.nodes);
  out.tri_index = std::move(b.leaf_tris);
  out.tri_pos.resize(out.tri_index.size() * 3);
  for (size_t s = 0; s < out.tri_index.size(); ++s) {
    uint32_t t = out.tri_index[s];
    for (int k = 0; k < 3; ++k) {
      const float* p = b.vertex(t, k);
      out.tri_pos[s * 3 + k] = F4{p[0], p[1], p[2], (k == 0) ? u2f(t) : 0.0f};
    }
  }
}

namespace {

struct WideChild {
  float lo[3], hi[3];
  int32_t ref;  // BVH2 reference: >= 0 inner node, < 0 leaf
};
float box_area(const WideChild& c) {
  float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
  return dx * dy + dy * dz + dz * dx;
}
void children_of(const BvhNode& n, WideChild& a, WideChild& b) {
  std::memcpy(a.lo, n.lo0, 12);
  std::memcpy(a.hi, n.hi0, 12);
  a.ref = n.child0;
  std::memcpy(b.lo, n.lo1, 12);
  std::memcpy(b.hi, n.hi1, 12);
  b.ref = n.child1;
}

struct WideBuilder {
  const Bvh& bvh;
  std::vector<WideNode> nodes;
  uint32_t max_stack = 0;

  uint32_t build(int32_t bvh2_node, uint32_t stack_above) {
    // the BVH2 node's children, the widest inner one opened until four are on the table
    WideChild kids[4];
    int count = 2;
    children_of(bvh.nodes[size_t(bvh2_node)], kids[0], kids[1]);
    if (kids[1].ref == kids[0].ref) count = 1;  // single-triangle root: both slots name the same leaf
    while (count < 4) {
      int best = -1;
      float best_area = -1.0f;
      for (int k = 0; k < count; ++k) {
        if ((kids[k].ref >= 0) && (box_area(kids[k]) > best_area)) {
          best_area = box_area(kids[k]);
          best = k;
        }
      }
      if (best < 0) break;
      WideChild a, b;
      children_of(bvh.nodes[size_t(kids[best].ref)], a, b);
      kids[best] = a;
      kids[count++] = b;
    }
    const uint32_t index = uint32_t(nodes.size());
    nodes.emplace_back();
    WideNode w = {};
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
      lo[a] = kids[0].lo[a];
      hi[a] = kids[0].hi[a];
      for (int k = 1; k < count; ++k) {
        lo[a] = std::min(lo[a], kids[k].lo[a]);
        hi[a] = std::max(hi[a], kids[k].hi[a]);
      }
      w.origin[a] = lo[a];
      // smallest power-of-two step with 255 steps covering the extent
      float extent = std::max(hi[a] - lo[a], 1.0e-30f);
      int e = 0;
      (void)std::frexp(extent / 255.0f, &e);  // extent / 255 = m * 2^e, m in [0.5, 1)  =>  2^e >= extent / 255
      e = std::max(-126, std::min(127, e));
      w.exp[a] = uint8_t(e + 127);
      const float step = std::ldexp(1.0f, e);
      for (int k = 0; k < count; ++k) {
        int ql = int(std::floor((kids[k].lo[a] - lo[a]) / step));
        int qh = int(std::ceil((kids[k].hi[a] - lo[a]) / step));
        ql = std::max(0, std::min(255, ql));
        qh = std::max(0, std::min(255, qh));
        // conservative in float arithmetic: the decoded corner (origin + q * step, as the device computes it) must not cut into the exact box
        while ((ql > 0) && (lo[a] + float(ql) * step > kids[k].lo[a])) ql--;
        while ((qh < 255) && (lo[a] + float(qh) * step < kids[k].hi[a])) qh++;
        w.qlo[k][a] = uint8_t(ql);
        w.qhi[k][a] = uint8_t(qh);
      }
    }
    w.count = uint8_t(count);
    max_stack = std::max(max_stack, stack_above + uint32_t(count - 1));
    for (int k = 0; k < count; ++k) {
      // when child k is being traversed, the other (count - 1 - k at most) siblings still wait on the stack: bound by count - 1
      w.child[k] = (kids[k].ref >= 0) ? int32_t(build(kids[k].ref, stack_above + uint32_t(count - 1))) : kids[k].ref;
    }
    for (int k = count; k < 4; ++k) w.child[k] = int32_t(0x80000000u);
    nodes[index] = w;
    return index;
  }
};

}  // namespace

void build_wide_bvh(const Bvh& bvh, WideBvh& out) {
  WideBuilder b{bvh};
  b.nodes.reserve(bvh.nodes.size() / 2 + 1);
  b.build(0, 0u);
  // top levels breadth-first (staged in shared memory by the persistent kernels), the rest in the builder's depth-first order
  constexpr size_t kTopNodes = 512;
  const size_t n = b.nodes.size();
  std::vector<uint32_t> order_new;
  order_new.reserve(n);
  std::vector<uint32_t> new_index(n, 0xffffffffu);
  order_new.push_back(0u);
  new_index[0] = 0u;
  for (size_t head = 0; (head < order_new.size()) && (order_new.size() < kTopNodes); ++head) {
    const WideNode& nd = b.nodes[order_new[head]];
    for (int k = 0; k < nd.count; ++k) {
      int32_t c = nd.child[k];
      if ((c >= 0) && (new_index[uint32_t(c)] == 0xffffffffu) && (order_new.size() < kTopNodes)) {
        new_index[uint32_t(c)] = uint32_t(order_new.size());
        order_new.push_back(uint32_t(c));
      }
    }
  }
  for (size_t k = 0; k < n; ++k) {
    if (new_index[k] == 0xffffffffu) {
      new_index[k] = uint32_t(order_new.size());
      order_new.push_back(uint32_t(k));
    }
  }
  out.nodes.resize(n);
  for (size_t k = 0; k < n; ++k) {
    WideNode nd = b.nodes[order_new[k]];
    for (int c = 0; c < nd.count; ++c)
      if (nd.child[c] >= 0) nd.child[c] = int32_t(new_index[uint32_t(nd.child[c])]);
    out.nodes[k] = nd;
  }
  out.max_stack = b.max_stack;
}

}  // namespace etxb
