// scene_loader_tangents.inl — per-vertex tangent frames of a mesh WITH texture coordinates (included by scene_loader.cpp inside its anonymous namespace).
//
// The reference hands such a mesh to Mikkelsen's tangent-space generator (build_tangents, scene_representation.cxx:337-398, thirdparty/mikktspace —
// genTangSpaceDefault: 180 degree angular threshold) and takes `tangent, sign` per triangle corner.  That published algorithm, restated:
//   1. corners with identical (position, normal, texture coordinate) are one vertex;
//   2. per triangle: the first-order derivatives dP/ds, dP/dt, their magnitudes, whether the mapping preserves orientation; a triangle with a zero
//      texture area "groups with anything";
//   3. triangles that share an edge traversed in opposite directions are neighbours (edges ordered by (lower vertex, higher vertex, triangle));
//   4. around every vertex, the triangles reachable through neighbours with the same orientation form a group (depth first, left neighbour first);
//   5. per group and corner: the members whose projected derivatives are within the angular threshold form a sub-group; its frame is the angle-
//      weighted sum of the members' projected, normalised derivatives, summed in ascending triangle order;
//   6. a corner written twice averages the two frames.
// Every float expression is evaluated in the generator's order, so the frames agree bit for bit — with one exception that is not reproduced: the
// generator leaves the LAST run of its edge list unsorted (its sub-sorts are triggered by a key change, which never comes for the final run), so
// for the few edges whose lower vertex is the highest-numbered one a neighbour can go unnoticed there; here the list is fully sorted.

namespace tangents {

struct TriInfo {
  F3 os, ot;
  float mag_s = 0.0f, mag_t = 0.0f;
  int neighbour[3] = {-1, -1, -1};
  int group[3] = {-1, -1, -1};
  bool orient_preserving = false, group_with_any = true;
};

struct Group {
  int vertex = -1;
  bool orient_preserving = false;
  std::vector<int> faces;
};

struct Frame {
  F3 os = {1.0f, 0.0f, 0.0f}, ot = {0.0f, 1.0f, 0.0f};
  float mag_s = 1.0f, mag_t = 1.0f;
  int counter = 0;
  bool orient = false;
};

inline bool not_zero(float x) { return fabsf(x) > FLT_MIN; }
inline bool v_not_zero(F3 v) { return not_zero(v.x) || not_zero(v.y) || not_zero(v.z); }
inline F3 scaled(float s, F3 v) { return {s * v.x, s * v.y, s * v.z}; }
inline float vdot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 normalized(F3 v) { return scaled(1 / sqrtf(v.x * v.x + v.y * v.y + v.z * v.z), v); }
inline bool same(F3 a, F3 b) { return (a.x == b.x) && (a.y == b.y) && (a.z == b.z); }
inline F3 project(F3 v, F3 n) {
  F3 p = v - scaled(vdot(n, v), n);
  return v_not_zero(p) ? normalized(p) : p;
}

struct CornerKey {
  uint32_t w[8];
  bool operator==(const CornerKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct CornerHash {
  size_t operator()(const CornerKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint32_t x : k.w) h = (h ^ x) * 1099511628211ull;
    return size_t(h);
  }
};

// `corner_vertex[3 f + i]` = index into `v` of corner i of triangle f.  Writes tan / btn of the vertices whose tangent is not valid yet.
void generate(std::vector<etxb_vertex>& v, const std::vector<etxb_triangle>& tris) {
  const int n_tris = int(tris.size());
  if (n_tris == 0) return;
  // 1. weld: `id[3 f + i]` = the first corner with the same attributes (== on floats: a negative zero equals a positive one, a NaN nothing)
  std::vector<int> id(size_t(n_tris) * 3);
  {
    std::unordered_map<CornerKey, int, CornerHash> first;
    first.reserve(size_t(n_tris) * 3);
    for (int c = 0; c < n_tris * 3; ++c) {
      const etxb_vertex& x = v[tris[c / 3].i[c % 3]];
      const float a[8] = {x.pos[0], x.pos[1], x.pos[2], x.nrm[0], x.nrm[1], x.nrm[2], x.tex[0], x.tex[1]};
      CornerKey key;
      bool has_nan = false;
      for (int k = 0; k < 8; ++k) {
        float f = (a[k] == 0.0f) ? 0.0f : a[k];
        has_nan = has_nan || (f != f);
        memcpy(&key.w[k], &f, 4);
      }
      if (has_nan) {
        id[c] = c;
        continue;
      }
      auto it = first.emplace(key, c);
      id[c] = it.first->second;
    }
  }
  auto vertex_of = [&](int corner) -> const etxb_vertex& { return v[tris[corner / 3].i[corner % 3]]; };
  auto position = [&](int corner) { return load3(vertex_of(corner).pos); };
  auto normal = [&](int corner) { return load3(vertex_of(corner).nrm); };

  // 2. per-triangle derivatives
  std::vector<TriInfo> info(n_tris);
  for (int f = 0; f < n_tris; ++f) {
    TriInfo& ti = info[f];
    const etxb_vertex &a = vertex_of(id[3 * f + 0]), &b = vertex_of(id[3 * f + 1]), &c = vertex_of(id[3 * f + 2]);
    const float t21x = b.tex[0] - a.tex[0], t21y = b.tex[1] - a.tex[1], t31x = c.tex[0] - a.tex[0], t31y = c.tex[1] - a.tex[1];
    const F3 d1 = load3(b.pos) - load3(a.pos), d2 = load3(c.pos) - load3(a.pos);
    const float signed_area = t21x * t31y - t21y * t31x;
    F3 os = scaled(t31y, d1) - scaled(t21y, d2);
    F3 ot = scaled(-t31x, d1) + scaled(t21x, d2);
    ti.orient_preserving = signed_area > 0;
    if (not_zero(signed_area)) {
      const float abs_area = fabsf(signed_area);
      const float len_os = sqrtf(os.x * os.x + os.y * os.y + os.z * os.z), len_ot = sqrtf(ot.x * ot.x + ot.y * ot.y + ot.z * ot.z);
      const float sign = ti.orient_preserving ? 1.0f : (-1.0f);
      if (not_zero(len_os)) ti.os = scaled(sign / len_os, os);
      if (not_zero(len_ot)) ti.ot = scaled(sign / len_ot, ot);
      ti.mag_s = len_os / abs_area;
      ti.mag_t = len_ot / abs_area;
      if (not_zero(ti.mag_s) && not_zero(ti.mag_t)) ti.group_with_any = false;
    }
  }

  // 3. neighbours over shared edges
  {
    struct Edge {
      int i0, i1, f;
    };
    std::vector<Edge> edges(size_t(n_tris) * 3);
    for (int f = 0; f < n_tris; ++f) {
      for (int i = 0; i < 3; ++i) {
        int a = id[3 * f + i], b = id[3 * f + (i < 2 ? (i + 1) : 0)];
        edges[size_t(f) * 3 + i] = {a < b ? a : b, !(a < b) ? a : b, f};
      }
    }
    std::sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) { return a.i0 != b.i0 ? a.i0 < b.i0 : (a.i1 != b.i1 ? a.i1 < b.i1 : a.f < b.f); });
    // which of a triangle's edges joins the two vertices, and in which direction the triangle walks it
    auto edge_of = [&](int f, int v0, int v1, int& from, int& to) {
      const int* ix = &id[3 * f];
      if (ix[0] == v0 || ix[0] == v1) {
        if (ix[1] == v0 || ix[1] == v1) {
          from = ix[0], to = ix[1];
          return 0;
        }
        from = ix[2], to = ix[0];
        return 2;
      }
      from = ix[1], to = ix[2];
      return 1;
    };
    const size_t n = edges.size();
    for (size_t i = 0; i < n; ++i) {
      const Edge& e = edges[i];
      int from_a, to_a;
      const int edge_a = edge_of(e.f, e.i0, e.i1, from_a, to_a);
      if (info[e.f].neighbour[edge_a] != -1) continue;
      for (size_t j = i + 1; j < n && edges[j].i0 == e.i0 && edges[j].i1 == e.i1; ++j) {
        int from_b, to_b;
        const int t = edges[j].f, edge_b = edge_of(t, e.i0, e.i1, from_b, to_b);
        if (from_a == to_b && to_a == from_b && info[t].neighbour[edge_b] == -1) {
          info[e.f].neighbour[edge_a] = t;
          info[t].neighbour[edge_b] = e.f;
          break;
        }
      }
    }
  }

  // 4. groups around vertices
  std::vector<Group> groups;
  {
    std::vector<int> stack;
    auto visit = [&](int tri, int g) {  // one entry of the depth-first walk; returns the corner when the triangle joined the group, else -1
      TriInfo& ti = info[tri];
      const int rep = groups[g].vertex;
      const int* ix = &id[3 * tri];
      const int i = (ix[0] == rep) ? 0 : ((ix[1] == rep) ? 1 : ((ix[2] == rep) ? 2 : -1));
      if (i < 0 || ti.group[i] != -1) return -1;
      if (ti.group_with_any && ti.group[0] == -1 && ti.group[1] == -1 && ti.group[2] == -1) ti.orient_preserving = groups[g].orient_preserving;
      if (ti.orient_preserving != groups[g].orient_preserving) return -1;
      groups[g].faces.push_back(tri);
      ti.group[i] = g;
      return i;
    };
    for (int f = 0; f < n_tris; ++f) {
      for (int i = 0; i < 3; ++i) {
        if (info[f].group_with_any || info[f].group[i] != -1) continue;
        const int g = int(groups.size());
        groups.emplace_back();
        groups[g].vertex = id[3 * f + i];
        groups[g].orient_preserving = info[f].orient_preserving;
        groups[g].faces.push_back(f);
        info[f].group[i] = g;
        // left neighbour's whole fan first, then the right one's: the order the members are listed in
        stack.clear();
        const int right = info[f].neighbour[i > 0 ? (i - 1) : 2], left = info[f].neighbour[i];
        if (right >= 0) stack.push_back(right);
        if (left >= 0) stack.push_back(left);
        while (!stack.empty()) {
          const int tri = stack.back();
          stack.pop_back();
          const int corner = visit(tri, g);
          if (corner < 0) continue;
          const int r = info[tri].neighbour[corner > 0 ? (corner - 1) : 2], l = info[tri].neighbour[corner];
          if (r >= 0) stack.push_back(r);
          if (l >= 0) stack.push_back(l);
        }
      }
    }
  }

  // 5. + 6. frames per corner
  const float threshold_cos = float(cos((180.0f * float(3.14159265358979323846)) / 180.0f));  // genTangSpaceDefault: 180 degrees
  std::vector<Frame> frames(size_t(n_tris) * 3);
  std::vector<int> members;
  std::vector<std::vector<int>> unique_members;
  std::vector<Frame> unique_frames;
  for (size_t g = 0; g < groups.size(); ++g) {
    const Group& group = groups[g];
    unique_members.clear();
    unique_frames.clear();
    for (int f : group.faces) {
      const TriInfo& tf = info[f];
      const int index = (tf.group[0] == int(g)) ? 0 : ((tf.group[1] == int(g)) ? 1 : 2);
      const F3 n = normal(id[3 * f + index]);
      const F3 os = project(tf.os, n), ot = project(tf.ot, n);
      members.clear();
      for (int t : group.faces) {
        const TriInfo& tt = info[t];
        const F3 os2 = project(tt.os, n), ot2 = project(tt.ot, n);
        const bool any = tf.group_with_any || tt.group_with_any;
        const float cos_s = vdot(os, os2), cos_t = vdot(ot, ot2);
        if (any || (f == t) || (cos_s > threshold_cos && cos_t > threshold_cos)) members.push_back(t);
      }
      std::sort(members.begin(), members.end());
      size_t l = 0;
      while (l < unique_members.size() && unique_members[l] != members) ++l;
      if (l == unique_members.size()) {
        Frame res;
        res.os = res.ot = {0.0f, 0.0f, 0.0f};
        res.mag_s = res.mag_t = 0.0f;
        float angle_sum = 0.0f;
        for (int m : members) {
          const TriInfo& tm = info[m];
          if (tm.group_with_any) continue;  // only triangles with a texture area contribute
          const int* ix = &id[3 * m];
          const int i = (ix[0] == group.vertex) ? 0 : ((ix[1] == group.vertex) ? 1 : 2);
          const F3 nm = normal(ix[i]);
          const F3 mos = project(tm.os, nm), mot = project(tm.ot, nm);
          const F3 p0 = position(ix[i > 0 ? (i - 1) : 2]), p1 = position(ix[i]), p2 = position(ix[i < 2 ? (i + 1) : 0]);
          const F3 v1 = project(p0 - p1, nm), v2 = project(p2 - p1, nm);
          float c = vdot(v1, v2);
          c = c > 1 ? 1 : (c < (-1) ? (-1) : c);
          const float angle = float(acos(double(c)));
          res.os = res.os + scaled(angle, mos);
          res.ot = res.ot + scaled(angle, mot);
          res.mag_s += (angle * tm.mag_s);
          res.mag_t += (angle * tm.mag_t);
          angle_sum += angle;
        }
        if (v_not_zero(res.os)) res.os = normalized(res.os);
        if (v_not_zero(res.ot)) res.ot = normalized(res.ot);
        if (angle_sum > 0) {
          res.mag_s /= angle_sum;
          res.mag_t /= angle_sum;
        }
        unique_members.push_back(members);
        unique_frames.push_back(res);
      }
      Frame& out = frames[size_t(f) * 3 + index];
      const Frame& sub = unique_frames[l];
      if (out.counter == 1) {
        Frame avg;
        if (out.mag_s == sub.mag_s && out.mag_t == sub.mag_t && same(out.os, sub.os) && same(out.ot, sub.ot)) {
          avg = out;
        } else {
          avg.mag_s = 0.5f * (out.mag_s + sub.mag_s);
          avg.mag_t = 0.5f * (out.mag_t + sub.mag_t);
          avg.os = out.os + sub.os;
          avg.ot = out.ot + sub.ot;
          if (v_not_zero(avg.os)) avg.os = normalized(avg.os);
          if (v_not_zero(avg.ot)) avg.ot = normalized(avg.ot);
        }
        out = avg;
        out.counter = 2;
      } else {
        out = sub;
        out.counter = 1;
      }
      out.orient = group.orient_preserving;
    }
  }

  // the reference's setTSpaceBasic callback (:385-392)
  for (int f = 0; f < n_tris; ++f) {
    for (int i = 0; i < 3; ++i) {
      etxb_vertex& x = v[tris[f].i[i]];
      if (valid_vector(load3(x.tan))) continue;
      const Frame& fr = frames[size_t(f) * 3 + i];
      const F3 tan = normalize(fr.os);
      store3(x.tan, tan);
      store3(x.btn, normalize(cross(tan, load3(x.nrm)) * (fr.orient ? 1.0f : (-1.0f))));
    }
  }
}

}  // namespace tangents
