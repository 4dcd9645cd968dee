// Tangent frames of an indexed triangle mesh that has no TANGENT attribute: Mikkelsen's tangent-space method as the reference
// applies it (S/gltf_scene.cpp:150-155 -> MikkTSpaceHandler::calcTangents, S/gltf_mikktspace_handler.cpp:14-66 ->
// genTangSpaceDefault of the vendored dependencies/MikkTSpace/mikktspace.c, angular threshold 180 degrees).  Triangles only
// (glTF mode 4), so the quad branches of the published method (shortest-diagonal split, averaging two spaces of a quad corner,
// quads with one degenerate half) never run and are not restated.
//
// The steps, in the published method's order, because the result depends on them bit for bit:
//   1. corners whose position, normal and texture coordinate compare equal (operator ==) are ONE vertex, named after one of its
//      corners.  WHICH corner is decided by the method's spatial weld (2048 cells along the longest axis, then median splits that
//      swap entries), and matters: the names are sort keys in step 4;
//   2. triangles with two equal corner positions are set aside; the others keep their order;
//   3. per triangle: first-order derivatives dP/ds, dP/dt from the UV parametrisation, normalised, with the sign of the UV area;
//      a triangle without UV area "groups with anything";
//   4. triangles are neighbours over an edge they traverse in opposite directions: the edge list is sorted by (lower name, higher
//      name, triangle) with the method's own quicksort (pivot from a fixed pseudo-random sequence) and partners are looked for among
//      ADJACENT entries only.  The method never sorts the LAST run of each pass, so the edges at the vertex with the highest lower
//      name stay in quicksort order and can miss their partner - reproduced here, pass by pass, because it changes the groups;
//   5. around each vertex, triangles reachable over such edges with the same UV orientation form a group (depth-first, the edge
//      leaving the vertex before the edge entering it); a group-with-anything triangle takes the orientation of the first group
//      that reaches it;
//   6. inside a group a corner's sub-group holds the members whose projected derivatives are less than 180 degrees from its own;
//      the sub-group's tangent is the angle-weighted sum of its members' projected derivatives, in ascending triangle order;
//   7. corners of set-aside triangles copy the frame of the first ordinary corner on the same vertex;
//   8. the frame of corner (f, k) is stored at tangents[index[3 f + k]] in face order: the last face touching an index wins
//      (setTSpaceBasic, S/gltf_mikktspace_handler.cpp:61-66).
// fp32 throughout, unfused (the library is built with -ffp-contract=off); the angle is acos in double, rounded to float.
// Pinned bit for bit against the reference's own mikktspace.c in tests/test_host_cpu.py (golden vectors, and the file itself compiled
// as a test-side library).
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <algorithm>
#include <unordered_map>
#include <vector>

#include "host.h"

namespace nmr {
namespace {

struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(float s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float len(V3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
inline bool same(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool nz(float v) { return fabsf(v) > FLT_MIN; }
inline bool nz(V3 v) { return nz(v.x) || nz(v.y) || nz(v.z); }
inline V3 unit(V3 v) { return mul(1 / len(v), v); }
inline V3 unit_if_nz(V3 v) { return nz(v) ? unit(v) : v; }
inline V3 tangential(V3 v, V3 n) { return sub(v, mul(dot(n, v), n)); }     // v minus its part along n

struct Frame { V3 s{1.f, 0.f, 0.f}; bool preserving = false; };

struct Tri {
    int v[3];                   // welded vertex of each corner
    int face;                   // face number in the index buffer
    V3 ds{0, 0, 0}, dt{0, 0, 0};
    float mag_s = 0, mag_t = 0;
    bool preserving = false, any = true;
    int nb[3] = {-1, -1, -1};   // neighbour over the edge k -> k+1
    int group[3] = {-1, -1, -1};
};

struct Group { int vertex; bool preserving; std::vector<int> tris; };
constexpr size_t kMaxFan = 8192;

// A vertex is named by one of its corners: name = (face << 2) | corner, as the method packs it.
struct Mesh {
    const float* pos; const float* nrm; const float* uv; const uint32_t* idx;
    uint32_t vertex(int name) const { return idx[(size_t)(name >> 2) * 3 + (name & 3)]; }
    V3 P(int name) const { const uint32_t v = vertex(name); return {pos[v * 3], pos[v * 3 + 1], pos[v * 3 + 2]}; }
    V3 N(int name) const { const uint32_t v = vertex(name); return {nrm[v * 3], nrm[v * 3 + 1], nrm[v * 3 + 2]}; }
    float U(int name, int k) const { return uv[(size_t)vertex(name) * 2 + k]; }
    bool equal(int a, int b) const {
        const uint32_t va = vertex(a), vb = vertex(b);
        for (int k = 0; k < 3; ++k) if (!(pos[va * 3 + k] == pos[vb * 3 + k]) || !(nrm[va * 3 + k] == nrm[vb * 3 + k])) return false;
        return uv[va * 2] == uv[vb * 2] && uv[va * 2 + 1] == uv[vb * 2 + 1];
    }
};

// step 1 ---------------------------------------------------------------------------------------------------------------------
struct Spot { float p[3]; int corner; };

inline int longest_axis(const float lo[3], const float hi[3]) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return (dy > dx && dy > dz) ? 1 : (dz > dx ? 2 : 0);
}

// Spots [L, R] share a cell.  Split at the middle of their longest extent (entries swapped pairwise from both ends) until the
// middle is no longer strictly inside; there, every spot takes the name of the first equal spot in front of it.
void weld_range(const Mesh& m, std::vector<int>& name, Spot* s, int L, int R) {
    float lo[3], hi[3];
    for (int c = 0; c < 3; ++c) lo[c] = hi[c] = s[L].p[c];
    for (int l = L + 1; l <= R; ++l) for (int c = 0; c < 3; ++c) {
        if (lo[c] > s[l].p[c]) lo[c] = s[l].p[c];
        if (hi[c] < s[l].p[c]) hi[c] = s[l].p[c];
    }
    const int ax = longest_axis(lo, hi);
    const float mid = 0.5f * (hi[ax] + lo[ax]);
    if (!std::isfinite(mid)) return;
    if (mid >= hi[ax] || mid <= lo[ax]) {
        for (int l = L; l <= R; ++l)
            for (int e = L; e < l; ++e)
                if (m.equal(name[s[l].corner], name[s[e].corner])) { name[s[l].corner] = name[s[e].corner]; break; }
        return;
    }
    int a = L, b = R;
    while (a < b) {
        bool left_ready = false, right_ready = false;
        while (!left_ready && a < b) { left_ready = !(s[a].p[ax] < mid); if (!left_ready) ++a; }
        while (!right_ready && a < b) { right_ready = s[b].p[ax] < mid; if (!right_ready) --b; }
        if (left_ready && right_ready) { std::swap(s[a], s[b]); ++a; --b; }
    }
    if (a == b) { if (s[b].p[ax] < mid) ++a; else --b; }
    if (L < b) weld_range(m, name, s, L, b);
    if (a < R) weld_range(m, name, s, a, R);
}

std::vector<int> weld(const Mesh& m, int n_face) {
    const int n = n_face * 3;
    std::vector<int> name(n);
    for (int f = 0; f < n_face; ++f) for (int k = 0; k < 3; ++k) name[f * 3 + k] = (f << 2) | k;
    float lo[3], hi[3];
    { const V3 p = m.P(name[0]); lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z; }
    for (int c = 1; c < n; ++c) {
        const V3 p = m.P(name[c]); const float q[3] = {p.x, p.y, p.z};
        for (int a = 0; a < 3; ++a) { if (lo[a] > q[a]) lo[a] = q[a]; else if (hi[a] < q[a]) hi[a] = q[a]; }
    }
    const int ax = longest_axis(lo, hi);
    const float fmin = lo[ax], fmax = hi[ax];
    constexpr int CELLS = 2048;
    auto cell_of = [&](int c) {
        const V3 p = m.P(name[c]);
        const float v = ax == 0 ? p.x : (ax == 1 ? p.y : p.z);
        const float fi = (float)CELLS * ((v - fmin) / (fmax - fmin));
        // (int) of a NaN or of a value outside int's range is INT_MIN on x86-64 (cvttss2si), which the method then clamps to cell 0
        const int i = (fi >= -2147483648.f && fi < 2147483648.f) ? (int)fi : INT32_MIN;
        return i < CELLS ? (i >= 0 ? i : 0) : CELLS - 1;
    };
    std::vector<int> cell(n), start(CELLS + 1, 0);
    for (int c = 0; c < n; ++c) { cell[c] = cell_of(c); ++start[cell[c] + 1]; }
    for (int k = 0; k < CELLS; ++k) start[k + 1] += start[k];
    std::vector<int> table(n), fill(start.begin(), start.end() - 1);
    for (int c = 0; c < n; ++c) table[fill[cell[c]]++] = c;
    std::vector<Spot> spots;
    for (int k = 0; k < CELLS; ++k) {
        const int cnt = start[k + 1] - start[k];
        if (cnt < 2) continue;
        spots.resize(cnt);
        for (int e = 0; e < cnt; ++e) {
            const int c = table[start[k] + e];
            const V3 p = m.P(name[c]);
            spots[e] = {{p.x, p.y, p.z}, c};
        }
        weld_range(m, name, spots.data(), 0, cnt - 1);
    }
    return name;
}

// step 4 ---------------------------------------------------------------------------------------------------------------------
struct Edge { int key[3]; };        // lower name, higher name, triangle

// The method's quicksort on one key: two entries are compared directly; otherwise the pivot is the entry at (seed' mod n), seed'
// from a fixed scramble of the seed handed down, and a Hoare partition follows.
void sort_edges(Edge* e, int left, int right, int ch, uint32_t seed) {
    // (the two halves of a partition are sorted independently with the same seed, so the smaller one is taken by recursion and the
    //  larger one by the loop: the same result as two recursive calls with a call depth of log2 n whatever the keys are)
    while (true) {
        const int n = right - left + 1;
        if (n < 2) return;
        if (n == 2) { if (e[left].key[ch] > e[right].key[ch]) std::swap(e[left], e[right]); return; }
        const uint32_t r = seed & 31;
        const uint32_t rot = r ? ((seed << r) | (seed >> (32 - r))) : seed;
        seed = seed + rot + 3;
        int a = left, b = right;
        const int pivot = e[left + (int)(seed % (uint32_t)n)].key[ch];
        do {
            while (e[a].key[ch] < pivot) ++a;
            while (e[b].key[ch] > pivot) --b;
            if (a <= b) { std::swap(e[a], e[b]); ++a; --b; }
        } while (a <= b);
        const bool lo = left < b, hi = a < right;
        if (lo && hi) {
            if (b - left < right - a) { sort_edges(e, left, b, ch, seed); left = a; }
            else { sort_edges(e, a, right, ch, seed); right = b; }
        } else if (lo) right = b;
        else if (hi) left = a;
        else return;
    }
}

// step 5: depth-first growth of one group around its vertex, from the triangles `first` and `second` (either may be -1).  The order
// is the published recursion's - a triangle joins, then everything behind the edge leaving the vertex, then behind the edge
// entering it - walked with an explicit stack: a fan of a million triangles around one vertex must not overflow the call stack.
void grow(std::vector<Tri>& tris, Group& G, int g, int first, int second, std::vector<int>& stack) {
    stack.clear();
    if (second >= 0) stack.push_back(second);
    if (first >= 0) stack.push_back(first);
    while (!stack.empty()) {
        const int t = stack.back(); stack.pop_back();
        Tri& T = tris[t];
        const int k = T.v[0] == G.vertex ? 0 : T.v[1] == G.vertex ? 1 : 2;
        if (T.group[k] != -1) continue;                             // already in this group, or in another one
        if (T.any && T.group[0] == -1 && T.group[1] == -1 && T.group[2] == -1) T.preserving = G.preserving;
        if (T.preserving != G.preserving) continue;
        G.tris.push_back(t);
        T.group[k] = g;
        const int out = T.nb[k], in = T.nb[k > 0 ? k - 1 : 2];
        if (in >= 0) stack.push_back(in);
        if (out >= 0) stack.push_back(out);
    }
}

}  // namespace

void mikk_tangents(const float* positions, const float* normals, const float* texcoords, size_t n_vert,
                   const uint32_t* indices, size_t n_idx, float* tangents) {
    const Mesh m{positions, normals, texcoords, indices};
    const size_t n_face = n_idx / 3;
    if (n_face == 0) return;
    if (n_face >= (size_t)1 << 29) throw std::length_error("mikk_tangents: more than 2^29 triangles");      // a vertex name is (face << 2) | corner in an int
    (void)n_vert;
    const std::vector<int> name = weld(m, (int)n_face);

    // step 2
    std::vector<Tri> tris; tris.reserve(n_face);
    std::vector<int> set_aside;
    for (size_t f = 0; f < n_face; ++f) {
        const V3 p0 = m.P(name[f * 3]), p1 = m.P(name[f * 3 + 1]), p2 = m.P(name[f * 3 + 2]);
        if (same(p0, p1) || same(p0, p2) || same(p1, p2)) { set_aside.push_back((int)f); continue; }
        Tri t; t.face = (int)f;
        for (int k = 0; k < 3; ++k) t.v[k] = name[f * 3 + k];
        tris.push_back(t);
    }
    const int n_tri = (int)tris.size();

    // step 3
    for (Tri& t : tris) {
        const V3 p1 = m.P(t.v[0]), p2 = m.P(t.v[1]), p3 = m.P(t.v[2]);
        const float s21 = m.U(t.v[1], 0) - m.U(t.v[0], 0), t21 = m.U(t.v[1], 1) - m.U(t.v[0], 1);
        const float s31 = m.U(t.v[2], 0) - m.U(t.v[0], 0), t31 = m.U(t.v[2], 1) - m.U(t.v[0], 1);
        const V3 e1 = sub(p2, p1), e2 = sub(p3, p1);
        const float area2 = s21 * t31 - t21 * s31;
        const V3 ds = sub(mul(t31, e1), mul(t21, e2)), dt = add(mul(-s31, e1), mul(s21, e2));
        t.preserving = area2 > 0;
        if (!nz(area2)) continue;
        const float a = fabsf(area2), ls = len(ds), lt = len(dt), sign = t.preserving ? 1.0f : -1.0f;
        if (nz(ls)) t.ds = mul(sign / ls, ds);
        if (nz(lt)) t.dt = mul(sign / lt, dt);
        t.mag_s = ls / a; t.mag_t = lt / a;
        if (nz(t.mag_s) && nz(t.mag_t)) t.any = false;
    }

    // step 4
    if (n_tri > 0) {
        std::vector<Edge> edges((size_t)n_tri * 3);
        for (int t = 0; t < n_tri; ++t) for (int k = 0; k < 3; ++k) {
            const int a = tris[t].v[k], b = tris[t].v[k < 2 ? k + 1 : 0];
            edges[(size_t)t * 3 + k] = {{a < b ? a : b, !(a < b) ? a : b, t}};
        }
        const int n_edge = n_tri * 3;
        const uint32_t seed = 39871946u;
        Edge* e = edges.data();
        sort_edges(e, 0, n_edge - 1, 0, seed);
        for (int i = 1, run = 0; i < n_edge; ++i)                      // runs of one lower name - all but the last one
            if (e[run].key[0] != e[i].key[0]) { const int l = run; run = i; sort_edges(e, l, i - 1, 1, seed); }
        for (int i = 1, run = 0; i < n_edge; ++i)                      // runs of one name pair - all but the last one
            if (e[run].key[0] != e[i].key[0] || e[run].key[1] != e[i].key[1]) { const int l = run; run = i; sort_edges(e, l, i - 1, 2, seed); }
        auto edge_of = [&](const Edge& E) {                            // which edge of its triangle an entry is
            const Tri& T = tris[E.key[2]];
            const bool v0 = T.v[0] == E.key[0] || T.v[0] == E.key[1], v1 = T.v[1] == E.key[0] || T.v[1] == E.key[1];
            return v0 ? (v1 ? 0 : 2) : 1;
        };
        for (int i = 0; i < n_edge; ++i) {
            const int ta = e[i].key[2], ka = edge_of(e[i]);
            if (tris[ta].nb[ka] != -1) continue;
            const int from = tris[ta].v[ka];
            for (int j = i + 1; j < n_edge && e[j].key[0] == e[i].key[0] && e[j].key[1] == e[i].key[1]; ++j) {
                const int tb = e[j].key[2], kb = edge_of(e[j]);
                if (tris[tb].nb[kb] != -1 || tris[tb].v[kb] == from) continue;      // claimed, or traversed the same way round
                tris[ta].nb[ka] = tb; tris[tb].nb[kb] = ta;
                break;
            }
        }
    }

    // step 5
    std::vector<Group> groups;
    {
        std::vector<int> stack;
        for (int t = 0; t < n_tri; ++t) for (int k = 0; k < 3; ++k) {
            if (tris[t].any || tris[t].group[k] != -1) continue;
            const int g = (int)groups.size();
            groups.push_back({tris[t].v[k], tris[t].preserving, {}});
            groups[g].tris.push_back(t);
            tris[t].group[k] = g;
            grow(tris, groups[g], g, tris[t].nb[k], tris[t].nb[k > 0 ? k - 1 : 2], stack);
        }
    }

    // step 6
    std::vector<Frame> corner(n_face * 3);
    const float cos_limit = (float)cos((double)((180.0f * (float)M_PI) / 180.0f));
    std::vector<int> members;
    std::vector<std::vector<int>> sub_members;
    std::vector<V3> sub_tangent;
    for (size_t g = 0; g < groups.size(); ++g) {
        const Group& G = groups[g];
        // the sub-groups compare every member with every other one (as the published method does): refuse a vertex whose fan would
        // take minutes - no modelled surface comes near this (the reference overflows its call stack long before)
        if (G.tris.size() > kMaxFan) throw std::length_error("mikk_tangents: more than 8192 triangles of one orientation around a vertex");
        const V3 n = m.N(G.vertex);
        sub_members.clear(); sub_tangent.clear();
        for (int f : G.tris) {
            const Tri& F = tris[f];
            const int k = F.group[0] == (int)g ? 0 : F.group[1] == (int)g ? 1 : 2;
            const V3 fs = unit_if_nz(tangential(F.ds, n)), ft = unit_if_nz(tangential(F.dt, n));
            members.clear();
            for (int t : G.tris) {
                const Tri& T = tris[t];
                const V3 ts = unit_if_nz(tangential(T.ds, n)), tt = unit_if_nz(tangential(T.dt, n));
                const float cs = dot(fs, ts), ct = dot(ft, tt);
                if (F.any || T.any || F.face == T.face || (cs > cos_limit && ct > cos_limit)) members.push_back(t);
            }
            std::sort(members.begin(), members.end());
            size_t l = 0;
            while (l < sub_members.size() && sub_members[l] != members) ++l;
            if (l == sub_members.size()) {
                V3 sum{0, 0, 0};
                for (int t : members) {
                    const Tri& T = tris[t];
                    if (T.any) continue;
                    const int i = T.v[0] == G.vertex ? 0 : T.v[1] == G.vertex ? 1 : 2;
                    const V3 nn = m.N(T.v[i]);
                    const V3 ts = unit_if_nz(tangential(T.ds, nn));
                    const V3 q0 = m.P(T.v[i > 0 ? i - 1 : 2]), q1 = m.P(T.v[i]), q2 = m.P(T.v[i < 2 ? i + 1 : 0]);
                    const V3 a = unit_if_nz(tangential(sub(q0, q1), nn)), b = unit_if_nz(tangential(sub(q2, q1), nn));
                    float c = dot(a, b); c = c > 1 ? 1 : (c < -1 ? -1 : c);
                    const float angle = (float)acos((double)c);
                    sum = add(sum, mul(angle, ts));
                }
                sub_members.push_back(members);
                sub_tangent.push_back(unit_if_nz(sum));
            }
            const size_t c = (size_t)F.face * 3 + k;
            corner[c].s = sub_tangent[l]; corner[c].preserving = G.preserving;
        }
    }

    // step 7
    if (!set_aside.empty()) {
        std::unordered_map<int, size_t> first_corner;                // welded vertex -> first ordinary corner on it
        for (int t = n_tri - 1; t >= 0; --t) for (int k = 2; k >= 0; --k) first_corner[tris[t].v[k]] = (size_t)tris[t].face * 3 + k;
        for (int f : set_aside) for (int k = 0; k < 3; ++k) {
            auto it = first_corner.find(name[(size_t)f * 3 + k]);
            if (it != first_corner.end()) corner[(size_t)f * 3 + k] = corner[it->second];
        }
    }

    // step 8
    for (size_t c = 0; c < n_face * 3; ++c) {
        float* o = tangents + (size_t)indices[c] * 4;
        o[0] = corner[c].s.x; o[1] = corner[c].s.y; o[2] = corner[c].s.z; o[3] = corner[c].preserving ? 1.0f : -1.0f;
    }
}

}  // namespace nmr
