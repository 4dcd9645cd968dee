/* TEST INFRASTRUCTURE ONLY - never linked into libnmr.so.
 * Drives the reference's OWN tangent generator (dependencies/MikkTSpace/mikktspace.c, compiled where it lies under /root/reference
 * by oracle/build_ref.py into oracle/_ref/libmikk_ref.so) through callbacks that read an indexed triangle list and store the result
 * per index, the way MikkTSpaceHandler does (S/gltf_mikktspace_handler.cpp:19-66: three vertices per face, vertex = indices[3 f + k],
 * setTSpaceBasic writes tangents[index] = (tangent, sign)).  Used to pin nmr_mikk_tangents bit for bit. */
#include <stddef.h>
#include <stdint.h>
#include "mikktspace.h"

typedef struct { const float* pos; const float* nrm; const float* uv; const uint32_t* idx; int n_face; float* out; } MeshView;

static const MeshView* view(const SMikkTSpaceContext* c) { return (const MeshView*)c->m_pUserData; }
static int n_faces(const SMikkTSpaceContext* c) { return view(c)->n_face; }
static int n_face_verts(const SMikkTSpaceContext* c, const int f) { (void)c; (void)f; return 3; }
static void get3(const float* a, const uint32_t v, float* o) { o[0] = a[v * 3]; o[1] = a[v * 3 + 1]; o[2] = a[v * 3 + 2]; }
static void position(const SMikkTSpaceContext* c, float o[], const int f, const int k) { get3(view(c)->pos, view(c)->idx[f * 3 + k], o); }
static void normal(const SMikkTSpaceContext* c, float o[], const int f, const int k) { get3(view(c)->nrm, view(c)->idx[f * 3 + k], o); }
static void texcoord(const SMikkTSpaceContext* c, float o[], const int f, const int k) {
    const uint32_t v = view(c)->idx[f * 3 + k]; o[0] = view(c)->uv[v * 2]; o[1] = view(c)->uv[v * 2 + 1];
}
static void store(const SMikkTSpaceContext* c, const float t[], const float sign, const int f, const int k) {
    float* o = view(c)->out + (size_t)view(c)->idx[f * 3 + k] * 4; o[0] = t[0]; o[1] = t[1]; o[2] = t[2]; o[3] = sign;
}

__attribute__((visibility("default")))
int ref_mikk_tangents(const float* pos, const float* nrm, const float* uv, const uint32_t* idx, int64_t n_idx, float* out) {
    MeshView mv = {pos, nrm, uv, idx, (int)(n_idx / 3), out};
    SMikkTSpaceInterface itf = {0};
    itf.m_getNumFaces = n_faces; itf.m_getNumVerticesOfFace = n_face_verts; itf.m_getPosition = position;
    itf.m_getNormal = normal; itf.m_getTexCoord = texcoord; itf.m_setTSpaceBasic = store;
    SMikkTSpaceContext ctx = {&itf, &mv};
    return genTangSpaceDefault(&ctx) ? 0 : 1;
}
