/*
 * nmr_oracle.c - scalar CPU restatement of the reference's hybrid NeRF + mesh render path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (nerf-glasses_b200/, libnmr.so, pynmr)
 * includes, links or calls this file; it is used by tests/, __graft_entry__.smoke() and by
 * bench.py's cpu_baseline / --impl reference legs as the checker / CPU baseline.
 *
 * Every function cites the reference code it restates.  Path shorthands:
 *   S/ = /root/reference/nerf_mesh_renderer/src/
 *   T/ = /root/reference/nerf_mesh_renderer/dependencies/tiny-cuda-nn/
 *
 * Parity status of this oracle (see DESIGN.md "Oracle"):
 *   - floatie removal and the orbit camera are pinned bit-for-bit against the reference's own
 *     headers compiled here (oracle/_ref, tests/test_oracle_vs_ref.py);
 *   - host/device helpers that are HOST_DEVICE in the reference (pixel_to_ray, AABB slab test,
 *     Sobol/LK jitter, Morton code, colour transfer functions) are pinned the same way;
 *   - everything that exists only as __device__/__global__ code in the reference (DDA helpers,
 *     hash-grid kernel, fused MLP, SH kernel, compositing, OptiX programs) cannot run in the
 *     authoring container (no GPU, no OptiX): "parity unpinned" for those, restated from source.
 *
 * Arithmetic conventions (the CUDA product follows the same ones so integer results are bit-exact):
 *   - IEEE fp32, round-to-nearest, NO fused multiply-add (build with -ffp-contract=off);
 *   - sums of three products are ((a*b + c*d) + e*f), the order Eigen's unrolled redux uses;
 *   - fp16 values are kept as uint16 bit patterns; half adds are float adds rounded once to fp16
 *     (exact for binary16 because fp32 carries >= 2*11+2 significand bits);
 *   - MLP layers accumulate in fp32 in k order and round to fp16 at every layer output (the
 *     reference accumulates in fp16 inside wmma, T/src/fully_fused_mlp.cu:66-68; the product uses
 *     fp32 accumulators in tensor memory, so this oracle documents the more precise variant).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

#define NERF_GRIDSIZE 128u
#define NERF_CASCADES 8u
#define GRID_CELLS (NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE)
#define MAX_LEVELS 16

/* ------------------------------------------------------------------------------------------ */
/* fp16 helpers                                                                                 */
/* ------------------------------------------------------------------------------------------ */
static inline float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do { man <<= 1; ++e; } while ((man & 0x400u) == 0);
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f; memcpy(&f, &bits, 4); return f;
}

/* round-to-nearest-even, like __float2half_rn */
static inline uint16_t f2h(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7FFFFFFFu;
    if (ax >= 0x7F800000u) {                       /* inf / nan */
        return (uint16_t)(sign | 0x7C00u | ((ax > 0x7F800000u) ? 0x200u : 0u));
    }
    if (ax >= 0x477FF000u) {                       /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7C00u);
    }
    if (ax < 0x38800000u) {                        /* subnormal half or zero */
        if (ax < 0x33000000u) return (uint16_t)sign; /* < 2^-25 -> 0 (2^-25 itself ties to even = 0) */
        uint32_t e = ax >> 23;
        uint32_t man = (ax & 0x7FFFFFu) | 0x800000u;
        uint32_t shift = 126u - e;                 /* 14..24 */
        uint32_t hm = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1u);
        uint32_t half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (hm & 1u))) ++hm;
        return (uint16_t)(sign | hm);
    }
    uint32_t e = (ax >> 23) - 112u;
    uint32_t man = ax & 0x7FFFFFu;
    uint32_t hm = (e << 10) | (man >> 13);
    uint32_t rem = man & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (hm & 1u))) ++hm;
    return (uint16_t)(sign | hm);
}

static inline uint16_t hadd(uint16_t a, uint16_t b) { return f2h(h2f(a) + h2f(b)); }

ORC_API void orc_f2h(const float* in, uint16_t* out, int64_t n) { for (int64_t i = 0; i < n; ++i) out[i] = f2h(in[i]); }
ORC_API void orc_h2f(const uint16_t* in, float* out, int64_t n) { for (int64_t i = 0; i < n; ++i) out[i] = h2f(in[i]); }

/* ------------------------------------------------------------------------------------------ */
/* small vector helpers                                                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;
static inline v3 v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
/* glm::dot: tmp = a*b; tmp.x + tmp.y + tmp.z (left to right) - used by the mesh stage */
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* Eigen's unrolled 3-element redux evaluates e0 + (e1 + e2) (pinned by tests/golden/ref_vectors.npz: p2r_out) */
static inline float edot3(v3 a, v3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
static inline v3 add3(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, float s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 div3(v3 a, float s) { return v3_make(a.x / s, a.y / s, a.z / s); }
static inline v3 cross3(v3 a, v3 b) { return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* Eigen normalized(): n / sqrt(squaredNorm) when > 0 */
static inline v3 normalize3(v3 a) { float z = edot3(a, a); return z > 0.0f ? div3(a, sqrtf(z)) : a; }
/* glm::normalize: v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x) */
static inline v3 glm_normalize3(v3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return mul3(a, inv); }
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); } /* T/.../common.h:245-248 */

/* ------------------------------------------------------------------------------------------ */
/* Morton code  (T/include/tiny-cuda-nn/common_device.h:338-362)                                */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
static inline uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}
ORC_API uint32_t orc_morton3D(uint32_t x, uint32_t y, uint32_t z) { return morton3D(x, y, z); }
ORC_API uint32_t orc_morton3D_invert(uint32_t x) { return morton3D_invert(x); }

/* ------------------------------------------------------------------------------------------ */
/* Low-discrepancy jitter  (S/ngp/random_val.cuh:163-294)                                       */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t reverse_bits(uint32_t x) {
    x = (((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1));
    x = (((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2));
    x = (((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4));
    x = (((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8));
    return ((x >> 16) | (x << 16));
}
static inline uint32_t laine_karras_permutation(uint32_t x, uint32_t seed) {
    x += seed;
    x ^= x * 0x6c50b47cu;
    x ^= x * 0xb82f1e52u;
    x ^= x * 0xc7afe638u;
    x ^= x * 0x8d22f6e6u;
    return x;
}
static inline uint32_t nested_uniform_scramble_base2(uint32_t x, uint32_t seed) {
    x = reverse_bits(x);
    x = laine_karras_permutation(x, seed);
    x = reverse_bits(x);
    return x;
}
static inline uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
/* sobol(index, 0): direction numbers of dimension 0 are 0x80000000 >> bit, i.e. a bit reversal */
static inline uint32_t sobol_dim0(uint32_t index) { return reverse_bits(index); }
static inline float ld_random_val(uint32_t index, uint32_t seed) {
    const float S = 2.3283064365386963e-10f; /* float(1.0/(1ull<<32)) */
    index = nested_uniform_scramble_base2(index, seed);
    return (float)nested_uniform_scramble_base2(sobol_dim0(index), hash_combine(seed, 0u)) * S;
}
ORC_API float orc_ld_random_val(uint32_t index, uint32_t seed) { return ld_random_val(index, seed); }

/* ------------------------------------------------------------------------------------------ */
/* Colour transfer  (S/ngp/ngp_common.cuh:125-146)                                              */
/* ------------------------------------------------------------------------------------------ */
static inline float linear_to_srgb(float linear) {
    if (linear < 0.0031308f) return 12.92f * linear;
    return 1.055f * powf(linear, 0.41666f) - 0.055f;
}
static inline float srgb_to_linear(float srgb) {
    if (srgb <= 0.04045f) return srgb / 12.92f;
    return powf((srgb + 0.055f) / 1.055f, 2.4f);
}
ORC_API float orc_linear_to_srgb(float x) { return linear_to_srgb(x); }
ORC_API float orc_srgb_to_linear(float x) { return srgb_to_linear(x); }

/* ------------------------------------------------------------------------------------------ */
/* Bounding box  (S/ngp/bounding_box.cuh:106-167; translation member is always zero here)       */
/* ------------------------------------------------------------------------------------------ */
typedef struct { v3 min, max; } aabb_t;

static inline void swapf(float* a, float* b) { float c = *a; *a = *b; *b = c; }

static void aabb_ray_intersect(const aabb_t* b, v3 pos, v3 dir, float* tmin_out, float* tmax_out) {
    const float FMAX = 3.402823466e+38f;
    float tmin = (b->min.x - pos.x) / dir.x;
    float tmax = (b->max.x - pos.x) / dir.x;
    if (tmin > tmax) swapf(&tmin, &tmax);
    float tymin = (b->min.y - pos.y) / dir.y;
    float tymax = (b->max.y - pos.y) / dir.y;
    if (tymin > tymax) swapf(&tymin, &tymax);
    if (tmin > tymax || tymin > tmax) { *tmin_out = FMAX; *tmax_out = FMAX; return; }
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (b->min.z - pos.z) / dir.z;
    float tzmax = (b->max.z - pos.z) / dir.z;
    if (tzmin > tzmax) swapf(&tzmin, &tzmax);
    if (tmin > tzmax || tzmin > tmax) { *tmin_out = FMAX; *tmax_out = FMAX; return; }
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    *tmin_out = tmin; *tmax_out = tmax;
}
static inline int aabb_contains(const aabb_t* b, v3 p) {
    return p.x >= b->min.x && p.x <= b->max.x && p.y >= b->min.y && p.y <= b->max.y && p.z >= b->min.z && p.z <= b->max.z;
}
ORC_API void orc_aabb_ray_intersect(const float* bmin, const float* bmax, const float* pos, const float* dir, float* out2) {
    aabb_t b = {v3_make(bmin[0], bmin[1], bmin[2]), v3_make(bmax[0], bmax[1], bmax[2])};
    aabb_ray_intersect(&b, v3_make(pos[0], pos[1], pos[2]), v3_make(dir[0], dir[1], dir[2]), &out2[0], &out2[1]);
}

/* ------------------------------------------------------------------------------------------ */
/* Camera  (S/orbit_camera.h:7-77, R/dependencies/flythrough_camera.h:256-333,                  */
/*          S/nerf_mesh_renderer.cu:919-939)                                                    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    float view[16];
    float eye[3];
    float look[3];
    float pivot[3];
    float up[3];
} orc_camera;

static void look_to(const float eye[3], const float look[3], const float up[3], float view[16]) {
    float look_len = sqrtf(look[0] * look[0] + look[1] * look[1] + look[2] * look[2]);
    float up_len = sqrtf(up[0] * up[0] + up[1] * up[1] + up[2] * up[2]);
    float up_norm[3] = { up[0] / up_len, up[1] / up_len, up[2] / up_len };
    float f[3] = { look[0] / look_len, look[1] / look_len, look[2] / look_len };
    float s[3] = {
        f[1] * up_norm[2] - f[2] * up_norm[1],
        f[2] * up_norm[0] - f[0] * up_norm[2],
        f[0] * up_norm[1] - f[1] * up_norm[0]
    };
    float s_len = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    s[0] /= s_len; s[1] /= s_len; s[2] /= s_len;
    float u[3] = {
        s[1] * f[2] - s[2] * f[1],
        s[2] * f[0] - s[0] * f[2],
        s[0] * f[1] - s[1] * f[0]
    };
    float u_len = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    u[0] /= u_len; u[1] /= u_len; u[2] /= u_len;
    f[0] = -f[0]; f[1] = -f[1]; f[2] = -f[2];          /* right-handed */
    float t[3] = {
        s[0] * -eye[0] + s[1] * -eye[1] + s[2] * -eye[2],
        u[0] * -eye[0] + u[1] * -eye[1] + u[2] * -eye[2],
        f[0] * -eye[0] + f[1] * -eye[1] + f[2] * -eye[2]
    };
    view[0] = s[0]; view[1] = u[0]; view[2] = f[0]; view[3] = 0.0f;
    view[4] = s[1]; view[5] = u[1]; view[6] = f[1]; view[7] = 0.0f;
    view[8] = s[2]; view[9] = u[2]; view[10] = f[2]; view[11] = 0.0f;
    view[12] = t[0]; view[13] = t[1]; view[14] = t[2]; view[15] = 1.0f;
}

/* NerfMeshRenderer ctor state: S/nerf_mesh_renderer.cuh:90-95, S/nerf_mesh_renderer.cu:365-372 */
ORC_API void orc_camera_init(orc_camera* c) {
    c->eye[0] = 0.f; c->eye[1] = 0.f; c->eye[2] = 2.f;
    c->look[0] = 0.f; c->look[1] = -0.000001f; c->look[2] = -0.999999f;
    c->up[0] = 0.f; c->up[1] = 1.f; c->up[2] = 0.f;
    c->pivot[0] = c->pivot[1] = c->pivot[2] = 0.f;
    look_to(c->eye, c->look, c->up, c->view);
}

/* orbitcam(): note NerfMeshRenderer::orbit passes (delta_polar, delta_azimuth, delta_zoom) */
ORC_API void orc_camera_orbit(orc_camera* c, float delta_azimuth, float delta_polar, float delta_scroll) {
    const double PI_D = 3.14159265359; /* the reference's #define M_PI, a double literal */
    float* eye = c->eye; float* pivot = c->pivot;
    float dir[3] = { eye[0] - pivot[0], eye[1] - pivot[1], eye[2] - pivot[2] };
    float radius = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    { float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]); dir[0] /= len; dir[1] /= len; dir[2] /= len; }
    float azimuth = atan2f(dir[2], dir[0]);
    float polar = atan2f(dir[1], sqrtf(dir[0] * dir[0] + dir[2] * dir[2]));
    azimuth += delta_azimuth;
    azimuth = fmodf(azimuth, (float)(2 * PI_D));
    if (azimuth < 0.f) azimuth = (float)(azimuth + 2 * PI_D);
    polar += delta_polar;
    const float polarCap = (float)(PI_D / 2.f - 0.001f);
    polar = fminf(polarCap, fmaxf(-polarCap, polar));
    radius -= delta_scroll * radius * 0.1f;
    if (radius < 1.f) radius = 1.f;
    const float sa = sinf(azimuth), ca = cosf(azimuth), sp = sinf(polar), cp = cosf(polar);
    eye[0] = pivot[0] + radius * cp * ca;
    eye[1] = pivot[1] + radius * sp;
    eye[2] = pivot[2] + radius * cp * sa;
    c->look[0] = pivot[0] - eye[0];
    c->look[1] = pivot[1] - eye[1];
    c->look[2] = pivot[2] - eye[2];
    look_to(eye, c->look, c->up, c->view);
}

/* One pose of the GUI's trajectory tool (S/nerf_mesh_renderer.cu:649-658): cam_pos on a circle, cam_look = glm::normalize(lookat -
 * cam_pos) = v * (1 / sqrt(x*x + y*y + z*z)) (glm func_geometric.inl:82-90, 48-55), flythrough_camera_look_to. */
ORC_API void orc_camera_trajectory_pose(orc_camera* c, float angle, float distance, float height, const float* lookat3) {
    c->eye[0] = cosf(angle) * distance;
    c->eye[1] = height;
    c->eye[2] = sinf(angle) * distance;
    const float d[3] = { lookat3[0] - c->eye[0], lookat3[1] - c->eye[1], lookat3[2] - c->eye[2] };
    const float inv = 1.0f / sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    c->look[0] = d[0] * inv; c->look[1] = d[1] * inv; c->look[2] = d[2] * inv;
    look_to(c->eye, c->look, c->up, c->view);
}

/* updateModelViewProj(): 3x4 camera, stored column-major in out12 (col0,col1,col2,col3) */
ORC_API void orc_camera_matrix(const orc_camera* c, int screen_w, int screen_h, float* out12) {
    float aspect = (float)(uint32_t)screen_w / (float)(uint32_t)screen_h;
    float vLength = tanf(0.5f * 45);          /* radians, as in the reference */
    float uLength = vLength * aspect;
    const float* vm = c->view;
    out12[0] = vm[0] * uLength; out12[1] = vm[4] * uLength; out12[2] = vm[8] * uLength;
    out12[3] = vm[1] * vLength; out12[4] = vm[5] * vLength; out12[5] = vm[9] * vLength;
    out12[6] = -vm[2]; out12[7] = -vm[6]; out12[8] = -vm[10];
    out12[9] = c->eye[0]; out12[10] = c->eye[1]; out12[11] = c->eye[2];
}

/* ------------------------------------------------------------------------------------------ */
/* Model                                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int n_levels, log2_hashmap_size, base_resolution, hash_type; /* 0 Prime, 1 CoherentPrime, 2 ReversedPrime */
    float per_level_scale;
    int aabb_scale, max_cascade;
    float cone_angle_constant;
    uint32_t offsets[MAX_LEVELS + 1];
    float scales[MAX_LEVELS];
    uint32_t res[MAX_LEVELS];
    uint32_t stride_y[MAX_LEVELS], stride_z[MAX_LEVELS];
    int dense[MAX_LEVELS];
    int density_hidden, rgb_hidden, width, enc_width;
    uint64_t n_params;
    uint16_t* params;         /* owned copy, params_binary order */
    const uint16_t* grid;     /* -> params + mlp params */
    /* fp32 transposed weight copies: wt[layer][k * n_out + j] */
    int n_dlayers, n_rlayers;
    float* dW[8]; int dIn[8], dOut[8];
    float* rW[8]; int rIn[8], rOut[8];
    uint8_t* bitfield;        /* NERF_CASCADES * 128^3 / 8 bytes */
    float density_mean;
} orc_model;

static uint32_t next_multiple_u32(uint32_t v, uint32_t d) { return ((v + d - 1) / d) * d; }

/* grid_scale / grid_resolution: T/.../grid.h:196-205 */
static float grid_scale(uint32_t level, float log2_per_level_scale, uint32_t base_resolution) {
    return exp2f(level * log2_per_level_scale) * base_resolution - 1.0f;
}
static uint32_t grid_resolution(float scale) { return (uint32_t)ceilf(scale) + 1; }

static float* transpose_to_f32(const uint16_t* w, int n_out, int n_in) {
    float* t = (float*)malloc(sizeof(float) * (size_t)n_out * n_in);
    for (int j = 0; j < n_out; ++j) for (int k = 0; k < n_in; ++k) t[(size_t)k * n_out + j] = h2f(w[(size_t)j * n_in + k]);
    return t;
}

/* Parameter order: density net | rgb net | hash grid  (S/ngp/nerf_network.cuh:359-392);
 * per network: first layer [width x in], hidden [width x width]..., last [16 x width], row-major [out][in]
 * (T/src/fully_fused_mlp.cu:661-679).  Level offsets: T/.../grid.h:985-1016.
 * per_level_scale <= 0 selects the automatic value of S/ngp/testbed.cu:1197-1204. */
ORC_API orc_model* orc_model_create(const uint16_t* params, uint64_t n_params, int n_levels, int log2_hashmap_size,
                                    int base_resolution, float per_level_scale, int aabb_scale, int hash_type,
                                    int density_hidden, int rgb_hidden) {
    if (n_levels < 1 || n_levels > MAX_LEVELS || density_hidden < 1 || rgb_hidden < 1 || density_hidden > 6 || rgb_hidden > 6) return NULL;
    orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
    m->n_levels = n_levels; m->log2_hashmap_size = log2_hashmap_size; m->base_resolution = base_resolution;
    m->hash_type = hash_type; m->aabb_scale = aabb_scale; m->width = 64; m->enc_width = n_levels * 2;
    m->density_hidden = density_hidden; m->rgb_hidden = rgb_hidden;
    if (per_level_scale <= 0.0f && n_levels > 1) {
        per_level_scale = expf(logf(2048.0f * (float)aabb_scale / (float)base_resolution) / (n_levels - 1));
    }
    m->per_level_scale = per_level_scale;
    m->max_cascade = 0;
    while ((1 << m->max_cascade) < aabb_scale) ++m->max_cascade;
    m->cone_angle_constant = aabb_scale <= 1 ? 0.0f : (1.0f / 256.0f);   /* S/ngp/testbed.cu:1115 */
    float log2_pls = log2f(per_level_scale);
    uint32_t offset = 0;
    for (int i = 0; i < n_levels; ++i) {
        float scale = grid_scale((uint32_t)i, log2_pls, (uint32_t)base_resolution);
        uint32_t resolution = grid_resolution(scale);
        uint32_t max_params = 0xFFFFFFFFu / 2;
        uint32_t params_in_level = powf((float)resolution, 3) > (float)max_params ? max_params : resolution * resolution * resolution;
        params_in_level = next_multiple_u32(params_in_level, 8u);
        uint32_t cap = 1u << log2_hashmap_size;
        if (params_in_level > cap) params_in_level = cap;
        m->offsets[i] = offset;
        offset += params_in_level;
        m->scales[i] = scale;
        m->res[i] = resolution;
        /* grid_index(): uint32 stride loop and the dense/hash decision, T/.../grid.h:164-186 */
        uint32_t stride = 1, st[3] = {0, 0, 0};
        for (int d = 0; d < 3 && stride <= params_in_level; ++d) { st[d] = stride; stride *= resolution; }
        m->dense[i] = !(params_in_level < stride);
        m->stride_y[i] = st[1]; m->stride_z[i] = st[2];
    }
    m->offsets[n_levels] = offset;
    uint64_t mlp = 0;
    int W = m->width;
    int din[8], dout[8], nd = 0;
    din[nd] = m->enc_width; dout[nd] = W; ++nd;
    for (int i = 0; i + 1 < density_hidden; ++i) { din[nd] = W; dout[nd] = W; ++nd; }
    din[nd] = W; dout[nd] = 16; ++nd;
    int rin[8], rout[8], nr = 0;
    rin[nr] = 32; rout[nr] = W; ++nr;
    for (int i = 0; i + 1 < rgb_hidden; ++i) { rin[nr] = W; rout[nr] = W; ++nr; }
    rin[nr] = W; rout[nr] = 16; ++nr;
    for (int i = 0; i < nd; ++i) mlp += (uint64_t)din[i] * dout[i];
    for (int i = 0; i < nr; ++i) mlp += (uint64_t)rin[i] * rout[i];
    uint64_t expect = mlp + (uint64_t)offset * 2;
    if (n_params != expect) { free(m); return NULL; }
    m->n_params = n_params;
    m->params = (uint16_t*)malloc(n_params * 2);
    memcpy(m->params, params, n_params * 2);
    const uint16_t* p = m->params;
    m->n_dlayers = nd; m->n_rlayers = nr;
    for (int i = 0; i < nd; ++i) { m->dIn[i] = din[i]; m->dOut[i] = dout[i]; m->dW[i] = transpose_to_f32(p, dout[i], din[i]); p += din[i] * dout[i]; }
    for (int i = 0; i < nr; ++i) { m->rIn[i] = rin[i]; m->rOut[i] = rout[i]; m->rW[i] = transpose_to_f32(p, rout[i], rin[i]); p += rin[i] * rout[i]; }
    m->grid = p;
    m->bitfield = (uint8_t*)calloc(NERF_CASCADES * GRID_CELLS / 8, 1);
    return m;
}

ORC_API void orc_model_destroy(orc_model* m) {
    if (!m) return;
    for (int i = 0; i < m->n_dlayers; ++i) free(m->dW[i]);
    for (int i = 0; i < m->n_rlayers; ++i) free(m->rW[i]);
    free(m->params); free(m->bitfield); free(m);
}

ORC_API void orc_model_level_table(const orc_model* m, uint32_t* offsets17, float* scales16, uint32_t* res16, int* dense16) {
    for (int i = 0; i <= m->n_levels; ++i) offsets17[i] = m->offsets[i];
    for (int i = 0; i < m->n_levels; ++i) { scales16[i] = m->scales[i]; res16[i] = m->res[i]; dense16[i] = m->dense[i]; }
}
ORC_API float orc_model_per_level_scale(const orc_model* m) { return m->per_level_scale; }
ORC_API float orc_model_cone_angle(const orc_model* m) { return m->cone_angle_constant; }
ORC_API uint8_t* orc_model_bitfield(orc_model* m) { return m->bitfield; }
ORC_API float orc_model_density_mean(const orc_model* m) { return m->density_mean; }

/* update_density_grid_mean_and_bitfield + grid_to_bitfield + bitfield_max_pool
 * (S/ngp/testbed.cu:119-166, 1120-1135).  grid: fp16, 128^3 * (max_cascade+1), Morton order.
 * The mean only covers cascade 0 (reduce_sum over n_elements = 128^3). */
ORC_API int orc_model_set_density_grid(orc_model* m, const uint16_t* grid, uint64_t n) {
    if (n == 0) { memset(m->bitfield, 0, NERF_CASCADES * GRID_CELLS / 8); return 0; }
    if (n != (uint64_t)GRID_CELLS * (uint64_t)(m->max_cascade + 1)) return -1;
    double sum = 0.0;
    for (uint32_t i = 0; i < GRID_CELLS; ++i) { float v = h2f(grid[i]); sum += (double)(fmaxf(v, 0.f) / (float)GRID_CELLS); }
    m->density_mean = (float)sum;
    float thresh = fminf(0.01f, m->density_mean);
    uint32_t n_bytes = GRID_CELLS / 8 * NERF_CASCADES, n_nonzero = GRID_CELLS / 8 * (uint32_t)(m->max_cascade + 1);
    for (uint32_t i = 0; i < n_bytes; ++i) {
        uint8_t bits = 0;
        if (i < n_nonzero) for (uint32_t j = 0; j < 8; ++j) bits |= h2f(grid[(uint64_t)i * 8 + j]) > thresh ? (uint8_t)(1u << j) : 0;
        m->bitfield[i] = bits;
    }
    for (uint32_t level = 1; level < NERF_CASCADES; ++level) {
        const uint8_t* prev = m->bitfield + (size_t)GRID_CELLS / 8 * (level - 1);
        uint8_t* next = m->bitfield + (size_t)GRID_CELLS / 8 * level;
        for (uint32_t i = 0; i < GRID_CELLS / 64; ++i) {
            uint8_t bits = 0;
            for (uint32_t j = 0; j < 8; ++j) bits |= prev[i * 8 + j] > 0 ? (uint8_t)(1u << j) : 0;
            uint32_t x = morton3D_invert(i >> 0) + NERF_GRIDSIZE / 8;
            uint32_t y = morton3D_invert(i >> 1) + NERF_GRIDSIZE / 8;
            uint32_t z = morton3D_invert(i >> 2) + NERF_GRIDSIZE / 8;
            next[morton3D(x, y, z)] |= bits;
        }
    }
    return 0;
}

ORC_API void orc_model_set_bitfield(orc_model* m, const uint8_t* bits) { memcpy(m->bitfield, bits, NERF_CASCADES * GRID_CELLS / 8); }

/* ------------------------------------------------------------------------------------------ */
/* Occupancy-grid DDA helpers  (S/ngp/testbed.cu:177-264, 293-315)                              */
/* ------------------------------------------------------------------------------------------ */
#define SQRT3 1.73205080757f
static inline float STEPSIZE(void) { return SQRT3 / 1024.0f; }
static inline float MIN_CONE_STEPSIZE(void) { return STEPSIZE(); }
static inline float MAX_CONE_STEPSIZE(void) { return STEPSIZE() * (float)(1 << (NERF_CASCADES - 1)) * 1024.0f / 128.0f; }
static inline float calc_dt(float t, float cone_angle) { return clampf(t * cone_angle, MIN_CONE_STEPSIZE(), MAX_CONE_STEPSIZE()); }

static inline int mip_from_pos(v3 pos) {
    int exponent;
    float maxval = fmaxf(fmaxf(fabsf(pos.x - 0.5f), fabsf(pos.y - 0.5f)), fabsf(pos.z - 0.5f));
    frexpf(maxval, &exponent);
    int v = exponent + 1; if (v < 0) v = 0;
    return v < (int)(NERF_CASCADES - 1) ? v : (int)(NERF_CASCADES - 1);
}
static inline int mip_from_dt(float dt, v3 pos) {
    int mip = mip_from_pos(pos);
    dt *= 2 * NERF_GRIDSIZE;
    if (dt < 1.f) return mip;
    int exponent;
    frexpf(dt, &exponent);
    int v = exponent > mip ? exponent : mip;
    return v < (int)(NERF_CASCADES - 1) ? v : (int)(NERF_CASCADES - 1);
}
static inline uint32_t cascaded_grid_idx_at(v3 pos, uint32_t mip) {
    float mip_scale = scalbnf(1.0f, -(int)mip);
    pos.x -= 0.5f; pos.y -= 0.5f; pos.z -= 0.5f;
    pos.x *= mip_scale; pos.y *= mip_scale; pos.z *= mip_scale;
    pos.x += 0.5f; pos.y += 0.5f; pos.z += 0.5f;
    int ix = (int)(pos.x * (float)NERF_GRIDSIZE), iy = (int)(pos.y * (float)NERF_GRIDSIZE), iz = (int)(pos.z * (float)NERF_GRIDSIZE);
    ix = ix < 0 ? 0 : (ix > 127 ? 127 : ix);
    iy = iy < 0 ? 0 : (iy > 127 ? 127 : iy);
    iz = iz < 0 ? 0 : (iz > 127 ? 127 : iz);
    return morton3D((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
}
static inline int density_grid_occupied_at(v3 pos, const uint8_t* bitfield, uint32_t mip) {
    uint32_t idx = cascaded_grid_idx_at(pos, mip);
    return bitfield[idx / 8 + (GRID_CELLS * mip) / 8] & (1u << (idx % 8));
}
static inline float signf1(float x) { return copysignf(1.0f, x); }
static inline float distance_to_next_voxel(v3 pos, v3 dir, v3 idir, uint32_t res) {
    float r = (float)res;
    v3 p = v3_make(r * pos.x, r * pos.y, r * pos.z);
    float tx = (floorf(p.x + 0.5f + 0.5f * signf1(dir.x)) - p.x) * idir.x;
    float ty = (floorf(p.y + 0.5f + 0.5f * signf1(dir.y)) - p.y) * idir.y;
    float tz = (floorf(p.z + 0.5f + 0.5f * signf1(dir.z)) - p.z) * idir.z;
    float t = fminf(fminf(tx, ty), tz);
    return fmaxf(t / r, 0.0f);
}
static inline float advance_to_next_voxel(float t, float cone_angle, v3 pos, v3 dir, v3 idir, uint32_t res) {
    float t_target = t + distance_to_next_voxel(pos, dir, idir, res);
    do { t += calc_dt(t, cone_angle); } while (t < t_target);
    return t;
}
static inline float warp_dt(float dt) {
    float max_stepsize = MIN_CONE_STEPSIZE() * (float)(1 << (NERF_CASCADES - 1));
    return (dt - MIN_CONE_STEPSIZE()) / (max_stepsize - MIN_CONE_STEPSIZE());
}
static inline float unwarp_dt(float dt) {
    float max_stepsize = MIN_CONE_STEPSIZE() * (float)(1 << (NERF_CASCADES - 1));
    return dt * (max_stepsize - MIN_CONE_STEPSIZE()) + MIN_CONE_STEPSIZE();
}
ORC_API uint32_t orc_cascaded_grid_idx_at(const float* pos, uint32_t mip) { return cascaded_grid_idx_at(v3_make(pos[0], pos[1], pos[2]), mip); }
ORC_API int orc_mip_from_pos(const float* pos) { return mip_from_pos(v3_make(pos[0], pos[1], pos[2])); }
ORC_API float orc_calc_dt(float t, float cone) { return calc_dt(t, cone); }

/* ------------------------------------------------------------------------------------------ */
/* Hash-grid encoding  (T/.../grid.h:111-186, 219-349; T/.../common_device.h:419-431)           */
/* ------------------------------------------------------------------------------------------ */
static inline uint32_t grid_hash(const orc_model* m, uint32_t x, uint32_t y, uint32_t z) {
    switch (m->hash_type) {
        case 0: return (x * 1958374283u) ^ (y * 2654435761u) ^ (z * 805459861u);       /* Prime */
        case 2: return (x * 2165219737u) ^ (y * 1434869437u) ^ (z * 2097192037u);      /* ReversedPrime */
        default: return (x * 1u) ^ (y * 2654435761u) ^ (z * 805459861u);               /* CoherentPrime */
    }
}

/* pos in [0,1]^3 -> enc[2*n_levels] fp16 bit patterns */
static void encode_position(const orc_model* m, v3 pos, uint16_t* enc) {
    const float in[3] = { pos.x, pos.y, pos.z };
    for (int level = 0; level < m->n_levels; ++level) {
        const uint16_t* grid = m->grid + (size_t)m->offsets[level] * 2;
        const uint32_t hashmap_size = m->offsets[level + 1] - m->offsets[level];
        const float scale = m->scales[level];
        float p[3]; uint32_t pg[3];
        for (int d = 0; d < 3; ++d) {
            float v = in[d] * scale + 0.5f;          /* two roundings: no fma */
            int tmp = (int)floorf(v);
            pg[d] = (uint32_t)tmp;
            p[d] = v - (float)tmp;
        }
        uint16_t r0 = 0, r1 = 0;
        for (uint32_t idx = 0; idx < 8; ++idx) {
            float weight = 1;
            uint32_t c[3];
            for (int d = 0; d < 3; ++d) {
                if ((idx & (1u << d)) == 0) { weight *= 1 - p[d]; c[d] = pg[d]; }
                else { weight *= p[d]; c[d] = pg[d] + 1; }
            }
            uint32_t index;
            if (m->dense[level]) index = c[0] + c[1] * m->stride_y[level] + c[2] * m->stride_z[level];
            else index = grid_hash(m, c[0], c[1], c[2]);
            index = (index % hashmap_size) * 2;
            float d0 = h2f(grid[index]), d1 = h2f(grid[index + 1]);
            r0 = hadd(r0, f2h(weight * d0));
            r1 = hadd(r1, f2h(weight * d1));
        }
        enc[level * 2 + 0] = r0;
        enc[level * 2 + 1] = r1;
    }
}

ORC_API void orc_encode(const orc_model* m, const float* pos, int64_t n, uint16_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) encode_position(m, v3_make(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]), out + i * m->enc_width);
}

/* ------------------------------------------------------------------------------------------ */
/* SH degree 4  (T/.../spherical_harmonics.h:65-98)                                             */
/* ------------------------------------------------------------------------------------------ */
static void sh4(v3 d01, uint16_t* out16) {
    float x = d01.x * 2.f - 1.f, y = d01.y * 2.f - 1.f, z = d01.z * 2.f - 1.f;
    float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    float o[16];
    o[0] = 0.28209479177387814f;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    for (int i = 0; i < 16; ++i) out16[i] = f2h(o[i]);
}
ORC_API void orc_sh4(const float* dir01, int64_t n, uint16_t* out) {
    for (int64_t i = 0; i < n; ++i) sh4(v3_make(dir01[i * 3], dir01[i * 3 + 1], dir01[i * 3 + 2]), out + i * 16);
}

/* ------------------------------------------------------------------------------------------ */
/* Fully fused MLP semantics  (T/src/fully_fused_mlp.cu:47-129, 316-476, 499-557):              */
/* y = act(W x) per layer, fp16 in/out, ReLU on hidden layers, no output activation.            */
/* ------------------------------------------------------------------------------------------ */
static void mlp_forward(int n_layers, float* const* W, const int* nin, const int* nout, const uint16_t* x16, uint16_t* y16) {
    float h[64], acc[64];
    for (int k = 0; k < nin[0]; ++k) h[k] = h2f(x16[k]);
    for (int l = 0; l < n_layers; ++l) {
        const int ni = nin[l], no = nout[l];
        const float* w = W[l];
        for (int j = 0; j < no; ++j) acc[j] = 0.0f;
        for (int k = 0; k < ni; ++k) {
            const float xk = h[k];
            const float* wk = w + (size_t)k * no;
            for (int j = 0; j < no; ++j) acc[j] = acc[j] + wk[j] * xk;
        }
        const int last = (l + 1 == n_layers);
        for (int j = 0; j < no; ++j) {
            float v = acc[j];
            if (!last) v = v > 0.0f ? v : 0.0f;
            uint16_t hv = f2h(v);
            if (last) y16[j] = hv; else h[j] = h2f(hv);
        }
    }
}

/* NerfNetwork::inference_mixed_precision_impl (S/ngp/nerf_network.cuh:101-135):
 * enc -> density net (16 out) ; [density out | SH16] -> rgb net ; out = (r,g,b raw, density raw ch 0). */
static void network_eval(const orc_model* m, v3 pos_warped, v3 dir01, uint16_t out4[4], uint16_t* enc_dbg) {
    uint16_t enc[2 * MAX_LEVELS];
    uint16_t rgb_in[32];
    uint16_t rgb_out[16];
    encode_position(m, pos_warped, enc);
    if (enc_dbg) memcpy(enc_dbg, enc, sizeof(uint16_t) * m->enc_width);
    mlp_forward(m->n_dlayers, m->dW, m->dIn, m->dOut, enc, rgb_in);
    sh4(dir01, rgb_in + 16);
    mlp_forward(m->n_rlayers, m->rW, m->rIn, m->rOut, rgb_in, rgb_out);
    out4[0] = rgb_out[0]; out4[1] = rgb_out[1]; out4[2] = rgb_out[2]; out4[3] = rgb_in[0];
}

ORC_API void orc_network(const orc_model* m, const float* pos, const float* dir01, int64_t n, uint16_t* out4) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        network_eval(m, v3_make(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]), v3_make(dir01[i * 3], dir01[i * 3 + 1], dir01[i * 3 + 2]), out4 + i * 4, NULL);
}

/* generic MLP entry for unit tests: which = 0 density net, 1 rgb net */
ORC_API void orc_mlp(const orc_model* m, int which, const uint16_t* x, int64_t n, uint16_t* y) {
    int ni = which ? 32 : m->enc_width;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        if (which) mlp_forward(m->n_rlayers, m->rW, m->rIn, m->rOut, x + i * ni, y + i * 16);
        else mlp_forward(m->n_dlayers, m->dW, m->dIn, m->dOut, x + i * ni, y + i * 16);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Render                                                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t width, height;
    float camera[12];             /* 3x4, column-major: col0 = right*uLen, col1 = up*vLen, col2 = fwd, col3 = eye */
    float aabb_min[3], aabb_max[3];               /* render aabb (crop box) */
    float render_aabb_to_local[9];                /* row-major 3x3 */
    float train_aabb_min[3], train_aabb_max[3];   /* m_aabb */
    float cone_angle;
    uint32_t spp_index;           /* CudaRenderBuffer::spp() of the frame being rendered */
    float min_transmittance;      /* 0.01 */
    int32_t rgb_activation;       /* 0 None 1 ReLU 2 Logistic 3 Exponential */
    int32_t density_activation;
    int32_t n_steps_mode;         /* 0: one sample per wavefront iteration; 1: the reference's clamp(N0/N_alive,1,8); 2: always 8 (what 1 gives
                                     while at most 1/8 of the FRAME's pixels are live - for windows of a larger frame) */
    int32_t x0, y0, x1, y1;       /* pixel window [x0,x1) x [y0,y1) to render (others untouched); all 0 = full frame */
    /* Testbed::m_model_rotation / m_model_translation as the model matrix of S/ngp/testbed.cu:1537-1542: rotation row-major 3x3
     * (all zeros = identity), translation in world units */
    float model_rot[9], model_trans[3];
} orc_render_params;

typedef struct {
    v3 origin, dir;
    float surf[4];
    float t_surface, t, t_start, max_weight;
    uint32_t idx;
    uint32_t n_steps;
    int alive;
    float rgba[4];
    float depth;
    uint32_t n_samples;           /* network evaluations consumed (composited or discarded) */
    int saturated;                /* the last composite call ended on the transmittance / surface-opacity threshold */
    int lens;                     /* lens ray (handled by march_lens_ray, not by the wavefront loop) */
} ray_t;

static inline float act_density(float v, int a) {
    switch (a) { case 0: return v; case 1: return v > 0.f ? v : 0.f; case 2: return 1.0f / (1.0f + expf(-v)); default: return expf(v); }
}
static inline float act_rgb(float v, int a) {
    switch (a) { case 0: return v; case 1: return v > 0.f ? v : 0.f; case 2: return 1.0f / (1.0f + expf(-v)); default: return expf(clampf(v, -10.f, 10.f)); }
}

static inline v3 mat3_mul(const float* m /* row-major */, v3 p) {
    return v3_make(m[0] * p.x + (m[1] * p.y + m[2] * p.z), m[3] * p.x + (m[4] * p.y + m[5] * p.z), m[6] * p.x + (m[7] * p.y + m[8] * p.z));
}
/* glm mat3 * vec3: left-to-right sums */
static inline v3 glm_mat3_mul(const float* m /* row-major */, v3 p) {
    return v3_make((m[0] * p.x + m[1] * p.y) + m[2] * p.z, (m[3] * p.x + m[4] * p.y) + m[5] * p.z, (m[6] * p.x + m[7] * p.y) + m[8] * p.z);
}

/* pixel_to_ray (S/ngp/ngp_common.cuh:334-394, perspective branch) + init_rays_with_payload_kernel_nerf
 * (S/ngp/testbed.cu:355-467).  Model matrix [R | t_model] (S/ngp/testbed.cu:1537-1542, consumed at :442-446):
 * dir = R d, NeRF-space origin = R eye + 0.5 + R t_model; the identity gives eye + 0.5. */
static void init_ray(const orc_render_params* P, const aabb_t* render_aabb, uint32_t x, uint32_t y, ray_t* r) {
    const float* c = P->camera;
    float ux = 2.0f * (((float)x + 0.5f) / (float)P->width) - 1.0f;
    float uy = 2.0f * (((float)y + 0.5f) / (float)P->height) - 1.0f;
    v3 d = v3_make(c[0] * ux + (c[3] * uy + c[6] * 1.0f), c[1] * ux + (c[4] * uy + c[7] * 1.0f), c[2] * ux + (c[5] * uy + c[8] * 1.0f));
    v3 o = v3_make(c[9], c[10], c[11]);
    d = normalize3(d);
    v3 to = v3_make(o.x + 0.5f, o.y + 0.5f, o.z + 0.5f);
    {
        const float* R = P->model_rot; int any = 0;
        for (int k = 0; k < 9; ++k) any |= (R[k] != 0.f);
        const int identity = !any || (R[0] == 1.f && R[4] == 1.f && R[8] == 1.f && R[1] == 0.f && R[2] == 0.f && R[3] == 0.f && R[5] == 0.f && R[6] == 0.f && R[7] == 0.f);
        const v3 tm = v3_make(P->model_trans[0], P->model_trans[1], P->model_trans[2]);
        if (!identity) {
            d = glm_mat3_mul(R, d);
            const v3 ro = glm_mat3_mul(R, o), rt = glm_mat3_mul(R, tm);
            to = v3_make((ro.x + 0.5f) + rt.x, (ro.y + 0.5f) + rt.y, (ro.z + 0.5f) + rt.z);
        } else if (tm.x != 0.f || tm.y != 0.f || tm.z != 0.f) {
            to = v3_make((o.x + 0.5f) + tm.x, (o.y + 0.5f) + tm.y, (o.z + 0.5f) + tm.z);
        }
    }
    float tmin, tmax;
    aabb_ray_intersect(render_aabb, to, d, &tmin, &tmax);
    float t = fmaxf(tmin, 0.0f) + 1e-6f;
    memset(r, 0, sizeof(*r));
    r->origin = to; r->dir = d; r->t = t; r->t_start = 0.f; r->idx = x + (uint32_t)P->width * y; r->n_steps = 0;
    r->max_weight = 0.f;
    r->alive = aabb_contains(render_aabb, add3(to, mul3(d, t)));
}

/* advance_pos_nerf (S/ngp/testbed.cu:470-537) */
static void advance_pos(const orc_model* m, const orc_render_params* P, const aabb_t* render_aabb, ray_t* r) {
    if (!r->alive) {
        if (r->t_surface != 0.0f) { r->t = r->t_surface; r->alive = 1; }
        return;
    }
    v3 origin = r->origin, dir = r->dir;
    v3 idir = v3_make(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    float cone_angle = P->cone_angle;
    float t = r->t;
    float dt = calc_dt(t, cone_angle);
    t += ld_random_val(P->spp_index, r->idx * 786433u) * dt;
    v3 pos;
    while (1) {
        if (r->t_surface != 0.0f && t > r->t_surface) { r->t = r->t_surface; return; }
        pos = add3(origin, mul3(dir, t));
        if (!aabb_contains(render_aabb, mat3_mul(P->render_aabb_to_local, pos))) {
            if (r->t_surface != 0.0f) { r->t = r->t_surface; return; }
            r->alive = 0;
            break;
        }
        dt = calc_dt(t, cone_angle);
        uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
        if (density_grid_occupied_at(pos, m->bitfield, mip)) break;
        uint32_t res = NERF_GRIDSIZE >> mip;
        t = advance_to_next_voxel(t, cone_angle, pos, dir, idir, res);
    }
    r->t = t;
    if (mip_from_pos(add3(origin, mul3(dir, t))) == 0) r->t_start = t;
}

typedef struct { v3 pos; float dt_warped; float t; uint32_t cell; uint32_t mip; } sample_t;

/* generate_next_nerf_network_inputs (S/ngp/testbed.cu:564-633); returns number of samples written */
/* t_stop (lens rays only, no reference equivalent): no sample beyond it; when the walk stops there or leaves the box, r->t
 * keeps the walk's position so that a later segment can resume it.  Pass INFINITY for the reference's behaviour. */
static uint32_t generate_samples_until(const orc_model* m, const orc_render_params* P, const aabb_t* render_aabb, const aabb_t* train_aabb,
                                       ray_t* r, uint32_t n_steps, sample_t* out, int ignore_surface, float t_stop, int keep_walk_state) {
    v3 origin = r->origin, dir = r->dir;
    v3 idir = v3_make(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    float cone_angle = P->cone_angle;
    float t = r->t;
    for (uint32_t j = 0; j < n_steps; ++j) {
        v3 pos; float dt = 0.0f; uint32_t mip = 0;
        while (1) {
            if (!ignore_surface && r->t_surface != 0.0f && t > r->t_surface && r->surf[3] == 1.f) {
                r->n_steps = j; r->t = r->t_surface; return j;
            }
            if (t > t_stop) { r->n_steps = j; r->t = t; return j; }
            pos = add3(origin, mul3(dir, t));
            if (!aabb_contains(render_aabb, mat3_mul(P->render_aabb_to_local, pos))) { r->n_steps = j; if (keep_walk_state) r->t = t; return j; }
            dt = calc_dt(t - r->t_start, cone_angle);
            mip = (uint32_t)mip_from_dt(dt, pos);
            if (density_grid_occupied_at(pos, m->bitfield, mip)) break;
            uint32_t res = NERF_GRIDSIZE >> mip;
            t = advance_to_next_voxel(t, cone_angle, pos, dir, idir, res);
        }
        /* warp_position = aabb.relative_pos (S/ngp/testbed.cu:204-208, bounding_box.cuh:83-85) */
        v3 diag = sub3(train_aabb->max, train_aabb->min);
        v3 rel = sub3(pos, train_aabb->min);
        out[j].pos = v3_make(rel.x / diag.x, rel.y / diag.y, rel.z / diag.z);
        out[j].dt_warped = warp_dt(dt);
        out[j].t = t;
        out[j].cell = cascaded_grid_idx_at(pos, mip);
        out[j].mip = mip;
        t += dt;
    }
    r->t = t;
    r->n_steps = n_steps;
    return n_steps;
}

static uint32_t generate_samples(const orc_model* m, const orc_render_params* P, const aabb_t* render_aabb, const aabb_t* train_aabb,
                                 ray_t* r, uint32_t n_steps, sample_t* out, int ignore_surface) {
    return generate_samples_until(m, P, render_aabb, train_aabb, r, n_steps, out, ignore_surface, INFINITY, 0);
}

/* composite_kernel_nerf (S/ngp/testbed.cu:784-905) for one ray and its batch of <= n_steps samples */
static void composite_ray(const orc_render_params* P, const aabb_t* train_aabb, ray_t* r, uint32_t n_steps,
                          const sample_t* smp, const uint16_t* net_out /* [n][4] */, uint32_t current_step) {
    float* c = r->rgba;
    float local_depth = r->depth;
    const v3 cam_origin = v3_make(P->camera[9], P->camera[10], P->camera[11]);
    uint32_t actual = r->n_steps, j = 0;
    r->saturated = 0;
    for (; j < actual; ++j) {
        float sr = h2f(net_out[j * 4 + 0]), sg = h2f(net_out[j * 4 + 1]), sb = h2f(net_out[j * 4 + 2]), sd = h2f(net_out[j * 4 + 3]);
        v3 diag = sub3(train_aabb->max, train_aabb->min);
        v3 pos = v3_make(train_aabb->min.x + smp[j].pos.x * diag.x, train_aabb->min.y + smp[j].pos.y * diag.y, train_aabb->min.z + smp[j].pos.z * diag.z);
        float T = 1.f - c[3];
        float dt = unwarp_dt(smp[j].dt_warped);
        if (r->t > r->t_surface && r->surf[3] > 0) {
            c[0] += r->surf[0] * r->surf[3] * T;
            c[1] += r->surf[1] * r->surf[3] * T;
            c[2] += r->surf[2] * r->surf[3] * T;
            c[3] += r->surf[3] * T;
            r->surf[3] = 0.f;
            T = 1.f - c[3];
            if (c[3] > 0.99f) {
                float a = c[3]; c[0] /= a; c[1] /= a; c[2] /= a; c[3] /= a;
                r->saturated = 1;
                break;
            }
        }
        float alpha = 1.f - expf(-act_density(sd, P->density_activation) * dt);
        float weight = alpha * T;
        c[0] += act_rgb(sr, P->rgb_activation) * weight;
        c[1] += act_rgb(sg, P->rgb_activation) * weight;
        c[2] += act_rgb(sb, P->rgb_activation) * weight;
        c[3] += weight;
        if (weight > r->max_weight) {
            r->max_weight = weight;
            v3 dd = sub3(pos, cam_origin);     /* mixes NeRF-space pos with world-space eye, as the reference does */
            local_depth = sqrtf(edot3(dd, dd));
        }
        if (c[3] > (1.0f - P->min_transmittance)) {
            float a = c[3]; c[0] /= a; c[1] /= a; c[2] /= a; c[3] /= a;
            r->saturated = 1;
            break;
        }
    }
    if (j < n_steps) {
        if (r->surf[3] > 0) {
            float k = 1.f - c[3];
            c[0] += r->surf[0] * k; c[1] += r->surf[1] * k; c[2] += r->surf[2] * k; c[3] += r->surf[3] * k;
        }
        r->alive = 0;
        r->n_steps = j + current_step;
    }
    r->depth = local_depth;
}

/* Renders one frame into the linear, premultiplied frame buffer exactly as Testbed::render_nerf leaves
 * it before accumulate/tonemap (S/ngp/testbed.cu:1521-1612, 1938-2053, 907-931).
 *   surf_rgba/t_surface: per-pixel mesh hand-off (A12) or NULL.
 *   frame[W*H*4], depth[W*H]: outputs (cleared here like clear_frame; depth gets 1e10 from init).
 *   n_samples[W*H] (optional): network evaluations per ray.  stats[4] (optional): {rays alive after
 *   first-hit, total samples, wavefront iterations, rays hit}. */
/* ---- lens rays (NEW functionality: the published reference traces primary rays only; this is the
 * executable specification of libnmr's secondary rays, DESIGN.md "Secondary rays") ---------------------------------- */
typedef struct { float f0, k[3], kmean, background[4]; int model; float thickness, ior; } lens_params_t;

/* Lens model 1 ("plate"): a pane of glass of `thickness` with parallel faces.  The ray refracts into the glass at the hit point
 * (Snell, relative index 1 / ior, normal turned towards the ray), crosses it, and leaves through the back face parallel to its old
 * direction: the transmitted segment is the primary ray shifted sideways by delta and resumed behind the pane.  Model 0 (thin
 * sheet): delta = 0.  The grazing-angle factor is capped (cos >= 0.05). */
static void lens_plate_shift(const lens_params_t* L, v3 dir, v3 n, float t_lens, v3* delta, float* t_behind) {
    *delta = v3_make(0.f, 0.f, 0.f); *t_behind = t_lens;
    if (L->model != 1 || !(L->thickness > 0.f)) return;
    float c = dot3(dir, n);
    if (c > 0.f) { n = mul3(n, -1.f); c = -c; }
    const float cosi = fmaxf(fminf(-c, 1.0f), 0.05f);
    const float eta = 1.0f / L->ior;
    const float cost = sqrtf(fmaxf(0.f, 1.0f - (eta * eta) * (1.0f - cosi * cosi)));
    const v3 td = add3(mul3(dir, eta), mul3(n, eta * cosi - cost));          /* unit direction inside the glass */
    const float d = L->thickness;
    *delta = sub3(mul3(td, d / cost), mul3(dir, d / cosi));
    *t_behind = t_lens + d / cosi;
}

/* Schlick reflectance of a thin lens for unit direction d and unit normal n (turned towards the ray); mirror direction out */
static float lens_fresnel(const lens_params_t* L, v3 d, v3 n, v3* refl) {
    float c = dot3(d, n);
    if (c > 0.f) { n = mul3(n, -1.f); c = -c; }
    float cosi = fminf(-c, 1.0f);
    *refl = glm_normalize3(sub3(d, mul3(n, 2.0f * c)));
    float mm = 1.0f - cosi, m2 = mm * mm;
    return L->f0 + (1.0f - L->f0) * (m2 * m2 * mm);
}

/* marches one segment to its end with batches of n samples; returns samples consumed */
static uint32_t march_segment(const orc_model* m, const orc_render_params* P, const aabb_t* render_aabb, const aabb_t* train_aabb,
                              ray_t* r, uint32_t n, float t_stop, int keep_walk_state) {
    uint32_t total = 0, step = 1;
    while (r->alive && step < 10000u) {
        sample_t smp[8]; uint16_t out[8 * 4];
        uint32_t cnt = generate_samples_until(m, P, render_aabb, train_aabb, r, n, smp, 0, t_stop, keep_walk_state);
        v3 dir01 = v3_make((r->dir.x + 1.0f) * 0.5f, (r->dir.y + 1.0f) * 0.5f, (r->dir.z + 1.0f) * 0.5f);
        for (uint32_t j = 0; j < cnt; ++j) network_eval(m, smp[j].pos, dir01, out + j * 4, NULL);
        total += cnt;
        composite_ray(P, train_aabb, r, n, smp, out, step);
        step += n;
    }
    return total;
}

/* One lens ray: r arrives after init_ray + advance_pos (clipped / revived at the lens like at a mesh surface).
 * Segment 1: samples in front of the lens (t <= t_lens), opaque mesh surface inactive.  If the ray saturates there it is an
 * ordinary ray.  Otherwise split: segment 2 = mirror ray from the hit point (NeRF only, over the background colour), taken when
 * coverage * F * transmittance >= 1/512; segment 3 = the primary ray carries on from where its walk stood, now with the opaque
 * surface (surf, t_surface).  Result = front + T_front * (w (F refl + (1 - F) k behind) + (1 - w) behind). */
static void march_lens_ray(const orc_model* m, const orc_render_params* P, const aabb_t* render_aabb, const aabb_t* train_aabb, const lens_params_t* L,
                           ray_t* r, uint32_t n, float w, float t_lens, v3 nrm, const float surf[4], float t_surf) {
    const v3 dir = r->dir, origin = r->origin;
    const float t_start = r->t_start;
    r->t_surface = 0.f; r->surf[0] = r->surf[1] = r->surf[2] = r->surf[3] = 0.f;
    const char* dbg_env = getenv("ORC_DEBUG_PIXEL");      /* diagnostics: the segments of one lens ray on stderr */
    const int dbg = dbg_env && (uint32_t)atoi(dbg_env) == r->idx;
    if (dbg) fprintf(stderr, "[oracle lens %u] n %u w %.4f t_lens %.7f t0 %.7f t_start %.7f t_surf %.7f surf_w %.3f\n", r->idx, n, w, t_lens, r->t, t_start, t_surf, surf[3]);
    r->n_samples += march_segment(m, P, render_aabb, train_aabb, r, n, t_lens, 1);   /* r->t must end where the walk stood: segment 3 resumes there */
    if (dbg) fprintf(stderr, "[oracle lens %u] segment 1: samples %u saturated %d alpha %.5f t_resume %.7f\n", r->idx, r->n_samples, r->saturated, r->rgba[3], r->t);
    if (r->saturated) return;
    float front[4]; memcpy(front, r->rgba, 16);
    const float t_resume = r->t;
    v3 refl;
    const float F = lens_fresnel(L, dir, nrm, &refl);
    float LB[3] = { 0.f, 0.f, 0.f };
    if (w * F * (1.f - front[3]) >= 1.0f / 512.0f) {
        ray_t b; memset(&b, 0, sizeof(b));
        b.origin = add3(origin, mul3(dir, t_lens)); b.dir = refl; b.t = 1e-3f; b.alive = 1; b.idx = r->idx;
        r->n_samples += march_segment(m, P, render_aabb, train_aabb, &b, n, INFINITY, 0);
        float k = (1.f - b.rgba[3]) * L->background[3];
        for (int c = 0; c < 3; ++c) LB[c] = b.rgba[c] + L->background[c] * k;
    }
    ray_t c3; memset(&c3, 0, sizeof(c3));
    v3 delta; float t_behind;
    lens_plate_shift(L, dir, nrm, t_lens, &delta, &t_behind);
    c3.origin = add3(origin, delta); c3.dir = dir; c3.t = fmaxf(t_resume, t_behind); c3.t_start = t_start; c3.t_surface = t_surf; memcpy(c3.surf, surf, 16); c3.alive = 1; c3.idx = r->idx;
    if (dbg) fprintf(stderr, "[oracle lens %u] F %.5f mirror segment taken %d, samples so far %u; segment 3 starts at t %.7f\n", r->idx, F, w * F * (1.f - front[3]) >= 1.0f / 512.0f, r->n_samples, c3.t);
    r->n_samples += march_segment(m, P, render_aabb, train_aabb, &c3, n, INFINITY, 0);   /* payload.t semantics of the reference (composite_ray) */
    if (dbg) fprintf(stderr, "[oracle lens %u] segment 3: samples total %u, behind rgba %.4f %.4f %.4f %.4f\n", r->idx, r->n_samples, c3.rgba[0], c3.rgba[1], c3.rgba[2], c3.rgba[3]);
    const float Tf = 1.f - front[3];
    for (int c = 0; c < 3; ++c) {
        float kc = (1.f - F) * L->k[c];
        r->rgba[c] = front[c] + Tf * (w * (F * LB[c] + kc * c3.rgba[c]) + (1.f - w) * c3.rgba[c]);
    }
    r->rgba[3] = front[3] + Tf * (w * (F + (1.f - F) * ((1.f - L->kmean) + L->kmean * c3.rgba[3])) + (1.f - w) * c3.rgba[3]);
    r->depth = c3.depth;
}

static int render_impl(const orc_model* m, const orc_render_params* P, const float* surf_rgba, const float* t_surface,
                       const float* lens_w, const float* lens_t, const float* lens_n, const lens_params_t* L,
                       float* frame, float* depth, uint32_t* n_samples, uint64_t* stats);

ORC_API int orc_render(const orc_model* m, const orc_render_params* P, const float* surf_rgba, const float* t_surface,
                       float* frame, float* depth, uint32_t* n_samples, uint64_t* stats) {
    return render_impl(m, P, surf_rgba, t_surface, NULL, NULL, NULL, NULL, frame, depth, n_samples, stats);
}

/* orc_render with lens hand-off: lens_w/lens_t [W*H], lens_n [W*H*3] (pixels with lens_w == 0 are ordinary), lens9 = f0, k[3],
 * kmean, background rgba (the colour the mirror rays see behind the NeRF, in the network's sRGB-like colour space). */
ORC_API int orc_render_lens(const orc_model* m, const orc_render_params* P, const float* surf_rgba, const float* t_surface,
                            const float* lens_w, const float* lens_t, const float* lens_n, const float* lens9,
                            float* frame, float* depth, uint32_t* n_samples, uint64_t* stats) {
    /* lens9[9..11] = model, thickness, ior (model 0: thin sheet) */
    lens_params_t L; L.f0 = lens9[0]; L.k[0] = lens9[1]; L.k[1] = lens9[2]; L.k[2] = lens9[3]; L.kmean = lens9[4]; memcpy(L.background, lens9 + 5, 16);
    L.model = (int)lens9[9]; L.thickness = lens9[10]; L.ior = lens9[11];
    return render_impl(m, P, surf_rgba, t_surface, lens_w, lens_t, lens_n, &L, frame, depth, n_samples, stats);
}

static int render_impl(const orc_model* m, const orc_render_params* P, const float* surf_rgba, const float* t_surface,
                       const float* lens_w, const float* lens_t, const float* lens_n, const lens_params_t* L,
                       float* frame, float* depth, uint32_t* n_samples, uint64_t* stats) {
    const int W = P->width, H = P->height;
    int x0 = P->x0, y0 = P->y0, x1 = P->x1, y1 = P->y1;
    if (x1 <= x0 || y1 <= y0) { x0 = 0; y0 = 0; x1 = W; y1 = H; }
    const int w = x1 - x0, h = y1 - y0;
    const size_t N = (size_t)w * h;
    aabb_t render_aabb = { v3_make(P->aabb_min[0], P->aabb_min[1], P->aabb_min[2]), v3_make(P->aabb_max[0], P->aabb_max[1], P->aabb_max[2]) };
    aabb_t train_aabb = { v3_make(P->train_aabb_min[0], P->train_aabb_min[1], P->train_aabb_min[2]), v3_make(P->train_aabb_max[0], P->train_aabb_max[1], P->train_aabb_max[2]) };
    ray_t* rays = (ray_t*)malloc(sizeof(ray_t) * N);
    if (!rays) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < (int64_t)N; ++i) {
        uint32_t x = (uint32_t)(x0 + (int)(i % w)), y = (uint32_t)(y0 + (int)(i / w));
        ray_t* r = &rays[i];
        init_ray(P, &render_aabb, x, y, r);
        if (t_surface) {
            r->t_surface = t_surface[r->idx];
            memcpy(r->surf, surf_rgba + (size_t)r->idx * 4, 16);
        }
        if (lens_w && lens_w[r->idx] > 0.f) { r->lens = 1; r->t_surface = lens_t[r->idx]; }   /* the first-hit walk is clipped / revived at the lens */
        depth[r->idx] = 1e10f;
        frame[(size_t)r->idx * 4 + 0] = frame[(size_t)r->idx * 4 + 1] = frame[(size_t)r->idx * 4 + 2] = frame[(size_t)r->idx * 4 + 3] = 0.f;
        advance_pos(m, P, &render_aabb, r);
    }
    uint64_t n_alive0 = 0;
    for (size_t i = 0; i < N; ++i) n_alive0 += rays[i].alive ? 1 : 0;
    uint64_t total_samples = 0, iterations = 0;
    uint32_t step = 1;
    /* Lens rays are marched on their own, outside the wavefront loop (which then holds the ordinary rays only).  Where the opaque
     * mesh surface behind a lens enters their compositing order: in batches of 8 samples while at most 1/8 of the pixels hold a
     * live ray - the regime in which the reference's own n_steps is the constant 8, i.e. a ray-local rule - and at the exact
     * sample otherwise (the reference's varying n_steps is a property of its wavefront, which lens rays are not part of). */
    uint32_t n_lens = 1;
    if (P->n_steps_mode == 1 && n_alive0 > 0) n_lens = ((uint64_t)N / n_alive0 >= 8) ? 8u : 1u;
    if (P->n_steps_mode == 2) n_lens = 8;
    uint64_t lens_samples = 0;
    if (L) {
#pragma omp parallel for schedule(dynamic, 4) reduction(+:lens_samples)
        for (int64_t i = 0; i < (int64_t)N; ++i) {
            ray_t* r = &rays[i];
            if (!r->lens || !r->alive) { if (r->lens) r->lens = 0; continue; }
            const float* sf = surf_rgba ? surf_rgba + (size_t)r->idx * 4 : NULL;
            float zero[4] = { 0.f, 0.f, 0.f, 0.f };
            march_lens_ray(m, P, &render_aabb, &train_aabb, L, r, n_lens, lens_w[r->idx], lens_t[r->idx],
                           v3_make(lens_n[(size_t)r->idx * 3], lens_n[(size_t)r->idx * 3 + 1], lens_n[(size_t)r->idx * 3 + 2]),
                           sf ? sf : zero, t_surface ? t_surface[r->idx] : 0.f);
            r->alive = 0;
            lens_samples += r->n_samples;
        }
    }
    total_samples += lens_samples;
    while (1) {
        uint64_t n_alive = 0;
        for (size_t i = 0; i < N; ++i) n_alive += rays[i].alive ? 1 : 0;
        if (n_alive == 0 || step >= 10000u) break;
        uint32_t n = 1;
        if (P->n_steps_mode == 1) {
            uint64_t q = (uint64_t)N / n_alive;     /* m_n_rays_initialized / n_alive, S/ngp/testbed.cu:1996 */
            n = (uint32_t)(q < 1 ? 1 : (q > 8 ? 8 : q));
        }
        if (P->n_steps_mode == 2) n = 8;
        uint64_t iter_samples = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+:iter_samples)
        for (int64_t i = 0; i < (int64_t)N; ++i) {
            ray_t* r = &rays[i];
            if (!r->alive) continue;
            sample_t smp[8]; uint16_t out[8 * 4];
            uint32_t cnt = generate_samples(m, P, &render_aabb, &train_aabb, r, n, smp, 0);
            v3 dir01 = v3_make((r->dir.x + 1.0f) * 0.5f, (r->dir.y + 1.0f) * 0.5f, (r->dir.z + 1.0f) * 0.5f);   /* warp_direction */
            for (uint32_t j = 0; j < cnt; ++j) network_eval(m, smp[j].pos, dir01, out + j * 4, NULL);
            r->n_samples += cnt; iter_samples += cnt;
            composite_ray(P, &train_aabb, r, n, smp, out, step);
        }
        total_samples += iter_samples;
        step += n;
        ++iterations;
    }
    uint64_t n_hit = 0;
    for (size_t i = 0; i < N; ++i) {
        ray_t* r = &rays[i];
        if (n_samples) n_samples[r->idx] = r->n_samples;
        /* compact_kernel_nerf: finished rays with alpha > 0.001 are shaded (S/ngp/testbed.cu:556) */
        if (!r->alive && r->rgba[3] > 0.001f) {
            /* shade_kernel_nerf (S/ngp/testbed.cu:907-931), train_in_linear_colors = false */
            float* fb = frame + (size_t)r->idx * 4;
            float tmp[4] = { srgb_to_linear(r->rgba[0]), srgb_to_linear(r->rgba[1]), srgb_to_linear(r->rgba[2]), r->rgba[3] };
            for (int k = 0; k < 4; ++k) fb[k] = tmp[k] + fb[k] * (1.0f - tmp[3]);
            if (tmp[3] > 0.2f) depth[r->idx] = r->depth;
            ++n_hit;
        }
    }
    if (stats) { stats[0] = n_alive0; stats[1] = total_samples; stats[2] = iterations; stats[3] = n_hit; }
    free(rays);
    return 0;
}

/* accumulate_kernel + tonemap_kernel (S/ngp/render_buffer.cu:232-267, 327-346, 537-566), colour space Linear,
 * tonemap curve Identity, exposure 0.  accum is updated in place; out gets the displayed float4 image. */
/* tonemap(x, curve) (S/ngp/render_buffer.cu:269-325): 0 Identity, 1 ACES, 2 Hable, 3 Reinhard */
static void tonemap_curve(float c[3], int curve) {
    if (curve == 0) return;
    for (int k = 0; k < 3; ++k) c[k] = fmaxf(c[k], 0.f);
    float k0, k1, k2, k3, k4, k5;
    if (curve == 1) {
        k0 = 0.6f * 0.6f * 2.51f; k1 = 0.6f * 0.03f; k2 = 0.0f; k3 = 0.6f * 0.6f * 2.43f; k4 = 0.6f * 0.59f; k5 = 0.14f;
    } else if (curve == 2) {
        const float A = 0.15f, B = 0.50f, Cc = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
        k0 = A * F - A * E; k1 = Cc * B * F - B * E; k2 = 0.0f; k3 = A * F; k4 = B * F; k5 = D * F * F;
        const float W = 11.2f;
        const float nom = k0 * (W * W) + k1 * W + k2, denom = k3 * (W * W) + k4 * W + k5;
        const float white_scale = denom / nom;
        k0 = 4.0f * k0 * white_scale; k1 = 2.0f * k1 * white_scale; k2 = k2 * white_scale; k3 = 4.0f * k3; k4 = 2.0f * k4;
    } else {
        const float Y = 0.2126f * c[0] + (0.7152f * c[1] + 0.0722f * c[2]);   /* Eigen 3-element dot */
        const float s = 1.f / (Y + 1.0f);
        for (int k = 0; k < 3; ++k) c[k] = c[k] * s;
        return;
    }
    for (int k = 0; k < 3; ++k) {
        const float sq = c[k] * c[k];
        const float nom = sq * k0 + k1 * c[k] + k2, denom = k3 * sq + k4 * c[k] + k5;
        c[k] = nom / denom;
    }
}

/* accumulate_kernel + tonemap_kernel (S/ngp/render_buffer.cu:232-267, 327-346, 537-566); exposure 0 (m_exposure is not
 * reachable from the Python API); tonemap_curve = Testbed.tonemap_curve */
ORC_API void orc_accumulate_tonemap_curve(const float* frame, float* accum, int64_t n_pixels, uint32_t spp_index,
                                          const float* background_rgba, int to_srgb, int curve, float* out) {
    const float sc = (float)spp_index;
    float bg[4] = { srgb_to_linear(background_rgba[0]), srgb_to_linear(background_rgba[1]), srgb_to_linear(background_rgba[2]), background_rgba[3] };
    const float expo = powf(2.0f, 0.0f);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_pixels; ++i) {
        float* a = accum + i * 4; const float* f = frame + i * 4;
        if (spp_index == 0) { a[0] = a[1] = a[2] = a[3] = 0.f; }
        for (int k = 0; k < 4; ++k) a[k] = (a[k] * sc + f[k]) / (sc + 1);
        float col[4] = { a[0], a[1], a[2], a[3] };
        float weight = (1 - col[3]) * bg[3];
        col[0] += bg[0] * weight; col[1] += bg[1] * weight; col[2] += bg[2] * weight; col[3] += weight;
        for (int k = 0; k < 3; ++k) col[k] *= expo;
        tonemap_curve(col, curve);
        if (to_srgb) for (int k = 0; k < 3; ++k) col[k] = linear_to_srgb(col[k]);
        if (to_srgb) for (int k = 0; k < 4; ++k) col[k] = fminf(fmaxf(col[k], 0.0f), 1.0f);
        memcpy(out + i * 4, col, 16);
    }
}
ORC_API void orc_accumulate_tonemap(const float* frame, float* accum, int64_t n_pixels, uint32_t spp_index,
                                    const float* background_rgba, int to_srgb, float* out) {
    orc_accumulate_tonemap_curve(frame, accum, n_pixels, spp_index, background_rgba, to_srgb, 0, out);
}

/* Traversal trace for bit-exactness tests: the first max_samples occupied samples of each listed pixel
 * (no network, no termination, mesh ignored), after the same init + first-hit advance as orc_render.
 * out_t/out_cell/out_mip: [n_pix][max_samples]; out_count[n_pix]; out_ray[n_pix][8] = origin3, dir3, t_first, alive. */
ORC_API void orc_trace_samples(const orc_model* m, const orc_render_params* P, const uint32_t* pixels, int64_t n_pix, uint32_t max_samples,
                               float* out_t, uint32_t* out_cell, uint32_t* out_mip, float* out_pos, uint32_t* out_count, float* out_ray) {
    aabb_t render_aabb = { v3_make(P->aabb_min[0], P->aabb_min[1], P->aabb_min[2]), v3_make(P->aabb_max[0], P->aabb_max[1], P->aabb_max[2]) };
    aabb_t train_aabb = { v3_make(P->train_aabb_min[0], P->train_aabb_min[1], P->train_aabb_min[2]), v3_make(P->train_aabb_max[0], P->train_aabb_max[1], P->train_aabb_max[2]) };
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_pix; ++i) {
        ray_t r; sample_t s;
        uint32_t x = pixels[i] % (uint32_t)P->width, y = pixels[i] / (uint32_t)P->width;
        init_ray(P, &render_aabb, x, y, &r);
        advance_pos(m, P, &render_aabb, &r);
        float* rr = out_ray + i * 8;
        rr[0] = r.origin.x; rr[1] = r.origin.y; rr[2] = r.origin.z; rr[3] = r.dir.x; rr[4] = r.dir.y; rr[5] = r.dir.z; rr[6] = r.t; rr[7] = (float)r.alive;
        uint32_t cnt = 0;
        while (r.alive && cnt < max_samples) {
            if (generate_samples(m, P, &render_aabb, &train_aabb, &r, 1, &s, 1) == 0) break;
            out_t[i * max_samples + cnt] = s.t;
            out_cell[i * max_samples + cnt] = s.cell;
            out_mip[i * max_samples + cnt] = s.mip;
            out_pos[(i * max_samples + cnt) * 3 + 0] = s.pos.x;
            out_pos[(i * max_samples + cnt) * 3 + 1] = s.pos.y;
            out_pos[(i * max_samples + cnt) * 3 + 2] = s.pos.z;
            ++cnt;
        }
        out_count[i] = cnt;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Density probes of the collision tool                                           */
/* ------------------------------------------------------------------------------------------ */
/* NerfTracer::intersects (S/ngp/testbed.cu:1891-1935) with the payloads NerfMeshRenderer::collide writes
 * (S/nerf_mesh_renderer.cu:1564-1574: origin = world point + 0.5): one network evaluation at the point with the minimum
 * step; out[i] = 1 - exp(-density * dt) when the cell under the point is occupied, untouched (0) otherwise. */
ORC_API void orc_probe_points(const orc_model* m, const orc_render_params* P, const float* points_world, const float* dir, int64_t n, float* out) {
    aabb_t train_aabb; memcpy(&train_aabb.min, P->train_aabb_min, 12); memcpy(&train_aabb.max, P->train_aabb_max, 12);
    v3 d = v3_make(dir[0], dir[1], dir[2]);
    v3 dir01 = v3_make((d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f);
    v3 diag = sub3(train_aabb.max, train_aabb.min);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        v3 origin = v3_make(points_world[i * 3] + 0.5f, points_world[i * 3 + 1] + 0.5f, points_world[i * 3 + 2] + 0.5f);
        v3 rel = sub3(origin, train_aabb.min);
        v3 warped = v3_make(rel.x / diag.x, rel.y / diag.y, rel.z / diag.z);
        uint16_t o4[4];
        network_eval(m, warped, dir01, o4, NULL);
        float dt = MIN_CONE_STEPSIZE();
        float alpha = 1.f - expf(-act_density(h2f(o4[3]), P->density_activation) * dt);
        v3 pos = v3_make(train_aabb.min.x + warped.x * diag.x, train_aabb.min.y + warped.y * diag.y, train_aabb.min.z + warped.z * diag.z);   /* unwarp_position */
        int mip = mip_from_dt(dt, pos); if (mip < 0) mip = 0;
        out[i] = 0.f;
        if (density_grid_occupied_at(pos, m->bitfield, (uint32_t)mip)) out[i] = alpha;
    }
}

/* NerfTracer::collide (S/ngp/testbed.cu:1814-1888) + check_collision (:721-782): rays from origin = world point + 0.5 along
 * `dir`, t = t_start = 0, batches of 8 samples; out[i] = |sample position - origin| of the first sample whose
 * alpha = 1 - exp(-density * dt) is positive, 0 for a ray that leaves the render box without one (the reference keeps such
 * a ray alive - its short batch is generated again and again until MARCH_ITER - and its distance stays at the memset 0). */
ORC_API void orc_probe_rays(const orc_model* m, const orc_render_params* P, const float* origins_world, const float* dir, int64_t n, float* out) {
    aabb_t render_aabb; memcpy(&render_aabb.min, P->aabb_min, 12); memcpy(&render_aabb.max, P->aabb_max, 12);
    aabb_t train_aabb; memcpy(&train_aabb.min, P->train_aabb_min, 12); memcpy(&train_aabb.max, P->train_aabb_max, 12);
    v3 d = v3_make(dir[0], dir[1], dir[2]);
    v3 dir01 = v3_make((d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f);
    v3 diag = sub3(train_aabb.max, train_aabb.min);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        ray_t r; memset(&r, 0, sizeof(r));
        r.origin = v3_make(origins_world[i * 3] + 0.5f, origins_world[i * 3 + 1] + 0.5f, origins_world[i * 3 + 2] + 0.5f);
        r.dir = d; r.alive = 1; r.idx = (uint32_t)i;
        out[i] = 0.f;
        int done = 0;
        while (!done) {
            sample_t smp[8];
            uint32_t got = generate_samples(m, P, &render_aabb, &train_aabb, &r, 8, smp, 1);
            for (uint32_t j = 0; j < got && !done; ++j) {
                uint16_t o4[4];
                network_eval(m, smp[j].pos, dir01, o4, NULL);
                float alpha = 1.f - expf(-act_density(h2f(o4[3]), P->density_activation) * unwarp_dt(smp[j].dt_warped));
                if (alpha > 0.f) {
                    v3 pos = v3_make(train_aabb.min.x + smp[j].pos.x * diag.x, train_aabb.min.y + smp[j].pos.y * diag.y, train_aabb.min.z + smp[j].pos.z * diag.z);
                    v3 dd = sub3(pos, r.origin);
                    out[i] = sqrtf(edot3(dd, dd));
                    done = 1;
                }
            }
            if (got < 8) done = 1;      /* left the render box: never collides */
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Mesh stage  (S/optix/optix_scene.cu:71-85, 120-325; S/optix/optix_util.cuh:23-29;            */
/*              S/gltf_scene.h:122-127; S/nerf_mesh_renderer.cu:64-100)                         */
/* OptiX's BVH traversal / triangle test live in the driver: "parity unpinned" at the           */
/* ray/triangle boundary; this is a brute-force Moeller-Trumbore with back-face culling.        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t n_verts, n_tris;
    v3* wpos;          /* world-space positions  */
    v3* wnrm;          /* world-space (inverse-transpose transformed, unnormalised) normals */
    float* uv;         /* [n_verts][2] */
    uint32_t* idx;     /* [n_tris][3] */
    float base_color[4], emissive[3], metallic, roughness;
    int tex_w, tex_h;
    float* tex_lin;    /* [h][w][4] linearised sRGB texture (alpha linear) or NULL */
    uint8_t* tri_lens;   /* per triangle: 1 = lens surface (NULL: none) */
    /* the other textures of __closesthit__ch (S/optix/optix_scene.cu:234-258): [0] emissive (sRGB), [1] metallicRoughness,
     * [2] normal, [3] occlusion (all linear); tangents as M3 t_obj / M3 n_obj with M3 = R S, and the normal matrix R S^-1 */
    float* xtex[4]; int xtex_w[4], xtex_h[4];
    float normal_scale, occlusion_strength;
    float* wtbn;       /* [n_verts][8]: M3 n (3), M3 t (3), tangent.w, 0; NULL without tangents */
    float nmat[9];
} orc_mesh;

/* r_wxyz: the (w,x,y,z) quaternion NerfMeshRenderer::loadMesh builds from its Vector4f argument
 * (S/nerf_mesh_renderer.cu:954).  M = T * R * S (S/gltf_scene.h:122-127); world = M * v. */
ORC_API orc_mesh* orc_mesh_create(const float* pos, const float* nrm, const float* uv, uint32_t n_verts, const uint16_t* indices, uint32_t n_idx,
                                  const float* t, const float* s, const float* r_wxyz, const float* base_color, float metallic, float roughness,
                                  const float* emissive, const uint8_t* tex_rgba8, int tex_w, int tex_h) {
    orc_mesh* M = (orc_mesh*)calloc(1, sizeof(orc_mesh));
    M->n_verts = n_verts; M->n_tris = n_idx / 3;
    M->wpos = (v3*)malloc(sizeof(v3) * n_verts); M->wnrm = (v3*)malloc(sizeof(v3) * n_verts);
    M->uv = (float*)malloc(sizeof(float) * 2 * n_verts); M->idx = (uint32_t*)malloc(sizeof(uint32_t) * n_idx);
    memcpy(M->uv, uv, sizeof(float) * 2 * n_verts);
    for (uint32_t i = 0; i < n_idx; ++i) M->idx[i] = indices[i];
    /* glm::mat3_cast(quat) */
    float qw = r_wxyz[0], qx = r_wxyz[1], qy = r_wxyz[2], qz = r_wxyz[3];
    float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    float R[9] = { /* row-major */
        1.f - 2.f * (qyy + qzz), 2.f * (qxy - qwz), 2.f * (qxz + qwy),
        2.f * (qxy + qwz), 1.f - 2.f * (qxx + qzz), 2.f * (qyz - qwx),
        2.f * (qxz - qwy), 2.f * (qyz + qwx), 1.f - 2.f * (qxx + qyy) };
    for (uint32_t i = 0; i < n_verts; ++i) {
        v3 p = v3_make(pos[i * 3] * s[0], pos[i * 3 + 1] * s[1], pos[i * 3 + 2] * s[2]);
        v3 w = glm_mat3_mul(R, p);
        M->wpos[i] = v3_make(w.x + t[0], w.y + t[1], w.z + t[2]);
        v3 n = v3_make(nrm[i * 3] / s[0], nrm[i * 3 + 1] / s[1], nrm[i * 3 + 2] / s[2]);   /* (R S)^-T n = R S^-1 n */
        M->wnrm[i] = glm_mat3_mul(R, n);
    }
    memcpy(M->base_color, base_color, 16); memcpy(M->emissive, emissive, 12);
    M->metallic = metallic; M->roughness = roughness;
    if (tex_rgba8 && tex_w > 0 && tex_h > 0) {
        M->tex_w = tex_w; M->tex_h = tex_h;
        M->tex_lin = (float*)malloc(sizeof(float) * 4 * (size_t)tex_w * tex_h);
        for (size_t i = 0; i < (size_t)tex_w * tex_h; ++i) {
            for (int k = 0; k < 3; ++k) M->tex_lin[i * 4 + k] = srgb_to_linear((float)tex_rgba8[i * 4 + k] / 255.0f);
            M->tex_lin[i * 4 + 3] = (float)tex_rgba8[i * 4 + 3] / 255.0f;
        }
    }
    return M;
}
ORC_API void orc_mesh_destroy(orc_mesh* M) {
    if (!M) return;
    free(M->wpos); free(M->wnrm); free(M->uv); free(M->idx); free(M->tex_lin); free(M->tri_lens); free(M->wtbn);
    for (int k = 0; k < 4; ++k) free(M->xtex[k]);
    free(M);
}
/* which: 0 emissive (sRGB-decoded like CudaTexture's bSrgb, S/gltf_scene.cpp:179), 1 metallicRoughness, 2 normal, 3 occlusion
 * (linear, :196-211).  factor: normalTexture.scale / occlusionTexture.strength for 2 / 3, ignored otherwise. */
ORC_API void orc_mesh_set_texture(orc_mesh* M, int which, const uint8_t* rgba8, int w, int h, float factor) {
    if (which < 0 || which > 3) return;
    free(M->xtex[which]); M->xtex[which] = NULL; M->xtex_w[which] = M->xtex_h[which] = 0;
    if (which == 2) M->normal_scale = factor;
    if (which == 3) M->occlusion_strength = factor;
    if (!rgba8 || w <= 0 || h <= 0) return;
    float* t = (float*)malloc(sizeof(float) * 4 * (size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        for (int k = 0; k < 3; ++k) { float sv = (float)rgba8[i * 4 + k] / 255.0f; t[i * 4 + k] = which == 0 ? srgb_to_linear(sv) : sv; }
        t[i * 4 + 3] = (float)rgba8[i * 4 + 3] / 255.0f;
    }
    M->xtex[which] = t; M->xtex_w[which] = w; M->xtex_h[which] = h;
}
/* object-space normals and tangents (xyz + handedness) + the mesh transform: what computeTbnMatrix needs */
ORC_API void orc_mesh_set_tangents(orc_mesh* M, const float* nrm_obj, const float* tan4_obj, const float* s, const float* r_wxyz) {
    free(M->wtbn); M->wtbn = NULL;
    float qw = r_wxyz[0], qx = r_wxyz[1], qy = r_wxyz[2], qz = r_wxyz[3];
    float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    float R[9] = { 1.f - 2.f * (qyy + qzz), 2.f * (qxy - qwz), 2.f * (qxz + qwy), 2.f * (qxy + qwz), 1.f - 2.f * (qxx + qzz), 2.f * (qyz - qwx),
                   2.f * (qxz - qwy), 2.f * (qyz + qwx), 1.f - 2.f * (qxx + qyy) };
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) M->nmat[i * 3 + j] = R[i * 3 + j] / s[j];
    if (!tan4_obj) return;
    M->wtbn = (float*)malloc(sizeof(float) * 8 * M->n_verts);
    for (uint32_t i = 0; i < M->n_verts; ++i) {
        v3 n = glm_mat3_mul(R, v3_make(nrm_obj[i * 3] * s[0], nrm_obj[i * 3 + 1] * s[1], nrm_obj[i * 3 + 2] * s[2]));
        v3 t = glm_mat3_mul(R, v3_make(tan4_obj[i * 4] * s[0], tan4_obj[i * 4 + 1] * s[1], tan4_obj[i * 4 + 2] * s[2]));
        float* o = M->wtbn + (size_t)i * 8;
        o[0] = n.x; o[1] = n.y; o[2] = n.z; o[3] = t.x; o[4] = t.y; o[5] = t.z; o[6] = tan4_obj[i * 4 + 3]; o[7] = 0.f;
    }
}
ORC_API void orc_mesh_world_positions(const orc_mesh* M, float* out) { memcpy(out, M->wpos, sizeof(v3) * M->n_verts); }

/* bilinear, wrap addressing, normalised coordinates, texel centres at +0.5 (S/cuda_texture.cu:22-28) */
static void tex_sample_any(const float* tex, int tw, int th, float u, float v, float out[4]);
static void tex_sample(const orc_mesh* M, float u, float v, float out[4]) { tex_sample_any(M->tex_lin, M->tex_w, M->tex_h, u, v, out); }
static void tex_sample_any(const float* tex, int tw, int th, float u, float v, float out[4]) {
    struct { int tex_w, tex_h; const float* tex_lin; } Ms = { tw, th, tex }, *M = &Ms;
    float fx = u * (float)M->tex_w - 0.5f, fy = v * (float)M->tex_h - 0.5f;
    float flx = floorf(fx), fly = floorf(fy);
    float ax = fx - flx, ay = fy - fly;
    int x0 = (int)flx, y0 = (int)fly;
    int xs[2], ys[2];
    xs[0] = ((x0 % M->tex_w) + M->tex_w) % M->tex_w; xs[1] = (((x0 + 1) % M->tex_w) + M->tex_w) % M->tex_w;
    ys[0] = ((y0 % M->tex_h) + M->tex_h) % M->tex_h; ys[1] = (((y0 + 1) % M->tex_h) + M->tex_h) % M->tex_h;
    for (int k = 0; k < 4; ++k) {
        float t00 = M->tex_lin[((size_t)ys[0] * M->tex_w + xs[0]) * 4 + k], t10 = M->tex_lin[((size_t)ys[0] * M->tex_w + xs[1]) * 4 + k];
        float t01 = M->tex_lin[((size_t)ys[1] * M->tex_w + xs[0]) * 4 + k], t11 = M->tex_lin[((size_t)ys[1] * M->tex_w + xs[1]) * 4 + k];
        float top = t00 * (1.0f - ax) + t10 * ax, bot = t01 * (1.0f - ax) + t11 * ax;
        out[k] = top * (1.0f - ay) + bot * ay;
    }
}

static inline float to_srgb_optix(float c) {   /* S/optix/optix_util.cuh:23-29 */
    float powed = powf(c, 1.0f / 2.4f);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}

/* Moeller-Trumbore, front faces only (det > 0 <=> counter-clockwise seen from the ray origin) */
static inline int ray_tri(v3 o, v3 d, v3 v0, v3 v1, v3 v2, float* t_out, float* u_out, float* v_out) {
    v3 e1 = sub3(v1, v0), e2 = sub3(v2, v0);
    v3 pv = cross3(d, e2);
    float det = dot3(e1, pv);
    if (!(det > 0.0f)) return 0;
    v3 tv = sub3(o, v0);
    float u = dot3(tv, pv);
    if (u < 0.0f || u > det) return 0;
    v3 qv = cross3(tv, e1);
    float v = dot3(d, qv);
    if (v < 0.0f || u + v > det) return 0;
    float t = dot3(e2, qv);
    if (!(t > 0.0f)) return 0;
    float inv = 1.0f / det;
    *t_out = t * inv; *u_out = u * inv; *v_out = v * inv;
    return 1;
}

static void shade_hit(const orc_mesh* M, uint32_t tri, float bu, float bv, float hitT, v3 eye, v3 dir, v3 light, float rgba[4]) {
    const uint32_t i0 = M->idx[tri * 3], i1 = M->idx[tri * 3 + 1], i2 = M->idx[tri * 3 + 2];
    float bw = 1.0f - bu - bv;
    v3 n = add3(add3(mul3(M->wnrm[i1], bu), mul3(M->wnrm[i2], bv)), mul3(M->wnrm[i0], bw));
    float uvx = (bu * M->uv[i1 * 2] + bv * M->uv[i2 * 2]) + bw * M->uv[i0 * 2];
    float uvy = (bu * M->uv[i1 * 2 + 1] + bv * M->uv[i2 * 2 + 1]) + bw * M->uv[i0 * 2 + 1];
    float base[4] = { M->base_color[0], M->base_color[1], M->base_color[2], M->base_color[3] };
    if (M->tex_lin) { float tx[4]; tex_sample(M, uvx, uvy, tx); for (int k = 0; k < 4; ++k) base[k] *= tx[k]; }
    float metallic = M->metallic, roughness = M->roughness, occlusion = 1.0f;
    float emissive[3] = { M->emissive[0], M->emissive[1], M->emissive[2] };
    if (M->xtex[0]) { float tx[4]; tex_sample_any(M->xtex[0], M->xtex_w[0], M->xtex_h[0], uvx, uvy, tx); for (int k = 0; k < 3; ++k) emissive[k] *= tx[k]; }
    if (M->xtex[1]) { float tx[4]; tex_sample_any(M->xtex[1], M->xtex_w[1], M->xtex_h[1], uvx, uvy, tx); metallic *= tx[2]; roughness *= tx[1]; }
    if (M->xtex[3]) { float tx[4]; tex_sample_any(M->xtex[3], M->xtex_w[3], M->xtex_h[3], uvx, uvy, tx); occlusion = 1.0f + M->occlusion_strength * (tx[0] - 1.0f); }
    if (M->xtex[2] && M->wtbn) {
        /* computeTbnMatrix (S/optix/optix_scene.cu:92-98), the mapped normal (:245-251), then the object-to-world normal
         * transform once more, as the reference applies it to the already transformed vector (:252) */
        const float* a0 = M->wtbn + (size_t)i0 * 8; const float* a1 = M->wtbn + (size_t)i1 * 8; const float* a2 = M->wtbn + (size_t)i2 * 8;
        float q[7];
        for (int k = 0; k < 7; ++k) q[k] = (bu * a1[k] + bv * a2[k]) + bw * a0[k];
        v3 tn = glm_normalize3(v3_make(q[3], q[4], q[5]));
        v3 nn = glm_normalize3(v3_make(q[0], q[1], q[2]));
        tn = glm_normalize3(sub3(tn, mul3(nn, dot3(tn, nn))));
        v3 bn = mul3(cross3(nn, tn), q[6]);
        float tx[4]; tex_sample_any(M->xtex[2], M->xtex_w[2], M->xtex_h[2], uvx, uvy, tx);
        float mx = (tx[0] * 2.0f - 1.0f) * M->normal_scale, my = (tx[1] * 2.0f - 1.0f) * M->normal_scale, mz = tx[2] * 2.0f - 1.0f;
        v3 mm = add3(add3(mul3(tn, mx), mul3(bn, my)), mul3(nn, mz));
        n = glm_mat3_mul(M->nmat, mm);
    }
    v3 hitPos = add3(eye, mul3(dir, hitT));
    v3 N = glm_normalize3(n);
    v3 V = glm_normalize3(sub3(eye, hitPos));
    v3 L = glm_normalize3(sub3(light, hitPos));
    v3 Hh = glm_normalize3(add3(V, L));
    float ndl = dot3(L, N);
    float dl = fmaxf(0.f, ndl);
    float fd[3] = { (1.0f - metallic) * base[0] * dl, (1.0f - metallic) * base[1] * dl, (1.0f - metallic) * base[2] * dl };
    float fr[3] = { 0.f, 0.f, 0.f };
    float dotNV = dot3(N, V), dotNL = ndl;
    if (dotNV > 0 && dotNL > 0) {
        float dotNH = clampf(dot3(N, Hh), 0.0f, 1.0f);
        float dotLH = clampf(dot3(L, Hh), 0.0f, 1.0f);
        float alpha = roughness * roughness;
        float a2 = alpha * alpha;                                   /* dGgx/gGgx square their "roughness" argument again */
        float f = (dotNH * a2 - dotNH) * dotNH + 1.0f;
        float D = a2 / (f * f);
        float lambdaV = fmaxf(0.f, dotNL) / sqrtf(a2 + (1.0f - a2) * dotNV * dotNV);
        float lambdaL = fmaxf(0.f, dotNV) / sqrtf(a2 + (1.0f - a2) * dotNL * dotNL);
        float G = 0.5f / (lambdaV + lambdaL + 0.0001f);
        float p5 = powf(1.0f - dotLH, 5.0f);
        for (int k = 0; k < 3; ++k) {
            float f0 = (0.5f * alpha) * (1.0f - metallic) + base[k] * metallic;   /* glm::mix */
            float F = f0 + (1.0f - f0) * p5;
            fr[k] = fabsf((D * G * F) / 3.14159265358979323846f);
        }
    }
    for (int k = 0; k < 3; ++k) {
        float ambient = base[k] * .2f * occlusion;
        float c = ambient + (fd[k] + fr[k]) + emissive[k];
        c = clampf(c, 0.f, 1.f);
        rgba[k] = to_srgb_optix(c);
    }
    rgba[3] = 1.f;
}

/* One primary ray per pixel of the (mesh_scale x) supersampled buffer; writes RGBA (alpha 1 hit / 0 miss) and
 * hitT (NaN on miss: the reference stores the uint payload -1 reinterpreted as float).  out_tri optional. */
/* window = {x0, y0, x1, y1} in supersampled pixels restricts the work to a sub-rectangle (others untouched); NULL = all */
ORC_API void orc_mesh_render(const orc_mesh* M, const float* camera12, const float* light_pos, int W2, int H2,
                             float* rgba, float* depth, int32_t* out_tri, const int32_t* window) {
    const v3 U = v3_make(camera12[0], camera12[1], camera12[2]), Vv = v3_make(camera12[3], camera12[4], camera12[5]);
    const v3 Wv = v3_make(camera12[6], camera12[7], camera12[8]), eye = v3_make(camera12[9], camera12[10], camera12[11]);
    const v3 light = v3_make(light_pos[0], light_pos[1], light_pos[2]);
    int wx0 = 0, wy0 = 0, wx1 = W2, wy1 = H2;
    if (window && window[2] > window[0] && window[3] > window[1]) { wx0 = window[0]; wy0 = window[1]; wx1 = window[2]; wy1 = window[3]; }
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = wy0; y < wy1; ++y) {
        for (int x = wx0; x < wx1; ++x) {
            float dx = 2.0f * (((float)x + 0.5f) / (float)W2) - 1.0f;
            float dy = 2.0f * (((float)y + 0.5f) / (float)H2) - 1.0f;
            v3 dir = glm_normalize3(v3_make((dx * U.x + dy * Vv.x) + Wv.x, (dx * U.y + dy * Vv.y) + Wv.y, (dx * U.z + dy * Vv.z) + Wv.z));
            float best_t = 1e16f, bu = 0, bv = 0; int32_t best = -1;
            for (uint32_t tri = 0; tri < M->n_tris; ++tri) {
                float t, u, v;
                if (ray_tri(eye, dir, M->wpos[M->idx[tri * 3]], M->wpos[M->idx[tri * 3 + 1]], M->wpos[M->idx[tri * 3 + 2]], &t, &u, &v) && t < best_t) {
                    best_t = t; bu = u; bv = v; best = (int32_t)tri;
                }
            }
            size_t i = (size_t)y * W2 + x;
            if (best >= 0) {
                shade_hit(M, (uint32_t)best, bu, bv, best_t, eye, dir, light, rgba + i * 4);
                depth[i] = best_t;
            } else {
                rgba[i * 4] = rgba[i * 4 + 1] = rgba[i * 4 + 2] = rgba[i * 4 + 3] = 0.f;
                uint32_t nanbits = 0xFFFFFFFFu; memcpy(&depth[i], &nanbits, 4);
            }
            if (out_tri) out_tri[i] = best;
        }
    }
}

/* ---- lens surfaces (new functionality, see march_lens_ray) ---- */
ORC_API void orc_mesh_set_lens(orc_mesh* M, const uint8_t* tri_lens) {
    free(M->tri_lens); M->tri_lens = NULL;
    if (tri_lens) { M->tri_lens = (uint8_t*)malloc(M->n_tris); memcpy(M->tri_lens, tri_lens, M->n_tris); }
}

/* Like orc_mesh_render but in two layers: the nearest OPAQUE hit (shaded RGBA + hitT, NaN on miss) and the nearest LENS hit
 * (hitT, 0 on miss, and the unit shading normal).  window as in orc_mesh_render (the caller pre-fills the rest as misses). */
ORC_API void orc_mesh_render_layers(const orc_mesh* M, const float* camera12, const float* light_pos, int W2, int H2,
                                    float* rgba, float* depth, float* lens_depth, float* lens_normal, const int32_t* window) {
    const v3 U = v3_make(camera12[0], camera12[1], camera12[2]), Vv = v3_make(camera12[3], camera12[4], camera12[5]);
    const v3 Wv = v3_make(camera12[6], camera12[7], camera12[8]), eye = v3_make(camera12[9], camera12[10], camera12[11]);
    const v3 light = v3_make(light_pos[0], light_pos[1], light_pos[2]);
    int wx0 = 0, wy0 = 0, wx1 = W2, wy1 = H2;
    if (window && window[2] > window[0] && window[3] > window[1]) { wx0 = window[0]; wy0 = window[1]; wx1 = window[2]; wy1 = window[3]; }
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = wy0; y < wy1; ++y) {
        for (int x = wx0; x < wx1; ++x) {
            float dx = 2.0f * (((float)x + 0.5f) / (float)W2) - 1.0f;
            float dy = 2.0f * (((float)y + 0.5f) / (float)H2) - 1.0f;
            v3 dir = glm_normalize3(v3_make((dx * U.x + dy * Vv.x) + Wv.x, (dx * U.y + dy * Vv.y) + Wv.y, (dx * U.z + dy * Vv.z) + Wv.z));
            float best_t[2] = { 1e16f, 1e16f }, bu[2] = { 0, 0 }, bv[2] = { 0, 0 }; int32_t best[2] = { -1, -1 };
            for (uint32_t tri = 0; tri < M->n_tris; ++tri) {
                float t, u, v;
                const int layer = (M->tri_lens && M->tri_lens[tri]) ? 1 : 0;
                if (ray_tri(eye, dir, M->wpos[M->idx[tri * 3]], M->wpos[M->idx[tri * 3 + 1]], M->wpos[M->idx[tri * 3 + 2]], &t, &u, &v) && t < best_t[layer]) {
                    best_t[layer] = t; bu[layer] = u; bv[layer] = v; best[layer] = (int32_t)tri;
                }
            }
            size_t i = (size_t)y * W2 + x;
            if (best[0] >= 0) { shade_hit(M, (uint32_t)best[0], bu[0], bv[0], best_t[0], eye, dir, light, rgba + i * 4); depth[i] = best_t[0]; }
            else { rgba[i * 4] = rgba[i * 4 + 1] = rgba[i * 4 + 2] = rgba[i * 4 + 3] = 0.f; uint32_t nanbits = 0xFFFFFFFFu; memcpy(&depth[i], &nanbits, 4); }
            lens_depth[i] = 0.f; lens_normal[i * 3] = 0.f; lens_normal[i * 3 + 1] = 0.f; lens_normal[i * 3 + 2] = 1.f;
            if (best[1] >= 0) {
                const uint32_t tri = (uint32_t)best[1], i0 = M->idx[tri * 3], i1 = M->idx[tri * 3 + 1], i2 = M->idx[tri * 3 + 2];
                float bw = 1.0f - bu[1] - bv[1];
                v3 n = glm_normalize3(add3(add3(mul3(M->wnrm[i1], bu[1]), mul3(M->wnrm[i2], bv[1])), mul3(M->wnrm[i0], bw)));
                lens_depth[i] = best_t[1]; lens_normal[i * 3] = n.x; lens_normal[i * 3 + 1] = n.y; lens_normal[i * 3 + 2] = n.z;
            }
        }
    }
}

/* Lens hand-off per pixel: a tap counts when it has a lens hit in front of its opaque hit; w = counted / taps, t = the nearest
 * counted hit (first in tap order on ties), n = its normal.  The event is dropped when the pixel's opaque t_surface (max over
 * the taps) lies in front of it. */
ORC_API void orc_lens_resolve(const float* depth2, const float* lens_depth2, const float* lens_normal2, const float* t_surface, int W, int H, int mesh_scale,
                              float* out_w, float* out_t, float* out_n) {
    const int W2 = W * mesh_scale;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < W * H; ++idx) {
        int px = idx % W, py = idx / W, n_l = 0; float best = 3.402823466e+38f; size_t bi = 0;
        for (int i = 0; i < mesh_scale; ++i) for (int j = 0; j < mesh_scale; ++j) {
            size_t o = (size_t)(py * mesh_scale + j) * W2 + (size_t)(px * mesh_scale + i);
            float tl = lens_depth2[o];
            if (!(tl > 0.f)) continue;
            float to = depth2[o];
            if (to == to && !(tl < to)) continue;     /* opaque hit (not NaN) at or in front of the lens */
            ++n_l;
            if (tl < best) { best = tl; bi = o; }
        }
        float w = n_l ? (float)n_l / (float)(mesh_scale * mesh_scale) : 0.f;
        if (w > 0.f && t_surface[idx] != 0.0f && t_surface[idx] < best) w = 0.f;
        out_w[idx] = w; out_t[idx] = w > 0.f ? best : 0.f;
        out_n[idx * 3] = n_l ? lens_normal2[bi * 3] : 0.f; out_n[idx * 3 + 1] = n_l ? lens_normal2[bi * 3 + 1] : 0.f; out_n[idx * 3 + 2] = n_l ? lens_normal2[bi * 3 + 2] : 1.f;
    }
}

/* copyRaytracingBuffersToNerfRays (S/nerf_mesh_renderer.cu:64-100): colour = mean of the taps, depth = max (NaN-ignoring) */
ORC_API void orc_mesh_resolve(const float* rgba2, const float* depth2, int W, int H, int mesh_scale, float* surf_rgba, float* t_surface) {
    const int mesh_pitch = W * mesh_scale;
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < W * H; ++idx) {
        int x = idx % W, y = idx / W;
        size_t mesh_idx = (size_t)y * mesh_pitch * mesh_scale + (size_t)x * mesh_scale;
        float c[4] = { 0.f, 0.f, 0.f, 0.f }; float d = 0.f;
        for (int i = 0; i < mesh_scale; ++i) for (int j = 0; j < mesh_scale; ++j) {
            const float* p = rgba2 + (mesh_idx + i + (size_t)j * mesh_pitch) * 4;
            c[0] += p[0]; c[1] += p[1]; c[2] += p[2]; c[3] += p[3];
            d = fmaxf(d, depth2[mesh_idx + i + (size_t)j * mesh_pitch]);
        }
        float q = (float)(mesh_scale * mesh_scale);
        surf_rgba[idx * 4] = c[0] / q; surf_rgba[idx * 4 + 1] = c[1] / q; surf_rgba[idx * 4 + 2] = c[2] / q; surf_rgba[idx * 4 + 3] = c[3] / q;
        t_surface[idx] = d;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Floatie removal  (S/floatyremover.h:11-267; S/nerf_mesh_renderer.cu:239-336, 901-917)        */
/* Works on the dumped representation: cells[8][128][128][128] bytes, index x + 128*(y + 128*(z + 128*lvl)). */
/* ------------------------------------------------------------------------------------------ */
static inline size_t cell_idx(int x, int y, int z, int lvl) { return (size_t)x + 128u * ((size_t)y + 128u * ((size_t)z + 128u * (size_t)lvl)); }

/* dumpDensityGrid: bitfield (Morton) -> byte cells */
ORC_API void orc_bitfield_to_cells(const uint8_t* bitfield, uint8_t* cells) {
    for (int mip = 0; mip < 8; ++mip) for (int z = 0; z < 128; ++z) for (int y = 0; y < 128; ++y) for (int x = 0; x < 128; ++x) {
        float xx = ((float)x / 128.0f - 0.5f) * (float)(1 << mip) + 0.5f;
        float yy = ((float)y / 128.0f - 0.5f) * (float)(1 << mip) + 0.5f;
        float zz = ((float)z / 128.0f - 0.5f) * (float)(1 << mip) + 0.5f;
        uint32_t idx = cascaded_grid_idx_at(v3_make(xx, yy, zz), (uint32_t)mip);
        cells[cell_idx(x, y, z, mip)] = (bitfield[idx / 8 + (size_t)GRID_CELLS * mip / 8] >> (idx % 8)) & 1u;
    }
}
/* loadDensityGrid: byte cells -> bitfield */
ORC_API void orc_cells_to_bitfield(const uint8_t* cells, uint8_t* bitfield) {
    memset(bitfield, 0, NERF_CASCADES * GRID_CELLS / 8);
    for (int mip = 0; mip < 8; ++mip) for (int z = 0; z < 128; ++z) for (int y = 0; y < 128; ++y) for (int x = 0; x < 128; ++x) {
        float xx = ((float)x / 128.0f - 0.5f) * (float)(1 << mip) + 0.5f;
        float yy = ((float)y / 128.0f - 0.5f) * (float)(1 << mip) + 0.5f;
        float zz = ((float)z / 128.0f - 0.5f) * (float)(1 << mip) + 0.5f;
        uint32_t idx = cascaded_grid_idx_at(v3_make(xx, yy, zz), (uint32_t)mip);
        if (cells[cell_idx(x, y, z, mip)]) bitfield[idx / 8 + (size_t)GRID_CELLS * mip / 8] |= (uint8_t)(1u << (idx % 8));
    }
}

/* is (x,y,z,lvl) a member of NgpGrid::density_points?  (ctor, floatyremover.h:35-52) */
static inline int is_point(const uint8_t* cells, int x, int y, int z, int lvl) {
    if (x < 0 || y < 0 || z < 0 || x > 127 || y > 127 || z > 127 || lvl < 0 || lvl > 7) return 0;
    if (lvl > 0 && x >= 32 && x < 96 && y >= 32 && y < 96 && z >= 32 && z < 96) return 0;
    return cells[cell_idx(x, y, z, lvl)] != 0;
}

/* get_neighbors (floatyremover.h:60-193), including its uint8 wrap-around of x-1 at 0 / cx arithmetic */
static int neighbors(const uint8_t* cells, int x, int y, int z, int mip, int out[][4]) {
    int n = 0;
#define PUSH(X, Y, Z, L) do { if (is_point(cells, (X), (Y), (Z), (L))) { out[n][0] = (X); out[n][1] = (Y); out[n][2] = (Z); out[n][3] = (L); ++n; } } while (0)
    /* `x - 1 >= 0` is evaluated in int, but the MipPoint ctor narrows to uint8: (uint8)(-1) = 255 is never a member */
    if (x - 1 >= 0) PUSH(x - 1, y, z, mip);
    if (x + 1 < 128) PUSH(x + 1, y, z, mip);
    if (y - 1 >= 0) PUSH(x, y - 1, z, mip);
    if (y + 1 < 128) PUSH(x, y + 1, z, mip);
    if (z - 1 >= 0) PUSH(x, y, z - 1, mip);
    if (z + 1 < 128) PUSH(x, y, z + 1, mip);
    if (mip < 7) {
        int mx = 32 + x / 2, my = 32 + y / 2, mz = 32 + z / 2;
        if (x == 0) PUSH(31, my, mz, mip + 1);
        if (x == 127) PUSH(96, my, mz, mip + 1);
        if (y == 0) PUSH(mx, 31, mz, mip + 1);
        if (y == 127) PUSH(mx, 96, mz, mip + 1);
        if (z == 0) PUSH(mx, my, 31, mip + 1);
        if (z == 127) PUSH(mx, my, 96, mip + 1);
    }
    if (mip > 0) {
        int cx = (uint8_t)(x * 2 - 64), cy = (uint8_t)(y * 2 - 64), cz = (uint8_t)(z * 2 - 64);
        /* cy + 1 etc. are narrowed to uint8 again by the MipPoint ctor */
#define U8(v) ((int)(uint8_t)(v))
        if (x == 31) { PUSH(0, U8(cy + 0), U8(cz + 0), mip - 1); PUSH(0, U8(cy + 0), U8(cz + 1), mip - 1); PUSH(0, U8(cy + 1), U8(cz + 0), mip - 1); PUSH(0, U8(cy + 1), U8(cz + 1), mip - 1); }
        if (x == 96) { PUSH(127, U8(cy + 0), U8(cz + 0), mip - 1); PUSH(127, U8(cy + 0), U8(cz + 1), mip - 1); PUSH(127, U8(cy + 1), U8(cz + 0), mip - 1); PUSH(127, U8(cy + 1), U8(cz + 1), mip - 1); }
        if (y == 31) { PUSH(U8(cx + 0), 0, U8(cz + 0), mip - 1); PUSH(U8(cx + 0), 0, U8(cz + 1), mip - 1); PUSH(U8(cx + 1), 0, U8(cz + 0), mip - 1); PUSH(U8(cx + 1), 0, U8(cz + 1), mip - 1); }
        if (y == 96) { PUSH(U8(cx + 0), 127, U8(cz + 0), mip - 1); PUSH(U8(cx + 0), 127, U8(cz + 1), mip - 1); PUSH(U8(cx + 1), 127, U8(cz + 0), mip - 1); PUSH(U8(cx + 1), 127, U8(cz + 1), mip - 1); }
        if (z == 31) { PUSH(U8(cx + 0), U8(cy + 0), 0, mip - 1); PUSH(U8(cx + 0), U8(cy + 1), 0, mip - 1); PUSH(U8(cx + 1), U8(cy + 0), 0, mip - 1); PUSH(U8(cx + 1), U8(cy + 1), 0, mip - 1); }
        if (z == 96) { PUSH(U8(cx + 0), U8(cy + 0), 127, mip - 1); PUSH(U8(cx + 0), U8(cy + 1), 127, mip - 1); PUSH(U8(cx + 1), U8(cy + 0), 127, mip - 1); PUSH(U8(cx + 1), U8(cy + 1), 127, mip - 1); }
#undef U8
    }
#undef PUSH
    return n;
}

/* cluster() + max_element(point_set_importance) + to_ngp_grid.
 * The reference's neighbour relation is not symmetric across cascade boundaries in all cases and its BFS
 * follows directed edges from an arbitrary unordered_set start point, so cluster membership can depend on
 * hash-set iteration order when the relation is asymmetric; this restatement grows clusters over the same
 * directed relation starting from the lowest linear index (x fastest), which coincides with the reference
 * whenever the relation is symmetric on the input (always the case for aabb_scale = 1, where only cascade 0
 * contributes points).  Importance = sum(16 - 2^level) (exact integer; the reference accumulates through a
 * float, identical below 2^24).  Ties: first cluster in discovery order (reference: unordered_set order).
 * cells is rewritten in place; returns the number of clusters (points with >= 1 neighbour), or -1 if none. */
ORC_API int orc_remove_floaties(uint8_t* cells, int64_t* out_best_size, int64_t* out_best_score) {
    const size_t total = (size_t)8 * GRID_CELLS;
    int32_t* label = (int32_t*)malloc(sizeof(int32_t) * total);
    uint32_t* queue = (uint32_t*)malloc(sizeof(uint32_t) * total);
    if (!label || !queue) { free(label); free(queue); return -2; }
    for (size_t i = 0; i < total; ++i) label[i] = -1;
    int n_clusters = 0; int64_t best_score = 0, best_size = 0; int best = -1;
    int nb[32][4];
    for (size_t start = 0; start < total; ++start) {
        int x = (int)(start % 128), y = (int)((start / 128) % 128), z = (int)((start / (128 * 128)) % 128), l = (int)(start / GRID_CELLS);
        if (label[start] != -1 || !is_point(cells, x, y, z, l)) continue;
        if (neighbors(cells, x, y, z, l, nb) == 0) { label[start] = -2; continue; }   /* isolated: never forms a cluster */
        size_t head = 0, tail = 0;
        queue[tail++] = (uint32_t)start; label[start] = n_clusters;
        int64_t score = 0, size = 0;
        while (head < tail) {
            uint32_t cur = queue[head++];
            int cx = (int)(cur % 128), cy = (int)((cur / 128) % 128), cz = (int)((cur / (128 * 128)) % 128), cl = (int)(cur / GRID_CELLS);
            score += 16 - (1 << cl); ++size;
            int k = neighbors(cells, cx, cy, cz, cl, nb);
            for (int i = 0; i < k; ++i) {
                size_t ni = cell_idx(nb[i][0], nb[i][1], nb[i][2], nb[i][3]);
                if (label[ni] < 0) { label[ni] = n_clusters; queue[tail++] = (uint32_t)ni; }
            }
        }
        if (best < 0 || score > best_score) { best = n_clusters; best_score = score; best_size = size; }
        ++n_clusters;
    }
    if (best >= 0) {
        uint8_t* out = (uint8_t*)calloc(total, 1);
        for (size_t i = 0; i < total; ++i) {
            if (label[i] != best) continue;
            int x = (int)(i % 128), y = (int)((i / 128) % 128), z = (int)((i / (128 * 128)) % 128), l = (int)(i / GRID_CELLS);
            out[i] = 1;
            for (int lvl = l + 1; lvl < 8; ++lvl) { x = 32 + x / 2; y = 32 + y / 2; z = 32 + z / 2; out[cell_idx(x, y, z, lvl)] = 1; }
        }
        memcpy(cells, out, total);
        free(out);
    }
    if (out_best_size) *out_best_size = best_size;
    if (out_best_score) *out_best_score = best_score;
    free(label); free(queue);
    return best >= 0 ? n_clusters : -1;
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm of bench.py runs on rank 0 alone and may use the host */
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
