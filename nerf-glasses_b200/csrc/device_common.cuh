// device_common.cuh - device-side building blocks shared by every kernel of libnmr (sm_100a).
//
// Arithmetic contract (DESIGN.md "Numerics"): this translation unit is compiled with --fmad=false, so every
// fp32 product and sum is rounded separately; the summation orders below are the ones the reference's Eigen/glm
// expressions evaluate to (pinned on the host against the reference headers, tests/golden/ref_vectors.npz).  That
// makes ray set-up, occupancy traversal (t, Morton cell, mip) and the fp16 hash-grid features reproducible to the bit.
#pragma once
#ifdef NMR_CHECKED
#include <cstdio>
#endif
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmr {

constexpr uint32_t NERF_GRIDSIZE = 128;
constexpr uint32_t NERF_CASCADES = 8;
constexpr uint32_t GRID_CELLS = NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE;
constexpr int N_LEVELS = 16;
constexpr int ENC_WIDTH = 32;

// ---- plain data handed to kernels by value ------------------------------------------------------------------
// Checked build (-DNMR_CHECKED, tools/all_paths.py): every index that reaches global memory is tested against the size of
// its buffer and a violation stops the kernel with file:line.  compute-sanitizer is not available on the B200 pool; this build
// is its stand-in for the index arithmetic of the traversal, the encoder, the ray queue and the frame writes.  Off: no code.
#ifdef NMR_CHECKED
#define NMR_DEVICE_CHECK(cond) do { if (!(cond)) { printf("nmr check failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define NMR_DEVICE_CHECK(cond) ((void)0)
#endif

struct DeviceModel {
    const __half2* grid;              // hash-table entries, all levels back to back
    const __half* mlp;                // density net | rgb net, row-major [out][in] (params_binary order)
    const __half* mlp_tc;             // the same 10 240 weights in the tensor-core operand layout (kernels.cu: stage_weights_tc), or null
    const uint8_t* bitfield;          // 2 MiB occupancy bits, Morton order per cascade
    const uint32_t* coarse;           // 32^3 "near" bits of cascade 0 (kCoarseRes; see coarse_near), or null: no empty-space jumps
    const __half2* level_ptr[N_LEVELS];  // grid + level_offset[l]: 64-bit base per level, so a gather is base + 32-bit index
    uint32_t level_offset[N_LEVELS];  // first entry of each level
    uint32_t level_size[N_LEVELS];    // entries in each level
    float level_scale[N_LEVELS];
    uint32_t stride_y[N_LEVELS], stride_z[N_LEVELS];
    uint32_t dense_mask;              // bit l set: level l is indexed densely
    uint32_t pow2_mask;               // bit l set: level_size[l] is a power of two (index & (size - 1))
    uint32_t prime[3];                // hash multipliers of the snapshot's hash type (CoherentPrime: 1, 2654435761, 805459861)
    // Levels [0, n_brick) re-laid out at load time as 2x2x2 BRICKS (brick_build_kernel): cell (gx, gy, gz) of a level of `brick_res`
    // cells per axis owns 32 bytes = the eight corner entries of that cell in corner order, i.e. exactly the values the plain
    // layout returns for the indices of grid.h:317-343 (index wrap included).  One 32-byte sector per sample and level instead of
    // eight 4-byte gathers with eight index computations; values and blend order unchanged, so the features stay bit-identical.
    const uint4* brick[N_LEVELS];
    uint32_t brick_res[N_LEVELS];
    uint32_t n_brick;
};

struct TexDev { const float* p; int w, h; };     // [h][w][4] floats (sRGB textures already linearised), p == nullptr: no texture
struct MeshDevice {
    const float* wpos;      // [n_verts][3] world space
    const float* wnrm;      // [n_verts][3]
    const float* uv;        // [n_verts][2]
    const uint32_t* idx;    // [n_tris][3]
    const float* tex_lin;   // [h][w][4] linearised base colour texture or nullptr
    const uint8_t* tri_lens; // [n_tris] 1 = lens surface (transmissive material), or nullptr when the scene has none / lenses are off
    uint32_t n_tris;
    int tex_w, tex_h;
    float base_color[4], emissive[3], metallic, roughness;
    // the other textures of the reference's closest-hit program (S/optix/optix_scene.cu:234-258)
    TexDev tex_emissive, tex_mr, tex_normal, tex_occ;
    float normal_scale, occlusion_strength;
    const float* wtbn;      // [n_verts][8]: M3 n_obj (3), M3 t_obj (3), tangent handedness, 0 - M3 = R S of the mesh; only with a normal texture
    float nmat[9];          // normal matrix R S^-1 of the first mesh, row-major (applied once more to the mapped normal, like the reference)
};

struct FrameParams {
    int width, height;
    float cam[12];                    // column-major 3x4
    float aabb_min[3], aabb_max[3];   // render aabb
    float r2l[9];                     // render_aabb_to_local, row-major
    float taabb_min[3], taabb_max[3]; // training aabb
    float cone_angle;
    uint32_t spp_index;
    float min_transmittance;
    int rgb_activation, density_activation;
    float background[4];
    float background_linear[3];       // srgb_to_linear(background), evaluated on the host
    float background_out[4];          // displayed value of a pixel that hit nothing (accumulate + tonemap of zero), see finish_pixel
    int to_srgb;
    int tonemap_curve;                // Testbed.tonemap_curve: 0 Identity, 1 ACES, 2 Hable, 3 Reinhard (S/ngp/render_buffer.cu:269-325)
    int shard_rank, shard_world, shard_band;
    int row0;                         // unsharded contexts: first image row of this pass (nmr_render's row ranges), normally 0
    int bg_filled_elsewhere;          // shared frame target (nmr_gather_*): pixels outside both screen rectangles are written by the destination rank's fill kernel, not by this pass
    int mesh_scale;                   // 0: no mesh stage
    float light[3];
    float cam_inv[9];                 // inverse of [U V W] (row-major), for the rasteriser's bounding boxes
    float occ_min[3], occ_max[3];     // box around every occupied grid cell a sample can test, inflated by one cell
    int occ_px[4];                    // pixels [x0, x1) x [y0, y1) = occ_px[0..3] whose ray can meet that box (host-projected, padded); the rest is background unless the mesh covers it
    int surface_mode;                 // where a partially covering mesh surface enters the compositing order: SurfaceMode
    // The mesh visibility buffer covers only the screen bounding box of the mesh (sub-pixel units of the mesh_scale x
    // supersampled frame, aligned to whole pixels): zb_w == 0 means "mesh not in view".  Entry (x, y) lives at
    // (y - zb_y0) * zb_w + (x - zb_x0).
    int zb_x0, zb_y0, zb_w, zb_h;
    // Lens surfaces (thin dielectric sheet, DESIGN.md "Secondary rays"): Schlick F0, transmitted colour factor
    // transmission * tint and its mean.  lens_on == 0: no lens triangles in this frame.
    int lens_on;
    float lens_f0, lens_k[3], lens_kmean;
    int lens_model;                   // 0 thin sheet (no bending), 1 plate of lens_thickness with parallel faces (two-interface Snell)
    float lens_thickness, lens_ior;
    int out_format;                   // PixelFormat of FrameOut::image (nmr_pixel_format of include/nmr.h)
    // Testbed::m_model_rotation / m_model_translation (S/ngp/testbed.cu:1537-1542, consumed at :442-446): ray direction = R d,
    // NeRF-space ray origin = R eye + 0.5 + R t_model (evaluated on the host, make_params).  The identity leaves eye + 0.5.
    float model_rot[9];               // row-major
    float ray_origin[3];
    int debug_pixel;                  // checked builds (-DNMR_CHECKED): the march kernel narrates this pixel's ray (NMR_DEBUG_PIXEL), -1: none
};

// Displayed image formats.  kPixelU8 is what render.py turns every frame into on the host (np.uint8(img * 255), V/render.py:62-66):
// trunc(clamp(v, 0, 1) * 255) with the product rounded to fp32 first, like numpy's float32 arithmetic followed by the C cast.
enum PixelFormat : int { kPixelF32 = 0, kPixelF16 = 1, kPixelU8 = 2 };
__device__ __forceinline__ void store_pixel(void* __restrict__ image, int fmt, uint32_t idx, float r, float g, float b, float a) {
    if (fmt == kPixelF32) {
        reinterpret_cast<float4*>(image)[idx] = make_float4(r, g, b, a);
    } else if (fmt == kPixelU8) {
        const uint32_t ur = __float2uint_rz(__saturatef(r) * 255.0f), ug = __float2uint_rz(__saturatef(g) * 255.0f);
        const uint32_t ub = __float2uint_rz(__saturatef(b) * 255.0f), ua = __float2uint_rz(__saturatef(a) * 255.0f);
        reinterpret_cast<uint32_t*>(image)[idx] = ur | (ug << 8) | (ub << 16) | (ua << 24);
    } else {
        const __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, a);
        reinterpret_cast<uint2*>(image)[idx] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
}

// The reference blends the mesh surface in front of the first sample of the n_steps BATCH whose end passed t_surface
// (composite_kernel_nerf tests payload.t, which generate_next_nerf_network_inputs has already advanced past the whole
// batch, S/ngp/testbed.cu:620-633, 843), and n_steps = clamp(pixels / live rays, 1, 8) (S/ngp/testbed.cu:1996).  While at
// most 1/8 of the pixels hold a live ray - render.py's framing - that is a fixed 8-sample batch counted from the ray's
// first sample, a ray-local rule (kSurfaceBatch8).  kSurfaceAuto picks it under that condition and the exact per-sample
// position otherwise; kSurfaceExact always inserts at the exact sample.
enum SurfaceMode : int { kSurfaceAuto = 0, kSurfaceExact = 1, kSurfaceBatch8 = 2 };

// ---- tiny vector helpers -------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 vmul(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 vcross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// glm::dot - left to right
__device__ __forceinline__ float gdot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// Eigen 3-element redux - e0 + (e1 + e2)
__device__ __forceinline__ float edot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
__device__ __forceinline__ V3 gnormalize(V3 a) { const float inv = 1.0f / sqrtf(gdot(a, a)); return vmul(a, inv); }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }

// ---- Morton code (T/include/tiny-cuda-nn/common_device.h:338-353) ---------------------------------------------
__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
__host__ __device__ __forceinline__ uint32_t morton3D_invert(uint32_t x) {
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

// ---- start jitter (S/ngp/random_val.cuh:163-294; Sobol dimension 0 is a bit reversal) -------------------------
__device__ __forceinline__ uint32_t lk_permute(uint32_t x, uint32_t seed) {
    x += seed;
    x ^= x * 0x6c50b47cu;
    x ^= x * 0xb82f1e52u;
    x ^= x * 0xc7afe638u;
    x ^= x * 0x8d22f6e6u;
    return x;
}
__device__ __forceinline__ uint32_t nested_scramble(uint32_t x, uint32_t seed) { return __brev(lk_permute(__brev(x), seed)); }
__device__ __forceinline__ float ld_random_val(uint32_t index, uint32_t seed) {
    index = nested_scramble(index, seed);
    const uint32_t s0 = seed ^ (0u + (seed << 6) + (seed >> 2));   // hash_combine(seed, 0)
    return (float)nested_scramble(__brev(index), s0) * 2.3283064365386963e-10f;
}

// ---- step sizes (S/ngp/testbed.cu:177-232) -------------------------------------------------------------------
#define NMR_SQRT3 1.73205080757f
__device__ __forceinline__ float min_cone_stepsize() { return NMR_SQRT3 / 1024.0f; }
__device__ __forceinline__ float max_cone_stepsize() { return (NMR_SQRT3 / 1024.0f) * 128.0f * 1024.0f / 128.0f; }
__device__ __forceinline__ float calc_dt(float t, float cone_angle) { return clampf(t * cone_angle, min_cone_stepsize(), max_cone_stepsize()); }
__device__ __forceinline__ float warp_dt(float dt) {
    const float max_stepsize = min_cone_stepsize() * 128.0f;
    return (dt - min_cone_stepsize()) / (max_stepsize - min_cone_stepsize());
}
__device__ __forceinline__ float unwarp_dt(float dt) {
    const float max_stepsize = min_cone_stepsize() * 128.0f;
    return dt * (max_stepsize - min_cone_stepsize()) + min_cone_stepsize();
}

// ---- occupancy-grid DDA (S/ngp/testbed.cu:188-202, 234-264, 293-315) -----------------------------------------
// frexpf exponent: exact, with the bit trick for normal numbers
__device__ __forceinline__ int frexp_exponent(float v) {
    const uint32_t b = __float_as_uint(v) & 0x7FFFFFFFu;
    if (b == 0u) return 0;
    const uint32_t e = b >> 23;
    if (e == 0u || e == 255u) { int ex; frexpf(v, &ex); return ex; }
    return (int)e - 126;
}
__device__ __forceinline__ int mip_from_pos(V3 pos) {
    const float maxval = fmaxf(fmaxf(fabsf(pos.x - 0.5f), fabsf(pos.y - 0.5f)), fabsf(pos.z - 0.5f));
    const int e = frexp_exponent(maxval) + 1;
    return min((int)NERF_CASCADES - 1, max(0, e));
}
__device__ __forceinline__ int mip_from_dt(float dt, V3 pos) {
    const int mip = mip_from_pos(pos);
    dt *= 2 * NERF_GRIDSIZE;
    if (dt < 1.f) return mip;
    return min((int)NERF_CASCADES - 1, max(frexp_exponent(dt), mip));
}
__device__ __forceinline__ uint32_t cascaded_grid_idx_at(V3 pos, uint32_t mip) {
    const float mip_scale = __uint_as_float((127u - mip) << 23);   // scalbnf(1.0f, -mip)
    pos.x -= 0.5f; pos.y -= 0.5f; pos.z -= 0.5f;
    pos.x *= mip_scale; pos.y *= mip_scale; pos.z *= mip_scale;
    pos.x += 0.5f; pos.y += 0.5f; pos.z += 0.5f;
    const int ix = min(max(__float2int_rz(pos.x * (float)NERF_GRIDSIZE), 0), 127);
    const int iy = min(max(__float2int_rz(pos.y * (float)NERF_GRIDSIZE), 0), 127);
    const int iz = min(max(__float2int_rz(pos.z * (float)NERF_GRIDSIZE), 0), 127);
    return morton3D((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
}
__device__ __forceinline__ bool occupied_at(V3 pos, const uint8_t* __restrict__ bitfield, uint32_t mip, uint32_t* cell_out = nullptr) {
    const uint32_t idx = cascaded_grid_idx_at(pos, mip);
    if (cell_out) *cell_out = idx;
    NMR_DEVICE_CHECK(idx < GRID_CELLS && mip < 8u);
    return __ldg(bitfield + idx / 8 + (GRID_CELLS / 8) * mip) & (1u << (idx % 8));
}
__device__ __forceinline__ float distance_to_next_voxel(V3 pos, V3 dir, V3 idir, uint32_t res) {
    const float r = (float)res;
    const V3 p = v3(r * pos.x, r * pos.y, r * pos.z);
    const float tx = (floorf(p.x + 0.5f + 0.5f * copysignf(1.0f, dir.x)) - p.x) * idir.x;
    const float ty = (floorf(p.y + 0.5f + 0.5f * copysignf(1.0f, dir.y)) - p.y) * idir.y;
    const float tz = (floorf(p.z + 0.5f + 0.5f * copysignf(1.0f, dir.z)) - p.z) * idir.z;
    const float t = fminf(fminf(tx, ty), tz);
    // res is a power of two (128 >> mip): multiplying by its reciprocal is the division, exactly (the quotient of a normal
    // float by 2^k is never inexact above the subnormal range, and t / r is then flushed to >= 0 by the max anyway)
    return fmaxf(t * __uint_as_float((254u << 23) - __float_as_uint(r)), 0.0f);
}
// `do { t += dt0; } while (t < t_target);` of the uniform-step walk, landed on directly.  While t and t_target share a binade the
// k-th addition yields the bit pattern bits(t) + k * inc (see lattice_advance below), and positive floats order like their bit
// patterns, so the loop ends at the smallest k >= 1 with bits(t) + k * inc >= bits(t_target): one rounded-down float quotient,
// corrected upwards.  Anything else (another binade, a landing beyond the binade's top, an increment of zero, a NaN) takes the
// additions themselves.  Bit-identical to the loop (tests/test_host_cpu.py restates it; the GPU traversal tests pass with it) but SLOWER on the B200 - the
// extra live values cost the march kernel 4 % and the set-up kernel as much (profiles/r2_ab_direct_landing.txt) - so it is only
// compiled with -DNMR_DIRECT_LANDING.
__device__ __forceinline__ float uniform_steps_past(float t, float t_target) {
    const float dt0 = min_cone_stepsize();
#ifdef NMR_DIRECT_LANDING
    const uint32_t b = __float_as_uint(t), e = b & 0xFF800000u, bt = __float_as_uint(t_target);
    const uint32_t inc = __float_as_uint(__uint_as_float(e) + dt0) - e;
    if ((bt & 0xFF800000u) == e && inc - 1u < 0x00400000u) {            // same binade, 0 < inc <= 2^22
        uint32_t k = 1u;
        if (bt > b) {
            const uint32_t diff = bt - b;                                // < 2^23
            k = (uint32_t)(__uint2float_rz(diff) * __frcp_rz(__uint2float_ru(inc)));      // <= diff / inc
            if (k * inc < diff) ++k;
            if (k * inc < diff) ++k;
        }
        const uint32_t r = b + k * inc;
        if (k * inc >= bt - b && (k == 1u || (k - 1u) * inc < bt - b) && r <= (e | 0x007FFFFFu)) return __uint_as_float(r);
    }
#endif
    do { t += dt0; } while (t < t_target);
    return t;
}
// uniform_dt: the cone angle is zero, so calc_dt(t, 0) = clamp(t * 0) is the constant minimum step for every finite t
__device__ __forceinline__ float advance_to_next_voxel(float t, float cone_angle, bool uniform_dt, V3 pos, V3 dir, V3 idir, uint32_t res) {
    const float t_target = t + distance_to_next_voxel(pos, dir, idir, res);
    if (uniform_dt) t = uniform_steps_past(t, t_target);
    else { do { t += calc_dt(t, cone_angle); } while (t < t_target); }
    return t;
}

// ---- exact empty-space jumps of the uniform-step walk ---------------------------------------------------------
// With a zero cone angle every step of the reference's walk is `t += dt0` in fp32 (S/ngp/testbed.cu:306-313).  Inside one
// binade [2^e, 2^(e+1)) t is a multiple of ulp = 2^(e-23) and dt0 = q ulp + r with a fixed r, so every such addition rounds
// the same way: it advances the BIT PATTERN of t by the constant inc(e) = bits(2^e + dt0) - bits(2^e) (r is never the tie
// ulp/2 for dt0 = sqrt(3)/1024 and any binade a walk can reach - checked in tests/test_host_cpu.py).  K sequential additions
// are therefore one integer multiply-add while the result stays inside the binade; a binade boundary is crossed with real
// additions.  The result is bit-identical to the K-fold loop.
__device__ __forceinline__ float lattice_advance(float t, int K) {
    const float dt0 = min_cone_stepsize();
    while (K > 0) {
        const uint32_t b = __float_as_uint(t);
        const uint32_t e = b & 0xFF800000u;
        const uint32_t inc = __float_as_uint(__uint_as_float(e) + dt0) - e;
        const uint32_t room = (e | 0x007FFFFFu) - b;                              // additions of one ulp that stay inside the binade
        // steps that certainly fit: (float) room / inc is within a few ulp of the quotient; two steps of slack
        const int fit = (int)(__uint2float_rz(room) * __frcp_rz(__uint2float_ru(inc))) - 2;
        if (K <= fit) return __uint_as_float(b + (uint32_t)K * inc);
        if (fit > 0) { t = __uint_as_float(b + (uint32_t)fit * inc); K -= fit; }
        // close to the top of the binade: real additions until the exponent changes (at most a handful)
        do { t += dt0; --K; } while (K > 0 && (__float_as_uint(t) & 0xFF800000u) == e);
    }
    return t;
}

// Coarse view of cascade 0 for those jumps: kCoarseRes^3 cells of (128 / kCoarseRes)^3 grid cells each.  A coarse cell is
// NEAR when a cell within Chebyshev distance 2 of it holds an occupied grid cell, or when it touches the surface of the unit
// cube (positions there may test cascade 1, and the walk's box tests stay exact).  From a cell that is not near, the walk
// can only step through empty cascade-0 cells until it leaves the coarse cell, so those steps are taken at once: the walk
// lands on the first step at or behind the coarse cell's exit, exactly where its own last step inside the cell would land
// up to the rounding of the two distance computations (they differ by ~1e-7, a step is 1.7e-3: the landing differs with
// probability ~1e-4, by one step) - and from any landing point the walk falls back onto the reference's own sequence of
// steps at the next voxel boundary with the same odds.  Two coarse cells (>= 4 voxel steps) separate a landing from the first
// occupied cell, which puts a different first sample at ~1e-16 per ray; tests compare t bit for bit with the oracle's plain walk.
#ifndef NMR_COARSE_RES
#define NMR_COARSE_RES 32
#endif
constexpr int kCoarseRes = NMR_COARSE_RES;               // 32: cells of 4 voxels, reach 2 cells; 64: cells of 2 voxels, reach 3 cells
constexpr int kCoarseReach = kCoarseRes == 32 ? 2 : 3;   // Chebyshev distance (in coarse cells) within which an occupied cell makes a cell "near"
constexpr int kCoarseRowWords = kCoarseRes / 32;
__device__ __forceinline__ bool coarse_near(const uint32_t* __restrict__ coarse, V3 pos) {
    const int cx = min(max(__float2int_rz(pos.x * (float)kCoarseRes), 0), kCoarseRes - 1);
    const int cy = min(max(__float2int_rz(pos.y * (float)kCoarseRes), 0), kCoarseRes - 1);
    const int cz = min(max(__float2int_rz(pos.z * (float)kCoarseRes), 0), kCoarseRes - 1);
    return (__ldg(coarse + (cz * kCoarseRes + cy) * kCoarseRowWords + (cx >> 5)) >> (cx & 31)) & 1u;
}
// the walk's steps up to the first one at or behind the exit of the (empty) coarse cell around pos
__device__ __forceinline__ float coarse_skip(float t, V3 pos, V3 dir, V3 idir) {
    const float dt0 = min_cone_stepsize();
    const float t_target = t + distance_to_next_voxel(pos, dir, idir, (uint32_t)kCoarseRes);
    const int K = (int)((t_target - t) * (1.0f / dt0)) - 2;          // steps that certainly stay in front of t_target
    if (K > 0) t = lattice_advance(t, K);
    do { t += dt0; } while (t < t_target);
    return t;
}

// ---- boxes (S/ngp/bounding_box.cuh:106-167) -------------------------------------------------------------------
__device__ __forceinline__ bool box_contains(const float* mn, const float* mx, V3 p) {
    return p.x >= mn[0] && p.x <= mx[0] && p.y >= mn[1] && p.y <= mx[1] && p.z >= mn[2] && p.z <= mx[2];
}
__device__ __forceinline__ float box_ray_tmin(const float* mn, const float* mx, V3 pos, V3 dir) {
    const float FMAXV = 3.402823466e+38f;
    float tmin = (mn[0] - pos.x) / dir.x, tmax = (mx[0] - pos.x) / dir.x;
    if (tmin > tmax) { const float c = tmin; tmin = tmax; tmax = c; }
    float tymin = (mn[1] - pos.y) / dir.y, tymax = (mx[1] - pos.y) / dir.y;
    if (tymin > tymax) { const float c = tymin; tymin = tymax; tymax = c; }
    if (tmin > tymax || tymin > tmax) return FMAXV;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (mn[2] - pos.z) / dir.z, tzmax = (mx[2] - pos.z) / dir.z;
    if (tzmin > tzmax) { const float c = tzmin; tzmin = tzmax; tzmax = c; }
    if (tmin > tzmax || tzmin > tmax) return FMAXV;
    if (tzmin > tmin) tmin = tzmin;
    return tmin;
}
__device__ __forceinline__ V3 r2l_mul(const float* m, V3 p) {   // Eigen Matrix3f * Vector3f
    // identity (the only value reachable without a custom snapshot): x*1 + (y*0 + z*0) == x exactly for finite inputs
    if (m[0] == 1.f && m[4] == 1.f && m[8] == 1.f && m[1] == 0.f && m[2] == 0.f && m[3] == 0.f && m[5] == 0.f && m[6] == 0.f && m[7] == 0.f) return p;
    return v3(m[0] * p.x + (m[1] * p.y + m[2] * p.z), m[3] * p.x + (m[4] * p.y + m[5] * p.z), m[6] * p.x + (m[7] * p.y + m[8] * p.z));
}

// ---- ray set-up (S/ngp/ngp_common.cuh:362-368; S/ngp/testbed.cu:435-464) ---------------------------------------
struct RayInit { V3 origin, dir; float t; float t_occ_in, t_limit; bool alive; };   // [t_occ_in, t_limit]: ray inside the box around the occupied cells

// Far intersection of the ray with the box around all occupied cells (negative when the ray misses it).  Past that
// parameter no sample can land in an occupied cell, so a walk may stop there: the reference's own walk would only step
// through empty cells until it leaves the render box (S/ngp/testbed.cu:506-532, 600-625) - same result, no arithmetic on t.
__device__ __forceinline__ float occupied_exit(const FrameParams& P, V3 o, V3 d, float& t_in) {
    float tmin = -3.402823466e+38f, tmax = 3.402823466e+38f;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    t_in = 3.402823466e+38f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (dd[k] != 0.f) {
            float a = (P.occ_min[k] - oo[k]) / dd[k], b = (P.occ_max[k] - oo[k]) / dd[k];
            if (a > b) { const float c = a; a = b; b = c; }
            tmin = fmaxf(tmin, a); tmax = fminf(tmax, b);
        } else if (oo[k] < P.occ_min[k] || oo[k] > P.occ_max[k]) {
            return -1.f;
        }
    }
    if (!(tmin <= tmax)) return -1.f;
    t_in = tmin;
    return tmax;
}

// R d (glm-style left-to-right sums); the identity - the only value reachable without nmr_set_model_transform - returns d itself
__device__ __forceinline__ V3 model_rotate(const FrameParams& P, V3 d) {
    const float* m = P.model_rot;
    if (m[0] == 1.f && m[4] == 1.f && m[8] == 1.f && m[1] == 0.f && m[2] == 0.f && m[3] == 0.f && m[5] == 0.f && m[6] == 0.f && m[7] == 0.f) return d;
    return v3((m[0] * d.x + m[1] * d.y) + m[2] * d.z, (m[3] * d.x + m[4] * d.y) + m[5] * d.z, (m[6] * d.x + m[7] * d.y) + m[8] * d.z);
}

__device__ __forceinline__ RayInit init_ray(const FrameParams& P, uint32_t x, uint32_t y) {
    const float* c = P.cam;
    const float ux = 2.0f * (((float)x + 0.5f) / (float)P.width) - 1.0f;
    const float uy = 2.0f * (((float)y + 0.5f) / (float)P.height) - 1.0f;
    V3 d = v3(c[0] * ux + (c[3] * uy + c[6] * 1.0f), c[1] * ux + (c[4] * uy + c[7] * 1.0f), c[2] * ux + (c[5] * uy + c[8] * 1.0f));
    const float z = edot(d, d);
    if (z > 0.0f) { const float n = sqrtf(z); d = v3(d.x / n, d.y / n, d.z / n); }
    RayInit r;
    r.origin = v3(P.ray_origin[0], P.ray_origin[1], P.ray_origin[2]);
    d = model_rotate(P, d);
    r.dir = d;
    r.t = fmaxf(box_ray_tmin(P.aabb_min, P.aabb_max, r.origin, d), 0.0f) + 1e-6f;
    r.alive = box_contains(P.aabb_min, P.aabb_max, vadd(r.origin, vmul(d, r.t)));
    r.t_limit = occupied_exit(P, r.origin, d, r.t_occ_in);
    // A walk gets nowhere once a step no longer changes t (t beyond ~4e4 with the scene in the unit cube: a camera in the wrong
    // units) and the reference's loops would spin for ever; such a ray - and one whose set-up produced a NaN - sees nothing.
    if (!(r.t_limit + min_cone_stepsize() > r.t_limit)) { r.alive = false; r.t_limit = -1.f; }
    return r;
}

// advance_pos_nerf (S/ngp/testbed.cu:470-537); returns alive
__device__ __forceinline__ bool advance_pos(const FrameParams& P, const uint8_t* __restrict__ bitfield, const uint32_t* __restrict__ coarse, V3 origin, V3 dir, uint32_t pixel_idx,
                                            float t_surface, float t_occ_in, float t_limit, bool alive, float& t_io, float& t_start) {
    t_start = 0.f;
    if (!alive) {
        if (t_surface != 0.0f) { t_io = t_surface; return true; }
        return false;
    }
    const V3 idir = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    const float cone = P.cone_angle;
    float t = t_io;
    float dt = calc_dt(t, cone);
    t += ld_random_val(P.spp_index, pixel_idx * 786433u) * dt;
    const bool uniform_dt = cone == 0.0f;
    if (uniform_dt) {
        // Zero cone angle: where the walk goes next does not depend on what the occupancy tests return, only whether it goes
        // on.  So the next kWalkBatch steps are laid out first (arithmetic only) with their occupancy loads in flight together,
        // and are then judged in order - the walk's critical path holds one memory latency per batch instead of one per step.
#ifndef NMR_WALK_BATCH
#define NMR_WALK_BATCH 4
#endif
        constexpr int kWalkBatch = NMR_WALK_BATCH;      // 4 measured best (profiles/r2_experiments.md)
        const float dt0 = min_cone_stepsize();
        bool hit = false;
        while (!hit) {
            float tv[kWalkBatch];
            uint32_t word[kWalkBatch], bit[kWalkBatch];
            int n = 0, stop = 0;                 // stop: 1 behind the mesh surface, 2 out of the box, 3 coarse cell without anything near
            float tc = t;
#pragma unroll
            for (int k = 0; k < kWalkBatch; ++k) {
                if (stop != 0) continue;
                if (t_surface != 0.0f && tc > t_surface) { stop = 1; continue; }
                const V3 pc = vadd(origin, vmul(dir, tc));
                if (tc > t_limit || !box_contains(P.aabb_min, P.aabb_max, r2l_mul(P.r2l, pc))) { stop = 2; continue; }
                if (coarse && !coarse_near(coarse, pc)) { stop = 3; continue; }
                const uint32_t mip = (uint32_t)mip_from_dt(dt0, pc);
                const uint32_t idx = cascaded_grid_idx_at(pc, mip);
                // before the ray enters the box around the occupied cells every test is known to fail: no load
                word[k] = tc >= t_occ_in ? (uint32_t)__ldg(bitfield + idx / 8 + (GRID_CELLS / 8) * mip) : 0u;
                bit[k] = 1u << (idx % 8);
                tv[k] = tc;
                n = k + 1;
                tc = advance_to_next_voxel(tc, cone, true, pc, dir, idir, NERF_GRIDSIZE >> mip);
            }
#pragma unroll
            for (int k = 0; k < kWalkBatch; ++k) {
                if (!hit && k < n && (word[k] & bit[k])) { t = tv[k]; hit = true; }
            }
            if (hit) break;
            t = tc;
            if (stop == 1) { t_io = t_surface; return true; }
            if (stop == 2) {                     // nothing occupied ahead == walked out of the box
                if (t_surface != 0.0f) { t_io = t_surface; return true; }
                alive = false;
                break;
            }
            if (stop == 3) t = coarse_skip(t, vadd(origin, vmul(dir, t)), dir, idir);
        }
    } else
    while (true) {
        if (t_surface != 0.0f && t > t_surface) { t_io = t_surface; return true; }
        const V3 pos = vadd(origin, vmul(dir, t));
        if (t > t_limit || !box_contains(P.aabb_min, P.aabb_max, r2l_mul(P.r2l, pos))) {   // nothing occupied ahead == walked out of the box
            if (t_surface != 0.0f) { t_io = t_surface; return true; }
            alive = false;
            break;
        }
        dt = calc_dt(t, cone);
        const uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
        // before the ray enters the box around the occupied cells every test is known to fail: no load, no Morton code
        if (t >= t_occ_in && occupied_at(pos, bitfield, mip)) break;
        t = advance_to_next_voxel(t, cone, false, pos, dir, idir, NERF_GRIDSIZE >> mip);
    }
    t_io = t;
    if (mip_from_pos(vadd(origin, vmul(dir, t))) == 0) t_start = t;
    return alive;
}

// One step of generate_next_nerf_network_inputs (S/ngp/testbed.cu:564-633) with n_steps = 1.
// Returns 1 with the sample (warped position, warped dt) and t advanced past it, or 0 when the ray produced no sample
// (left the box, or reached an opaque mesh surface, in which case t is snapped to t_surface).
struct Sample { V3 pos; float dt_warped; float t; uint32_t cell, mip; };
// Returns 1 with the sample and t advanced past it; 0 when the ray produced no sample (left the box, or reached an opaque
// mesh surface, in which case t is snapped to t_surface); 2 when `budget` voxel tests were spent in empty space without
// either outcome - t_io then holds the walk's state and the same call resumes it (the walk is a pure function of t).
__device__ __forceinline__ int next_sample(const FrameParams& P, const uint8_t* __restrict__ bitfield, V3 origin, V3 dir, V3 idir,
                                        float t_start, float t_surface, float surf_w, float t_limit, bool ignore_surface, int budget, float& t_io, Sample& s) {
    const float cone = P.cone_angle;
    const bool uniform_dt = cone == 0.0f;
    float t = t_io;
    V3 pos; float dt; uint32_t mip, cell;
    while (true) {
        if (!ignore_surface && t_surface != 0.0f && t > t_surface && surf_w == 1.f) { t_io = t_surface; return 0; }
        if (t > t_limit) { t_io = t; return 0; }   // no occupied cell ahead: the walk could only run out of the render box
        pos = vadd(origin, vmul(dir, t));
        if (!box_contains(P.aabb_min, P.aabb_max, r2l_mul(P.r2l, pos))) { t_io = t; return 0; }
        dt = uniform_dt ? min_cone_stepsize() : calc_dt(t - t_start, cone);
        mip = (uint32_t)mip_from_dt(dt, pos);
        if (occupied_at(pos, bitfield, mip, &cell)) break;
        t = advance_to_next_voxel(t, cone, uniform_dt, pos, dir, idir, NERF_GRIDSIZE >> mip);
        if (--budget <= 0) { t_io = t; return 2; }
    }
    const V3 diag = v3(P.taabb_max[0] - P.taabb_min[0], P.taabb_max[1] - P.taabb_min[1], P.taabb_max[2] - P.taabb_min[2]);
    s.pos = v3((pos.x - P.taabb_min[0]) / diag.x, (pos.y - P.taabb_min[1]) / diag.y, (pos.z - P.taabb_min[2]) / diag.z);
    s.dt_warped = warp_dt(dt);
    s.t = t; s.cell = cell; s.mip = mip;
    t_io = t + dt;
    return 1;
}

// ---- multiresolution hash-grid encoding (T/.../encodings/grid.h:111-186, 219-349) ------------------------------
// Per level: cell coordinates and trilinear weights, the eight corner entries (three ways to fetch them, below), and the blend
// accumulated IN FP16 in corner order 0..7 with every product rounded to fp32 and then to fp16 first (grid.h:317-343).
struct LevelCell { uint32_t gx, gy, gz; float wx0, wx1, wy0, wy1, wz0, wz1; };
__device__ __forceinline__ LevelCell level_cell(float scale, V3 p01) {
    const float fx = p01.x * scale + 0.5f, fy = p01.y * scale + 0.5f, fz = p01.z * scale + 0.5f;
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    LevelCell c;
    c.gx = (uint32_t)(int)flx; c.gy = (uint32_t)(int)fly; c.gz = (uint32_t)(int)flz;
    c.wx1 = fx - flx; c.wy1 = fy - fly; c.wz1 = fz - flz;
    c.wx0 = 1 - c.wx1; c.wy0 = 1 - c.wy1; c.wz0 = 1 - c.wz1;
    return c;
}

// indices of the eight corners of cell (gx, gy, gz) in the plain layout of `level`: dense (x + y * stride_y + z * stride_z, mod
// size) or hashed (x*p0 ^ y*p1 ^ z*p2 with the multipliers of the snapshot's hash type, prime_hash / reversed_prime_hash)
__device__ __forceinline__ void level_indices(const DeviceModel& M, int level, uint32_t gx, uint32_t gy, uint32_t gz, uint32_t index[8]) {
    const uint32_t size = M.level_size[level];
    if ((M.dense_mask >> level) & 1u) {
        const uint32_t sy = M.stride_y[level], sz = M.stride_z[level];
        const uint32_t y0 = gy * sy, z0 = gz * sz;
        const bool pow2 = (M.pow2_mask >> level) & 1u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t i = (gx + (c & 1)) + (y0 + ((c >> 1) & 1) * sy) + (z0 + ((c >> 2) & 1) * sz);
            // index % size.  A dense level holds >= res^3 entries and coordinates never exceed res, so i < 2 * size and one
            // conditional subtraction is the exact modulo; the one exception is the uint32-wrapped stride quirk of
            // res = 2048 levels (grid.h:164-186), whose size is a power of two.
            if (pow2) i &= size - 1u; else if (i >= size) i -= size;
            index[c] = i;
        }
    } else {
        const uint32_t p0 = M.prime[0], p1 = M.prime[1], p2 = M.prime[2];
        const uint32_t hx0 = gx * p0, hy0 = gy * p1, hz0 = gz * p2;
        const uint32_t hx1 = hx0 + p0, hy1 = hy0 + p1, hz1 = hz0 + p2;
        const uint32_t mask = size - 1u;       // a hashed level always has exactly 2^log2_hashmap_size entries
        // the four (y, z) partial hashes once, then one (a ^ b) & mask per corner
        const uint32_t yz00 = hy0 ^ hz0, yz10 = hy1 ^ hz0, yz01 = hy0 ^ hz1, yz11 = hy1 ^ hz1;
        index[0] = (hx0 ^ yz00) & mask; index[1] = (hx1 ^ yz00) & mask; index[2] = (hx0 ^ yz10) & mask; index[3] = (hx1 ^ yz10) & mask;
        index[4] = (hx0 ^ yz01) & mask; index[5] = (hx1 ^ yz01) & mask; index[6] = (hx0 ^ yz11) & mask; index[7] = (hx1 ^ yz11) & mask;
    }
}

// two fp32 products in one instruction (FMUL2, sm_100): each lane is an ordinary round-to-nearest fp32 multiply, so results equal
// two separate multiplies bit for bit
__device__ __forceinline__ unsigned long long f32x2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ void split_f32x2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// acc = sum over corners c = 0..7, in that order, of fp16(w_c * v[c]) with w_c = ((1 * wx) * wy) * wz (grid.h:326-343).
// Corners 2k and 2k + 1 differ in x only: their weights are one FMUL2 pair, and so are their products per feature.
__device__ __forceinline__ __half2 blend_corners(const __half2 v[8], const LevelCell& c) {
    const unsigned long long wx = f32x2(c.wx0, c.wx1);
    const unsigned long long wxy0 = mul_f32x2(wx, f32x2(c.wy0, c.wy0)), wxy1 = mul_f32x2(wx, f32x2(c.wy1, c.wy1));
    const unsigned long long wz0 = f32x2(c.wz0, c.wz0), wz1 = f32x2(c.wz1, c.wz1);
    const unsigned long long w[4] = {mul_f32x2(wxy0, wz0), mul_f32x2(wxy1, wz0), mul_f32x2(wxy0, wz1), mul_f32x2(wxy1, wz1)};
    __half2 acc = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 a = __half22float2(v[2 * k]), b = __half22float2(v[2 * k + 1]);
        float x0, x1, y0, y1;
        split_f32x2(mul_f32x2(w[k], f32x2(a.x, b.x)), x0, x1);
        split_f32x2(mul_f32x2(w[k], f32x2(a.y, b.y)), y0, y1);
        acc = __hadd2(acc, __floats2half2_rn(x0, y0));
        acc = __hadd2(acc, __floats2half2_rn(x1, y1));
    }
    return acc;
}

// One level of one sample.  BRICKS: levels below M.n_brick fetch their eight corners with a single 32-byte load from the brick
// layout (a cell coordinate outside the brick grid - a position outside the unit cube - takes the plain path); the others, and
// every level when !BRICKS, gather them one by one from the plain layout.  One copy of the cell arithmetic and of the blend.
template <bool BRICKS>
__device__ __forceinline__ __half2 encode_level_t(const DeviceModel& M, int level, V3 p01) {
    const LevelCell c = level_cell(M.level_scale[level], p01);
    __half2 v[8];
    bool bricked = false;
    if (BRICKS && (uint32_t)level < M.n_brick) {
        const uint32_t res = M.brick_res[level];
        if (c.gx < res && c.gy < res && c.gz < res) {
            const uint4* __restrict__ b = M.brick[level] + 2u * ((c.gz * res + c.gy) * res + c.gx);
            NMR_DEVICE_CHECK(M.brick[level] != nullptr && res > 0u && res <= 256u);
            uint32_t r[8];
            asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(b));
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const __half2*>(&r[k]);
            bricked = true;
        }
    }
    if (!bricked) {
        uint32_t index[8];
        level_indices(M, level, c.gx, c.gy, c.gz, index);
        const __half2* __restrict__ grid = M.level_ptr[level];
#pragma unroll
        for (int k = 0; k < 8; ++k) { NMR_DEVICE_CHECK(index[k] < M.level_size[level]); v[k] = __ldg(grid + index[k]); }
    }
    return blend_corners(v, c);
}
__device__ __forceinline__ __half2 encode_level(const DeviceModel& M, int level, V3 p01) { return encode_level_t<false>(M, level, p01); }

// All 16 levels of one sample.  Features (2l, 2l+1) of level l go to dst + (l / 4) * chunk_stride + (l % 4) * 4, i.e. four
// 16-byte chunks of 8 features.  UNROLL levels are kept in flight per thread (8 * UNROLL gathers); the loop itself stays
// rolled so the kernel fits the instruction cache.  (Parity probes and the probe kernel; plain layout only.)
template <int UNROLL>
__device__ __forceinline__ void encode_chunks(const DeviceModel& M, V3 p01, char* dst, int chunk_stride) {
#pragma unroll 1
    for (int l0 = 0; l0 < N_LEVELS; l0 += UNROLL) {
        __half2 e[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) e[j] = encode_level(M, l0 + j, p01);
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const int l = l0 + j;
            *reinterpret_cast<__half2*>(dst + (size_t)(l >> 2) * chunk_stride + (l & 3) * 4) = e[j];
        }
    }
}

// Levels first, first + step, ... of one sample (step lanes share a sample when a warp has few of them, see march_kernel step 2);
// same destination layout as encode_chunks.  Bricked levels take the brick path.  The loop stays rolled: one copy of each path.
__device__ __forceinline__ void encode_levels_strided(const DeviceModel& M, V3 p01, char* dst, int chunk_stride, uint32_t first, uint32_t step) {
#pragma unroll 1
    for (uint32_t l = first; l < (uint32_t)N_LEVELS; l += step) {
        const __half2 e = encode_level_t<true>(M, (int)l, p01);
        *reinterpret_cast<__half2*>(dst + (size_t)(l >> 2) * chunk_stride + (l & 3u) * 4u) = e;
    }
}

// ---- SH degree 4 (T/.../encodings/spherical_harmonics.h:65-98) -------------------------------------------------
__device__ __forceinline__ void sh4(V3 d01, __half2 out[8]) {
    const float x = d01.x * 2.f - 1.f, y = d01.y * 2.f - 1.f, z = d01.z * 2.f - 1.f;
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    out[0] = __floats2half2_rn(0.28209479177387814f, -0.48860251190291987f * y);
    out[1] = __floats2half2_rn(0.48860251190291987f * z, -0.48860251190291987f * x);
    out[2] = __floats2half2_rn(1.0925484305920792f * xy, -1.0925484305920792f * yz);
    out[3] = __floats2half2_rn(0.94617469575755997f * z2 - 0.31539156525251999f, -1.0925484305920792f * xz);
    out[4] = __floats2half2_rn(0.54627421529603959f * x2 - 0.54627421529603959f * y2, 0.59004358992664352f * y * (-3.0f * x2 + y2));
    out[5] = __floats2half2_rn(2.8906114426405538f * xy * z, 0.45704579946446572f * y * (1.0f - 5.0f * z2));
    out[6] = __floats2half2_rn(0.3731763325901154f * z * (5.0f * z2 - 3.0f), 0.45704579946446572f * x * (1.0f - 5.0f * z2));
    out[7] = __floats2half2_rn(1.4453057213202769f * z * (x2 - y2), 0.59004358992664352f * x * (-x2 + 3.0f * y2));
}

// ---- activations and colour transfer ---------------------------------------------------------------------------
__device__ __forceinline__ float act_density(float v, int a) {
    switch (a) { case 0: return v; case 1: return v > 0.f ? v : 0.f; case 2: return 1.0f / (1.0f + __expf(-v)); default: return __expf(v); }
}
__device__ __forceinline__ float act_rgb(float v, int a) {
    switch (a) { case 0: return v; case 1: return v > 0.f ? v : 0.f; case 2: return 1.0f / (1.0f + __expf(-v)); default: return __expf(clampf(v, -10.f, 10.f)); }
}
// x^y for x > 0 through the SFU (ex2.approx(y * lg2.approx(x))): relative error ~1e-6, three orders of magnitude below the
// 2/255 pixel tolerance; the accurate powf expands to ~100 instructions per call and there are six calls per pixel.
__device__ __forceinline__ float fast_pow(float x, float y) { return exp2f(y * __log2f(x)); }
__device__ __forceinline__ float linear_to_srgb(float l) { return l < 0.0031308f ? 12.92f * l : 1.055f * fast_pow(l, 0.41666f) - 0.055f; }
__device__ __forceinline__ float srgb_to_linear(float s) { return s <= 0.04045f ? s / 12.92f : fast_pow((s + 0.055f) / 1.055f, 2.4f); }

// tonemap(x, curve) (S/ngp/render_buffer.cu:269-325); host and device (make_params evaluates the background pixel with it)
__host__ __device__ __forceinline__ void tonemap_curve_apply(float& r, float& g, float& b, int curve) {
    if (curve == 0) return;
    r = fmaxf(r, 0.f); g = fmaxf(g, 0.f); b = fmaxf(b, 0.f);
    float k0, k1, k2, k3, k4, k5;
    if (curve == 1) {
        k0 = 0.6f * 0.6f * 2.51f; k1 = 0.6f * 0.03f; k2 = 0.0f; k3 = 0.6f * 0.6f * 2.43f; k4 = 0.6f * 0.59f; k5 = 0.14f;
    } else if (curve == 2) {
        const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
        k0 = A * F - A * E; k1 = C * B * F - B * E; k2 = 0.0f; k3 = A * F; k4 = B * F; k5 = D * F * F;
        const float W = 11.2f;
        const float nom = k0 * (W * W) + k1 * W + k2, denom = k3 * (W * W) + k4 * W + k5;
        const float white_scale = denom / nom;
        k0 = 4.0f * k0 * white_scale; k1 = 2.0f * k1 * white_scale; k2 = k2 * white_scale; k3 = 4.0f * k3; k4 = 2.0f * k4;
    } else {
        const float Y = 0.2126f * r + (0.7152f * g + 0.0722f * b);
        const float s = 1.f / (Y + 1.0f);
        r = r * s; g = g * s; b = b * s;
        return;
    }
    float* c[3] = {&r, &g, &b};
    for (int k = 0; k < 3; ++k) {
        const float x = *c[k], sq = x * x;
        const float nom = sq * k0 + k1 * x + k2, denom = k3 * sq + k4 * x + k5;
        *c[k] = nom / denom;
    }
}

// ---- mesh stage: Moeller-Trumbore with back-face culling + the reference's PBR shading ------------------------
// (S/optix/optix_scene.cu:71-85, 182-325; OptiX's own triangle test is not available: see DESIGN.md)
__device__ __forceinline__ V3 mesh_ray_dir(const FrameParams& P, int x, int y, int W2, int H2) {
    const float dx = 2.0f * (((float)x + 0.5f) / (float)W2) - 1.0f;
    const float dy = 2.0f * (((float)y + 0.5f) / (float)H2) - 1.0f;
    const float* c = P.cam;
    return gnormalize(v3((dx * c[0] + dy * c[3]) + c[6], (dx * c[1] + dy * c[4]) + c[7], (dx * c[2] + dy * c[5]) + c[8]));
}
__device__ __forceinline__ V3 ld3(const float* p, uint32_t i) { return v3(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)); }

__device__ __forceinline__ bool ray_tri(V3 o, V3 d, V3 v0, V3 v1, V3 v2, float& t_out, float& u_out, float& v_out) {
    const V3 e1 = vsub(v1, v0), e2 = vsub(v2, v0);
    const V3 pv = vcross(d, e2);
    const float det = gdot(e1, pv);
    if (!(det > 0.0f)) return false;
    const V3 tv = vsub(o, v0);
    const float u = gdot(tv, pv);
    if (u < 0.0f || u > det) return false;
    const V3 qv = vcross(tv, e1);
    const float v = gdot(d, qv);
    if (v < 0.0f || u + v > det) return false;
    const float t = gdot(e2, qv);
    if (!(t > 0.0f)) return false;
    const float inv = 1.0f / det;
    t_out = t * inv; u_out = u * inv; v_out = v * inv;
    return true;
}

// bilinear, wrap addressing, normalised coordinates (cudaTextureDesc of S/cuda_texture.cu:19-27; float weights instead of the
// texture unit's 8-bit ones)
__device__ __forceinline__ void tex_sample(const float* __restrict__ tex, int tw, int th, float u, float v, float out[4]) {
    const float fx = u * (float)tw - 0.5f, fy = v * (float)th - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float ax = fx - flx, ay = fy - fly;
    const int x0 = (int)flx, y0 = (int)fly;
    const int xa = ((x0 % tw) + tw) % tw, xb = (((x0 + 1) % tw) + tw) % tw;
    const int ya = ((y0 % th) + th) % th, yb = (((y0 + 1) % th) + th) % th;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float t00 = __ldg(tex + ((size_t)ya * tw + xa) * 4 + k), t10 = __ldg(tex + ((size_t)ya * tw + xb) * 4 + k);
        const float t01 = __ldg(tex + ((size_t)yb * tw + xa) * 4 + k), t11 = __ldg(tex + ((size_t)yb * tw + xb) * 4 + k);
        const float top = t00 * (1.0f - ax) + t10 * ax, bot = t01 * (1.0f - ax) + t11 * ax;
        out[k] = top * (1.0f - ay) + bot * ay;
    }
}

__device__ __forceinline__ float to_srgb_mesh(float c) {   // S/optix/optix_util.cuh:23-29
    const float powed = powf(c, 1.0f / 2.4f);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}

__device__ __forceinline__ void shade_hit(const MeshDevice& M, const FrameParams& P, uint32_t tri, float bu, float bv, float hitT, V3 dir, float rgba[4]) {
    const uint32_t i0 = __ldg(M.idx + tri * 3), i1 = __ldg(M.idx + tri * 3 + 1), i2 = __ldg(M.idx + tri * 3 + 2);
    const float bw = 1.0f - bu - bv;
    const V3 n = vadd(vadd(vmul(ld3(M.wnrm, i1), bu), vmul(ld3(M.wnrm, i2), bv)), vmul(ld3(M.wnrm, i0), bw));
    const float uvx = (bu * __ldg(M.uv + i1 * 2) + bv * __ldg(M.uv + i2 * 2)) + bw * __ldg(M.uv + i0 * 2);
    const float uvy = (bu * __ldg(M.uv + i1 * 2 + 1) + bv * __ldg(M.uv + i2 * 2 + 1)) + bw * __ldg(M.uv + i0 * 2 + 1);
    float base[4] = {M.base_color[0], M.base_color[1], M.base_color[2], M.base_color[3]};
    if (M.tex_lin) { float tx[4]; tex_sample(M.tex_lin, M.tex_w, M.tex_h, uvx, uvy, tx); for (int k = 0; k < 4; ++k) base[k] *= tx[k]; }
    float emissive[3] = {M.emissive[0], M.emissive[1], M.emissive[2]};
    if (M.tex_emissive.p) { float tx[4]; tex_sample(M.tex_emissive.p, M.tex_emissive.w, M.tex_emissive.h, uvx, uvy, tx); for (int k = 0; k < 3; ++k) emissive[k] *= tx[k]; }
    float metallic = M.metallic, roughness = M.roughness, occlusion = 1.0f;
    if (M.tex_mr.p) { float tx[4]; tex_sample(M.tex_mr.p, M.tex_mr.w, M.tex_mr.h, uvx, uvy, tx); metallic *= tx[2]; roughness *= tx[1]; }   // B = metallic, G = roughness
    if (M.tex_occ.p) { float tx[4]; tex_sample(M.tex_occ.p, M.tex_occ.w, M.tex_occ.h, uvx, uvy, tx); occlusion = 1.0f + M.occlusion_strength * (tx[0] - 1.0f); }
    V3 normal = n;
    if (M.tex_normal.p && M.wtbn) {
        // computeTbnMatrix (S/optix/optix_scene.cu:92-98) of the interpolated normal / tangent under the mesh's matrix, the mapped
        // normal, and then - like the reference - the object-to-world normal transform on top (S/optix/optix_scene.cu:252-258)
        const float* a0 = M.wtbn + (size_t)i0 * 8; const float* a1 = M.wtbn + (size_t)i1 * 8; const float* a2 = M.wtbn + (size_t)i2 * 8;
        float q[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) q[k] = (bu * __ldg(a1 + k) + bv * __ldg(a2 + k)) + bw * __ldg(a0 + k);
        V3 tn = gnormalize(v3(q[3], q[4], q[5]));
        const V3 nn = gnormalize(v3(q[0], q[1], q[2]));
        tn = gnormalize(vsub(tn, vmul(nn, gdot(tn, nn))));
        const V3 bn = vmul(vcross(nn, tn), q[6]);
        float tx[4]; tex_sample(M.tex_normal.p, M.tex_normal.w, M.tex_normal.h, uvx, uvy, tx);
        const float mx = (tx[0] * 2.0f - 1.0f) * M.normal_scale, my = (tx[1] * 2.0f - 1.0f) * M.normal_scale, mz = tx[2] * 2.0f - 1.0f;
        const V3 m = vadd(vadd(vmul(tn, mx), vmul(bn, my)), vmul(nn, mz));
        normal = v3((M.nmat[0] * m.x + M.nmat[1] * m.y) + M.nmat[2] * m.z, (M.nmat[3] * m.x + M.nmat[4] * m.y) + M.nmat[5] * m.z, (M.nmat[6] * m.x + M.nmat[7] * m.y) + M.nmat[8] * m.z);
    }
    const V3 eye = v3(P.cam[9], P.cam[10], P.cam[11]);
    const V3 light = v3(P.light[0], P.light[1], P.light[2]);
    const V3 hitPos = vadd(eye, vmul(dir, hitT));
    const V3 N = gnormalize(normal);
    const V3 V = gnormalize(vsub(eye, hitPos));
    const V3 L = gnormalize(vsub(light, hitPos));
    const V3 H = gnormalize(vadd(V, L));
    const float ndl = gdot(L, N);
    const float dl = fmaxf(0.f, ndl);
    float fr[3] = {0.f, 0.f, 0.f};
    const float dotNV = gdot(N, V), dotNL = ndl;
    if (dotNV > 0 && dotNL > 0) {
        const float dotNH = clampf(gdot(N, H), 0.0f, 1.0f);
        const float dotLH = clampf(gdot(L, H), 0.0f, 1.0f);
        const float alpha = roughness * roughness;
        const float a2 = alpha * alpha;
        const float f = (dotNH * a2 - dotNH) * dotNH + 1.0f;
        const float D = a2 / (f * f);
        const float lambdaV = fmaxf(0.f, dotNL) / sqrtf(a2 + (1.0f - a2) * dotNV * dotNV);
        const float lambdaL = fmaxf(0.f, dotNV) / sqrtf(a2 + (1.0f - a2) * dotNL * dotNL);
        const float G = 0.5f / (lambdaV + lambdaL + 0.0001f);
        const float p5 = powf(1.0f - dotLH, 5.0f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float f0 = (0.5f * alpha) * (1.0f - metallic) + base[k] * metallic;
            const float F = f0 + (1.0f - f0) * p5;
            fr[k] = fabsf((D * G * F) / 3.14159265358979323846f);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float fd = (1.0f - metallic) * base[k] * dl;
        const float ambient = base[k] * .2f * occlusion;
        float c = ambient + (fd + fr[k]) + emissive[k];
        c = clampf(c, 0.f, 1.f);
        rgba[k] = to_srgb_mesh(c);
    }
    rgba[3] = 1.f;
}

}  // namespace nmr
