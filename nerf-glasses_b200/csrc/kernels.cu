// kernels.cu - the CUDA kernels of libnmr's render path (sm_100a only).
//
//   occupancy_*            snapshot density grid -> 2 MiB occupancy bitfield (load time)
//   mesh_raster_kernel     glasses mesh -> 2x supersampled visibility buffer (64-bit atomicMin of hitT|triangle)
//   init_rays_kernel       per pixel: ray set-up, mesh hand-off (shade + 2x2 resolve), first-hit DDA; dead pixels are
//                          finished in place, live rays are appended to a compact queue
//   march_kernel           persistent fused kernel: DDA sample generation, hash-grid encoding, both MLPs (tcgen05 tensor
//                          cores, accumulators in tensor memory), SH, front-to-back compositing with mesh clipping,
//                          shade/accumulate/tonemap - one thread per ray slot, 128 rays per CTA, rays refilled from the queue
//
// Reference behaviour restated per kernel: see the citations at each function and SURVEY.md section 8a.
#include "kernels.cuh"
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include <cstdio>

namespace nmr {

int rows_owned_by(int height, int rank, int world, int band) {
    if (world <= 1) return height;
    int n = 0;
    for (int y0 = rank * band; y0 < height; y0 += world * band) n += (height - y0 < band) ? (height - y0) : band;
    return n;
}

// local (owned) row -> image row
__device__ __forceinline__ int shard_row(const FrameParams& P, int local_row) {
    if (P.shard_world <= 1) return local_row + P.row0;
    const int b = local_row / P.shard_band;
    return (b * P.shard_world + P.shard_rank) * P.shard_band + local_row % P.shard_band;
}

// =================================================================================================================
// occupancy grid  (grid_to_bitfield / bitfield_max_pool, S/ngp/testbed.cu:119-166, 1120-1135)
// =================================================================================================================
__global__ void occupancy_mean_kernel(const __half* __restrict__ grid, double* __restrict__ sum) {
    // mean over cascade 0 only, each term divided by n like the reference's reduce_sum functor
    double local = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < GRID_CELLS; i += gridDim.x * blockDim.x)
        local += (double)(fmaxf(__half2float(grid[i]), 0.f) / (float)GRID_CELLS);
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(sum, local);
}

__global__ void occupancy_bits_kernel(const __half* __restrict__ grid, uint32_t n_nonzero_bytes, const double* __restrict__ sum, uint8_t* __restrict__ bitfield) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= GRID_CELLS / 8 * NERF_CASCADES) return;
    uint8_t bits = 0;
    if (i < n_nonzero_bytes) {
        const float thresh = fminf(0.01f, (float)*sum);
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) bits |= __half2float(grid[(size_t)i * 8 + j]) > thresh ? (uint8_t)(1u << j) : 0;
    }
    bitfield[i] = bits;
}

__global__ void occupancy_pool_kernel(const uint8_t* __restrict__ prev, uint8_t* __restrict__ next) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= GRID_CELLS / 64) return;
    uint8_t bits = 0;
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) bits |= prev[i * 8 + j] > 0 ? (uint8_t)(1u << j) : 0;
    const uint32_t x = morton3D_invert(i >> 0) + NERF_GRIDSIZE / 8;
    const uint32_t y = morton3D_invert(i >> 1) + NERF_GRIDSIZE / 8;
    const uint32_t z = morton3D_invert(i >> 2) + NERF_GRIDSIZE / 8;
    next[morton3D(x, y, z)] |= bits;   // distinct i map to distinct bytes: no race
}

// min / max cell coordinate of the set cells of every cascade: out[c*6 + {0,1,2}] = min xyz, [3,4,5] = max xyz
__global__ void occupancy_bounds_kernel(const uint8_t* __restrict__ bitfield, int* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;          // one byte = 8 Morton-consecutive cells = a 2x2x2 block
    if (i >= GRID_CELLS / 8 * NERF_CASCADES) return;
    const uint8_t bits = bitfield[i];
    if (!bits) return;
    const uint32_t c = i / (GRID_CELLS / 8), m = (i % (GRID_CELLS / 8)) * 8;
    const int x = (int)morton3D_invert(m), y = (int)morton3D_invert(m >> 1), z = (int)morton3D_invert(m >> 2);
    atomicMin(&out[c * 6 + 0], x); atomicMin(&out[c * 6 + 1], y); atomicMin(&out[c * 6 + 2], z);
    atomicMax(&out[c * 6 + 3], x + 1); atomicMax(&out[c * 6 + 4], y + 1); atomicMax(&out[c * 6 + 5], z + 1);
}

void launch_occupancy_bounds(const uint8_t* d_bitfield, int* d_out48, cudaStream_t s) {
    int init[48];
    for (int c = 0; c < 8; ++c) { for (int k = 0; k < 3; ++k) { init[c * 6 + k] = 1 << 20; init[c * 6 + 3 + k] = -1; } }
    cudaMemcpyAsync(d_out48, init, sizeof(init), cudaMemcpyHostToDevice, s);
    cudaStreamSynchronize(s);   // `init` lives on this stack frame
    const uint32_t total = GRID_CELLS / 8 * NERF_CASCADES;
    occupancy_bounds_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_bitfield, d_out48);
}

// coarse "near" bits of cascade 0 (device_common.cuh: coarse_near).  A coarse cell is a (128 / kCoarseRes)^3 block of grid cells
// aligned to its size, i.e. a run of consecutive Morton codes: 64 codes = 8 bytes of the bitfield (kCoarseRes 32) or 8 codes =
// one byte (kCoarseRes 64).
__global__ void coarse_occupied_kernel(const uint8_t* __restrict__ bitfield, uint8_t* __restrict__ occ) {
    constexpr uint32_t V = NERF_GRIDSIZE / kCoarseRes;
    static_assert(V == 4 || V == 2, "a coarse cell is 4 x 4 x 4 or 2 x 2 x 2 grid cells");
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kCoarseRes * kCoarseRes * kCoarseRes) return;
    const uint32_t cx = i % kCoarseRes, cy = (i / kCoarseRes) % kCoarseRes, cz = i / (kCoarseRes * kCoarseRes);
    const uint32_t m = morton3D(cx * V, cy * V, cz * V);                    // first of the block's V^3 codes
    if (V == 4) { const uint2 v = *reinterpret_cast<const uint2*>(bitfield + m / 8u); occ[i] = (v.x | v.y) != 0u; }
    else occ[i] = bitfield[m / 8u] != 0u;
}
__global__ void coarse_near_kernel(const uint8_t* __restrict__ occ, uint32_t* __restrict__ near_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;                    // one thread per coarse cell; the words were zeroed
    if (i >= kCoarseRes * kCoarseRes * kCoarseRes) return;
    const int cx = i % kCoarseRes, cy = (i / kCoarseRes) % kCoarseRes, cz = i / (kCoarseRes * kCoarseRes);
    bool near_cell = cx == 0 || cy == 0 || cz == 0 || cx == kCoarseRes - 1 || cy == kCoarseRes - 1 || cz == kCoarseRes - 1;
    for (int dz = -kCoarseReach; dz <= kCoarseReach && !near_cell; ++dz)
        for (int dy = -kCoarseReach; dy <= kCoarseReach && !near_cell; ++dy)
            for (int dx = -kCoarseReach; dx <= kCoarseReach; ++dx) {
                const int x = cx + dx, y = cy + dy, z = cz + dz;
                if (x < 0 || y < 0 || z < 0 || x >= kCoarseRes || y >= kCoarseRes || z >= kCoarseRes) continue;
                if (occ[(z * kCoarseRes + y) * kCoarseRes + x]) { near_cell = true; break; }
            }
    if (near_cell) atomicOr(near_bits + (cz * kCoarseRes + cy) * kCoarseRowWords + (cx >> 5), 1u << (cx & 31));
}
void launch_coarse_build(const uint8_t* d_bitfield, uint8_t* d_occ_scratch, uint32_t* d_near_bits, cudaStream_t s) {
    constexpr int n = kCoarseRes * kCoarseRes * kCoarseRes;
    cudaMemsetAsync(d_near_bits, 0, sizeof(uint32_t) * kCoarseRes * kCoarseRes * kCoarseRowWords, s);
    coarse_occupied_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_bitfield, d_occ_scratch);
    coarse_near_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_occ_scratch, d_near_bits);
}

// Brick layout of one level (DeviceModel::brick): cell (gx, gy, gz) -> the eight corner entries the plain layout returns for it,
// in corner order.  Load time only.
__global__ void brick_build_kernel(DeviceModel M, int level, uint32_t res, uint4* __restrict__ out) {
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= res * res * res) return;
    const uint32_t gx = cell % res, gy = (cell / res) % res, gz = cell / (res * res);
    uint32_t index[8];
    level_indices(M, level, gx, gy, gz, index);
    const uint32_t* __restrict__ grid = reinterpret_cast<const uint32_t*>(M.level_ptr[level]);
    uint32_t v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __ldg(grid + index[c]);
    out[2 * (size_t)cell] = make_uint4(v[0], v[1], v[2], v[3]);
    out[2 * (size_t)cell + 1] = make_uint4(v[4], v[5], v[6], v[7]);
}
void launch_brick_build(const DeviceModel& M, int level, uint32_t res, void* d_out, cudaStream_t s) {
    const uint32_t cells = res * res * res;
    brick_build_kernel<<<(cells + 255) / 256, 256, 0, s>>>(M, level, res, reinterpret_cast<uint4*>(d_out));
}

void launch_occupancy_build(const uint16_t* d_grid, int n_cascades_present, uint8_t* d_bitfield, float* d_scratch, cudaStream_t s) {
    double* sum = reinterpret_cast<double*>(d_scratch);
    cudaMemsetAsync(sum, 0, sizeof(double), s);
    const __half* g = reinterpret_cast<const __half*>(d_grid);
    occupancy_mean_kernel<<<296, 256, 0, s>>>(g, sum);
    const uint32_t total = GRID_CELLS / 8 * NERF_CASCADES;
    occupancy_bits_kernel<<<(total + 255) / 256, 256, 0, s>>>(g, GRID_CELLS / 8 * (uint32_t)n_cascades_present, sum, d_bitfield);
    for (uint32_t level = 1; level < NERF_CASCADES; ++level)
        occupancy_pool_kernel<<<(GRID_CELLS / 64 + 255) / 256, 256, 0, s>>>(d_bitfield + (size_t)GRID_CELLS / 8 * (level - 1), d_bitfield + (size_t)GRID_CELLS / 8 * level);
}

// =================================================================================================================
// mesh stage
// =================================================================================================================
constexpr unsigned long long kZMiss = ~0ull;

// `slices` warps per triangle (2 for ordinary meshes, 16 when the scene has lens panes - a few huge triangles): conservative screen bounding box, then the exact ray/triangle test per covered
// sub-pixel; the box's sub-pixels are dealt to the triangle's warps in interleaved groups of 32, so that a triangle covering
// tens of thousands of sub-pixels (a lens pane at 4K) does not hang on one warp.
// Replaces optixLaunch(2W x 2H) + RT-core traversal (S/nerf_mesh_renderer.cu:1454-1487, S/optix/optix_scene.cu:120-174).
__global__ void __launch_bounds__(256) mesh_raster_kernel(MeshDevice mesh, FrameParams P, int W2, int H2, unsigned long long* __restrict__ zbuf, uint32_t slices) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // (the set-up kernel waits for this grid's end by itself)
    const uint32_t warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t tri = warp_id / slices, slice = warp_id % slices;
    const uint32_t lane = threadIdx.x & 31;
    if (tri >= mesh.n_tris) return;
    const V3 eye = v3(P.cam[9], P.cam[10], P.cam[11]);
    const V3 v0 = ld3(mesh.wpos, __ldg(mesh.idx + tri * 3)), v1 = ld3(mesh.wpos, __ldg(mesh.idx + tri * 3 + 1)), v2 = ld3(mesh.wpos, __ldg(mesh.idx + tri * 3 + 2));
    // the visibility buffer only exists inside the mesh's screen bounding box
    const int bx0 = P.zb_x0, by0 = P.zb_y0, bx1 = P.zb_x0 + P.zb_w - 1, by1 = P.zb_y0 + P.zb_h - 1;
    int x0 = bx0, y0 = by0, x1 = bx1, y1 = by1;
    {
        const V3 vs[3] = {v0, v1, v2};
        float minx = 1e30f, miny = 1e30f, maxx = -1e30f, maxy = -1e30f;
        bool behind = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const V3 p = vsub(vs[k], eye);
            const float a = P.cam_inv[0] * p.x + P.cam_inv[1] * p.y + P.cam_inv[2] * p.z;
            const float b = P.cam_inv[3] * p.x + P.cam_inv[4] * p.y + P.cam_inv[5] * p.z;
            const float c = P.cam_inv[6] * p.x + P.cam_inv[7] * p.y + P.cam_inv[8] * p.z;
            if (!(c > 1e-4f)) { behind = true; break; }
            const float px = (a / c + 1.0f) * 0.5f * (float)W2 - 0.5f, py = (b / c + 1.0f) * 0.5f * (float)H2 - 0.5f;
            minx = fminf(minx, px); maxx = fmaxf(maxx, px); miny = fminf(miny, py); maxy = fmaxf(maxy, py);
        }
        if (!behind) {
            if (maxx < -2.f || maxy < -2.f || minx > (float)W2 + 1.f || miny > (float)H2 + 1.f) return;
            x0 = max(bx0, (int)floorf(minx) - 1); y0 = max(by0, (int)floorf(miny) - 1);
            x1 = min(bx1, (int)ceilf(maxx) + 1); y1 = min(by1, (int)ceilf(maxy) + 1);
            if (x1 < x0 || y1 < y0) return;
        }
    }
    // lens surfaces have their own visibility window right behind the opaque one (nearest opaque hit and nearest lens hit
    // are both needed per sub-pixel)
    unsigned long long* __restrict__ zwin = zbuf + ((mesh.tri_lens != nullptr && __ldg(mesh.tri_lens + tri)) ? (size_t)P.zb_w * P.zb_h : (size_t)0);
    // shard: only sub-pixel rows belonging to owned image rows are needed, but testing ownership per row is cheap enough
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    const int ms = P.mesh_scale;
    // the triangle set-up above only reads the mesh; the visibility window is the clear kernel's to reset first (launch_chained)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = (int)(slice * 32u + lane); i < bw * bh; i += 32 * (int)slices) {
        const int x = x0 + i % bw, y = y0 + i / bw;
        if (P.shard_world > 1 && ((y / ms) / P.shard_band) % P.shard_world != P.shard_rank) continue;
        const V3 dir = mesh_ray_dir(P, x, y, W2, H2);
        float t, u, v;
        if (ray_tri(eye, dir, v0, v1, v2, t, u, v) && t < 1e16f) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | tri;
            NMR_DEVICE_CHECK(y >= by0 && y - by0 < P.zb_h && x >= bx0 && x - bx0 < P.zb_w);
            atomicMin(zwin + (size_t)(y - by0) * P.zb_w + (x - bx0), key);
        }
    }
}

// Programmatic dependent launch of a frame's small leading kernels (clear -> mesh raster -> set-up): each is launched while its
// predecessor still runs, sits in griddepcontrol.wait until that one has finished and flushed, and so starts without the launch
// gap.  (griddepcontrol.wait is a no-op in a grid that was launched the ordinary way.)  NMR_NO_PDL_CHAIN=1: ordinary launches.
static bool pdl_chain() { static const bool off = std::getenv("NMR_NO_PDL_CHAIN") != nullptr; return !off; }
template <typename K, typename... Args>
static void launch_chained(K kernel, dim3 grid, dim3 block, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_chain() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

// One launch instead of three memset nodes at the head of a frame: device counters, the schedule histogram and the mesh
// visibility window (every stream node costs 2-4 us of device time on a 250 us frame).
__global__ void frame_clear_kernel(uint32_t* __restrict__ counters, uint32_t* __restrict__ hist, ulonglong2* __restrict__ zbuf2, size_t zbuf_pairs, float4* __restrict__ queue) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // (the mesh raster behind this kernel waits for its end by itself)
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    // ready words of the records the previous frame queued (its march kernel left their number; nobody writes that word here)
    const uint32_t prev = counters[kCntPrevCount];
    if (queue) for (size_t k = i; k < (size_t)prev; k += stride) reinterpret_cast<uint32_t*>(queue + k * kRayRecordFloat4s + 1)[2] = kEmptyRecord;
    if (i < (size_t)kFrameCounters) counters[i] = 0u;
    if (i == 0) { counters[kCntLensRays] = 0u; counters[kCntMarchStart] = 0xFFFFFFFFu; counters[kCntMarchStart + 1] = 0xFFFFFFFFu; counters[kCntMarchEnd] = 0u; counters[kCntMarchEnd + 1] = 0u; }
    if (hist) for (size_t k = i; k < kSchedBins; k += stride) hist[k] = 0u;
    if (zbuf2) for (size_t k = i; k < zbuf_pairs; k += stride) zbuf2[k] = make_ulonglong2(~0ull, ~0ull);
}
size_t zbuf_window_words(const MeshDevice& mesh, const FrameParams& P) {
    if (P.mesh_scale <= 0 || mesh.n_tris == 0 || P.zb_w <= 0 || P.zb_h <= 0) return 0;
    return (size_t)P.zb_w * P.zb_h * (mesh.tri_lens ? 2 : 1);
}
void launch_frame_clear(uint32_t* d_counters, uint32_t* d_hist, unsigned long long* d_zbuf, size_t zbuf_words, float4* d_queue, cudaStream_t s) {
    const size_t pairs = (zbuf_words + 1) / 2;        // the buffer is allocated for the whole 2W x 2H frame (even), the window never fills it to the last word
    const size_t floor_work = (size_t)148 * 256;      // (the previous frame's record count is only known on the device: keep a grid that copes with a busy frame)
    const size_t work = std::max(std::max(pairs, (size_t)kSchedBins), floor_work);
    const unsigned blocks = (unsigned)((work + 255) / 256 < 592 ? (work + 255) / 256 : 592);
    frame_clear_kernel<<<blocks ? blocks : 1, 256, 0, s>>>(d_counters, d_hist, zbuf_words ? reinterpret_cast<ulonglong2*>(d_zbuf) : nullptr, pairs, d_queue);
}

void launch_mesh_raster(const MeshDevice& mesh, const FrameParams& P, int rows_owned, unsigned long long* d_zbuf, cudaStream_t s, bool clear) {
    (void)rows_owned;
    const int W2 = P.width * P.mesh_scale, H2 = P.height * P.mesh_scale;
    if (mesh.n_tris == 0 || P.zb_w <= 0 || P.zb_h <= 0) return;
    if (clear) cudaMemsetAsync(d_zbuf, 0xFF, (size_t)P.zb_w * P.zb_h * sizeof(unsigned long long) * (mesh.tri_lens ? 2 : 1), s);
    const uint32_t slices = mesh.tri_lens ? 16u : 2u;
    const uint32_t threads = mesh.n_tris * 32u * slices;
    launch_chained(mesh_raster_kernel, dim3((threads + 255) / 256), dim3(256), s, mesh, P, W2, H2, d_zbuf, slices);
}

// closest hit of one sub-pixel -> shaded RGBA (alpha 1) and hitT; false on miss
__device__ __forceinline__ bool mesh_tap(const MeshDevice& mesh, const FrameParams& P, const unsigned long long* __restrict__ zbuf, int x, int y, int W2, int H2,
                                         float rgba[4], float& hit_t, int32_t* tri_out = nullptr) {
    const int rx = x - P.zb_x0, ry = y - P.zb_y0;
    const unsigned long long key = (rx >= 0 && ry >= 0 && rx < P.zb_w && ry < P.zb_h) ? __ldg(zbuf + (size_t)ry * P.zb_w + rx) : kZMiss;
    if (key == kZMiss) { if (tri_out) *tri_out = -1; return false; }
    const uint32_t tri = (uint32_t)(key & 0xFFFFFFFFull);
    const V3 eye = v3(P.cam[9], P.cam[10], P.cam[11]);
    const V3 dir = mesh_ray_dir(P, x, y, W2, H2);
    const V3 v0 = ld3(mesh.wpos, __ldg(mesh.idx + tri * 3)), v1 = ld3(mesh.wpos, __ldg(mesh.idx + tri * 3 + 1)), v2 = ld3(mesh.wpos, __ldg(mesh.idx + tri * 3 + 2));
    float t = 0.f, u = 0.f, v = 0.f;
    ray_tri(eye, dir, v0, v1, v2, t, u, v);      // same arithmetic as the raster pass: reproduces the stored hit
    shade_hit(mesh, P, tri, u, v, t, dir, rgba);
    hit_t = t;
    if (tri_out) *tri_out = (int32_t)tri;
    return true;
}

// copyRaytracingBuffersToNerfRays (S/nerf_mesh_renderer.cu:64-100) fused with the shading of the taps
__device__ __forceinline__ void mesh_resolve(const MeshDevice& mesh, const FrameParams& P, const unsigned long long* __restrict__ zbuf, int px, int py,
                                             float surf[4], float& t_surface) {
    const int ms = P.mesh_scale, W2 = P.width * ms, H2 = P.height * ms;
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f, depth = 0.f;
    // the box is aligned to whole pixels: a pixel outside it has no tap inside it (0 + 0 + ... / q == 0 exactly)
    if (px * ms < P.zb_x0 || py * ms < P.zb_y0 || px * ms >= P.zb_x0 + P.zb_w || py * ms >= P.zb_y0 + P.zb_h) {
        surf[0] = surf[1] = surf[2] = surf[3] = 0.f; t_surface = 0.f;
        return;
    }
    for (int i = 0; i < ms; ++i) {
        for (int j = 0; j < ms; ++j) {
            float rgba[4]; float ht;
            if (mesh_tap(mesh, P, zbuf, px * ms + i, py * ms + j, W2, H2, rgba, ht)) {
                c0 += rgba[0]; c1 += rgba[1]; c2 += rgba[2]; c3 += rgba[3];
                depth = fmaxf(depth, ht);
            } else {
                c0 += 0.f; c1 += 0.f; c2 += 0.f; c3 += 0.f;
            }
        }
    }
    const float q = (float)(ms * ms);
    surf[0] = c0 / q; surf[1] = c1 / q; surf[2] = c2 / q; surf[3] = c3 / q;
    t_surface = depth;
}

// Lens hand-off of one pixel (new functionality, no reference equivalent - DESIGN.md "Secondary rays").  A tap counts when
// its nearest lens hit lies in front of its nearest opaque hit; w = counted taps / taps, t = the smallest such hit distance,
// n = unit shading normal there, turned towards the eye.  Hit distances are read back from the visibility keys (exact bits).
struct LensHit { float w, t; V3 n; };
__device__ __forceinline__ void lens_resolve(const MeshDevice& mesh, const FrameParams& P, const unsigned long long* __restrict__ zbuf, int px, int py, LensHit& L) {
    L.w = 0.f; L.t = 0.f; L.n = v3(0.f, 0.f, 1.f);
    const int ms = P.mesh_scale, W2 = P.width * ms, H2 = P.height * ms;
    if (px * ms < P.zb_x0 || py * ms < P.zb_y0 || px * ms >= P.zb_x0 + P.zb_w || py * ms >= P.zb_y0 + P.zb_h) return;
    const unsigned long long* __restrict__ zl = zbuf + (size_t)P.zb_w * P.zb_h;
    int n_l = 0, bx = 0, by = 0; uint32_t btri = 0; float best = 3.402823466e+38f;
    for (int i = 0; i < ms; ++i) {
        for (int j = 0; j < ms; ++j) {
            const int x = px * ms + i, y = py * ms + j;
            const size_t o = (size_t)(y - P.zb_y0) * P.zb_w + (x - P.zb_x0);
            const unsigned long long kl = __ldg(zl + o);
            if (kl == kZMiss) continue;
            const unsigned long long ko = __ldg(zbuf + o);
            const float tl = __uint_as_float((uint32_t)(kl >> 32));
            if (ko != kZMiss && !(tl < __uint_as_float((uint32_t)(ko >> 32)))) continue;
            ++n_l;
            if (tl < best) { best = tl; bx = x; by = y; btri = (uint32_t)(kl & 0xFFFFFFFFull); }
        }
    }
    if (n_l == 0) return;
    const V3 eye = v3(P.cam[9], P.cam[10], P.cam[11]);
    const V3 dir = mesh_ray_dir(P, bx, by, W2, H2);
    const uint32_t i0 = __ldg(mesh.idx + btri * 3), i1 = __ldg(mesh.idx + btri * 3 + 1), i2 = __ldg(mesh.idx + btri * 3 + 2);
    float t = 0.f, u = 0.f, v = 0.f;
    ray_tri(eye, dir, ld3(mesh.wpos, i0), ld3(mesh.wpos, i1), ld3(mesh.wpos, i2), t, u, v);
    const float bw = 1.0f - u - v;
    V3 n = gnormalize(vadd(vadd(vmul(ld3(mesh.wnrm, i1), u), vmul(ld3(mesh.wnrm, i2), v)), vmul(ld3(mesh.wnrm, i0), bw)));
    L.w = (float)n_l / (float)(ms * ms); L.t = best; L.n = n;
}

// Fresnel split at the lens for a unit ray direction d: returns the Schlick reflectance and the mirrored direction
__device__ __forceinline__ float lens_fresnel(const FrameParams& P, V3 d, V3 n, V3& refl) {
    float c = gdot(d, n);
    if (c > 0.f) { n = vmul(n, -1.f); c = -c; }        // normal towards the incoming ray
    const float cosi = fminf(-c, 1.0f);
    refl = gnormalize(vsub(d, vmul(n, 2.0f * c)));
    const float m = 1.0f - cosi, m2 = m * m;
    return P.lens_f0 + (1.0f - P.lens_f0) * (m2 * m2 * m);
}

// Lens model 1 ("plate", oracle: lens_plate_shift): a pane of thickness d with parallel faces.  The ray refracts into the glass at
// the hit point (Snell, relative index 1 / ior), crosses it and leaves parallel to its old direction: the transmitted segment is the
// primary ray shifted sideways by `delta`, resumed at t_behind.  Model 0: no shift.
__device__ __forceinline__ void lens_plate_shift(const FrameParams& P, V3 dir, V3 n, float t_lens, V3& delta, float& t_behind) {
    delta = v3(0.f, 0.f, 0.f); t_behind = t_lens;
    if (P.lens_model != 1 || !(P.lens_thickness > 0.f)) return;
    float c = gdot(dir, n);
    if (c > 0.f) { n = vmul(n, -1.f); c = -c; }
    const float cosi = fmaxf(fminf(-c, 1.0f), 0.05f);
    const float eta = 1.0f / P.lens_ior;
    const float cost = sqrtf(fmaxf(0.f, 1.0f - (eta * eta) * (1.0f - cosi * cosi)));
    const V3 td = vadd(vmul(dir, eta), vmul(n, eta * cosi - cost));
    const float d = P.lens_thickness;
    delta = vsub(vmul(td, d / cost), vmul(dir, d / cosi));
    t_behind = t_lens + d / cosi;
}

// =================================================================================================================
// pixel finish: shade_kernel_nerf + accumulate_kernel + tonemap_kernel
// (S/ngp/testbed.cu:907-931; S/ngp/render_buffer.cu:232-267, 327-346, 537-566) fused into the ray's last step
// =================================================================================================================
__device__ __forceinline__ void finish_pixel(const FrameParams& P, const FrameOut& out, uint32_t idx, float r, float g, float b, float a, float depth, uint32_t n_samples) {
    NMR_DEVICE_CHECK(idx < (uint32_t)(P.width * P.height));
    float4 fb = make_float4(0.f, 0.f, 0.f, 0.f);
    float d = 1e10f;
    if (a > 0.001f) {   // compact_kernel_nerf's hit criterion
        fb = make_float4(srgb_to_linear(r), srgb_to_linear(g), srgb_to_linear(b), a);
        if (a > 0.2f) d = depth;
    }
    if (out.frame) out.frame[idx] = fb;
    if (out.depth) out.depth[idx] = d;
    if (out.n_samples) out.n_samples[idx] = n_samples;
    if (a <= 0.001f && P.spp_index == 0) {
        // nothing hit on the first sample of an accumulation: the pixel is the tonemapped background, a per-frame constant
        out.accum[idx] = fb;
        store_pixel(out.image, P.out_format, idx, P.background_out[0], P.background_out[1], P.background_out[2], P.background_out[3]);
        return;
    }
    float4 acc = fb;
    if (P.spp_index != 0) {
        const float sc = (float)P.spp_index;
        const float4 prev = out.accum[idx];
        acc = make_float4((prev.x * sc + fb.x) / (sc + 1), (prev.y * sc + fb.y) / (sc + 1), (prev.z * sc + fb.z) / (sc + 1), (prev.w * sc + fb.w) / (sc + 1));
    }
    out.accum[idx] = acc;
    const float w = (1 - acc.w) * P.background[3];
    float cr = acc.x + P.background_linear[0] * w;
    float cg = acc.y + P.background_linear[1] * w;
    float cb = acc.z + P.background_linear[2] * w;
    float ca = acc.w + w;
    tonemap_curve_apply(cr, cg, cb, P.tonemap_curve);
    if (P.to_srgb) {
        cr = fminf(fmaxf(linear_to_srgb(cr), 0.f), 1.f); cg = fminf(fmaxf(linear_to_srgb(cg), 0.f), 1.f);
        cb = fminf(fmaxf(linear_to_srgb(cb), 0.f), 1.f); ca = fminf(fmaxf(ca, 0.f), 1.f);
    }
    store_pixel(out.image, P.out_format, idx, cr, cg, cb, ca);
}

// =================================================================================================================
// init_rays_kernel: init_rays_with_payload_kernel_nerf + mesh hand-off + advance_pos_nerf
// (S/ngp/testbed.cu:355-537; S/nerf_mesh_renderer.cu:64-100)
// =================================================================================================================
__device__ __forceinline__ void init_one_ray(const FrameParams& P, const DeviceModel& M, const MeshDevice& mesh, const unsigned long long* __restrict__ zbuf,
                                             float4* __restrict__ queue, uint32_t* __restrict__ counters, const FrameOut& out, int x, int y, uint32_t* __restrict__ surf_list) {
    const uint32_t idx = (uint32_t)x + (uint32_t)P.width * (uint32_t)y;

    float surf[4] = {0.f, 0.f, 0.f, 0.f};
    float t_surface = 0.f;
    if (P.mesh_scale > 0) mesh_resolve(mesh, P, zbuf, x, y, surf, t_surface);
    LensHit L; L.w = 0.f; L.t = 0.f; L.n = v3(0.f, 0.f, 1.f);
    if (P.lens_on) {
        lens_resolve(mesh, P, zbuf, x, y, L);
        if (L.w > 0.f && t_surface != 0.0f && t_surface < L.t) L.w = 0.f;   // an opaque part of the mesh is in front: no lens event
    }
    // the first mesh surface the primary ray can meet: it clips / revives the first-hit walk exactly like t_surface does
    const float t_first = L.w > 0.f ? L.t : t_surface;

    // Most pixels see neither the mesh nor any occupied cell: their ray misses the box around the occupied cells, so
    // advance_pos_nerf could only walk it out of the render box (dead ray, background pixel).  Decide that with the
    // un-normalised direction and approximate reciprocals - the box carries a whole grid cell of margin - before paying for
    // the exact ray set-up.
    if (t_first == 0.0f) {
        const float* c = P.cam;
        const float ux = 2.0f * (((float)x + 0.5f) / (float)P.width) - 1.0f;
        const float uy = 2.0f * (((float)y + 0.5f) / (float)P.height) - 1.0f;
        const V3 dm = model_rotate(P, v3(c[0] * ux + (c[3] * uy + c[6]), c[1] * ux + (c[4] * uy + c[7]), c[2] * ux + (c[5] * uy + c[8])));
        const float d[3] = {dm.x, dm.y, dm.z};
        const float o[3] = {P.ray_origin[0], P.ray_origin[1], P.ray_origin[2]};
        float tmin = 0.f, tmax = 3.402823466e+38f;        // only the part of the line in front of the eye counts
        bool miss = false;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (fabsf(d[k]) > 1e-12f) {
                const float inv = __frcp_rn(d[k]);
                const float a = (P.occ_min[k] - o[k]) * inv, b = (P.occ_max[k] - o[k]) * inv;
                tmin = fmaxf(tmin, fminf(a, b)); tmax = fminf(tmax, fmaxf(a, b));
            } else if (o[k] < P.occ_min[k] || o[k] > P.occ_max[k]) {
                miss = true;
            }
        }
        if (miss || tmin > tmax) {
            finish_pixel(P, out, idx, 0.f, 0.f, 0.f, 0.f, 0.f, 0u);
            return;
        }
    }

    RayInit r = init_ray(P, (uint32_t)x, (uint32_t)y);
    float t = r.t, t_start;
    const bool alive = advance_pos(P, M.bitfield, M.coarse, r.origin, r.dir, idx, t_first, r.t_occ_in, r.t_limit, r.alive, t, t_start);
    if (!alive) {
        finish_pixel(P, out, idx, 0.f, 0.f, 0.f, 0.f, 0.f, 0u);
        return;
    }
    const uint32_t slot = atomicAdd(&counters[0], 1u);
    NMR_DEVICE_CHECK(slot < (uint32_t)(P.width * P.height));
    queue[(size_t)slot * kRayRecordFloat4s + 0] = make_float4(r.dir.x, r.dir.y, r.dir.z, t);
    const bool carries_surface = L.w == 0.f && t_surface != 0.0f && surf[3] > 0.f;
    if (carries_surface && surf_list) surf_list[atomicAdd(&counters[7], 1u)] = slot;
    float* q1 = reinterpret_cast<float*>(queue + (size_t)slot * kRayRecordFloat4s + 1);
    q1[0] = t_start; q1[1] = t_surface; q1[3] = r.t_limit;
    queue[(size_t)slot * kRayRecordFloat4s + 2] = make_float4(surf[0], surf[1], surf[2], surf[3]);
    if (L.w > 0.f) {
        atomicAdd(&counters[kCntLensRays], 1u);
        const V3 ln = model_rotate(P, L.n);     // the mirror direction is taken in NeRF space
        out.lens[(size_t)idx * 2] = make_float4(ln.x, ln.y, ln.z, L.t);
        out.lens[(size_t)idx * 2 + 1] = make_float4(L.w, 0.f, 0.f, 0.f);
    }
    // the ready word last, with release semantics: whoever reads it (acquire) also sees the rest of the record and the lens entry
    const uint32_t ready = idx | (L.w > 0.f ? kLensRayFlag : 0u) | (carries_surface ? kSurfRayFlag : 0u);
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(q1 + 2), "r"(ready) : "memory");
}

// A pixel that can only be background (outside both screen rectangles): the finished pixel of an empty ray
__device__ __forceinline__ void background_pixel(const FrameParams& P, const FrameOut& out, uint32_t idx) {
    if (P.bg_filled_elsewhere) {
        // the constant background of these pixels is written into the shared image by its owner (fill_background_kernel): only
        // pixels inside the rectangles cross NVLink.  The local accumulator still gets its value (zero stays zero under accumulation).
        out.accum[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (out.frame) out.frame[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (out.depth) out.depth[idx] = 1e10f;
        if (out.n_samples) out.n_samples[idx] = 0u;
    } else {
        finish_pixel(P, out, idx, 0.f, 0.f, 0.f, 0.f, 0.f, 0u);
    }
}

// The set-up kernel only covers the TILE BOX: the 16 x 8 pixel tiles [tile_x0, tile_x0 + gridDim.x) x [tile_y0, ...) of local
// (owned) rows that the union of the two screen rectangles touches.  Everything else is background, written by background_kernel.

// 16 x 8 pixel CTA, each warp an 8 x 4 tile so queue neighbours are screen neighbours.  A pixel outside both the screen
// rectangle of the box around the occupied cells (FrameParams::occ_px, projected on the host) and the mesh's screen rectangle
// can only be background: integer compares, one store pair, no ray arithmetic.
__device__ __forceinline__ void init_rays_body(const FrameParams& P, const DeviceModel& M, const MeshDevice& mesh, const unsigned long long* __restrict__ zbuf, int rows_owned,
                                               float4* __restrict__ queue, uint32_t* __restrict__ counters, const FrameOut& out, uint32_t* __restrict__ surf_list, TileBox box, uint2 block_rot) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTAs are handed out in index order; the ones over the occupied cells carry the kernel's long dependent chains (first-hit
    // walks), so the index space is rotated to start there: those chains begin at time zero.
    const int bx = (int)((blockIdx.x + block_rot.x) % gridDim.x) + box.x0, by = (int)((blockIdx.y + block_rot.y) % gridDim.y) + box.y0;
    const int x = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int ly = by * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (x >= P.width || ly >= rows_owned) return;
    const int y = shard_row(P, ly), ms = P.mesh_scale;
    const bool in_occ = x >= P.occ_px[0] && x < P.occ_px[2] && y >= P.occ_px[1] && y < P.occ_px[3];
    const bool in_mesh = ms > 0 && P.zb_w > 0 && x * ms >= P.zb_x0 && x * ms < P.zb_x0 + P.zb_w && y * ms >= P.zb_y0 && y * ms < P.zb_y0 + P.zb_h;
    if (!in_occ && !in_mesh) { background_pixel(P, out, (uint32_t)x + (uint32_t)P.width * (uint32_t)y); return; }
    init_one_ray(P, M, mesh, zbuf, queue, counters, out, x, y, surf_list);
}
__global__ void __launch_bounds__(128) init_rays_kernel(FrameParams P, DeviceModel M, MeshDevice mesh, const unsigned long long* __restrict__ zbuf, int rows_owned,
                                                        float4* __restrict__ queue, uint32_t* __restrict__ counters, FrameOut out, uint32_t* __restrict__ surf_list, TileBox box, uint2 block_rot) {
    // Overlapped frames: the march kernel behind this one is launched with programmatic stream serialisation - it may start as
    // soon as every CTA of this grid has got here (i.e. is resident or done; the tile box is at most a few waves), and then
    // consumes the queue while the first-hit walks are still running.  Harmless when the next kernel is an ordinary launch.
    // This grid itself may have been launched early (launch_chained): first wait for the mesh raster (and, through it, the clear
    // kernel) - only then may the march kernel be let loose on the counters and ready words they reset.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    init_rays_body(P, M, mesh, zbuf, rows_owned, queue, counters, out, surf_list, box, block_rot);
    // this CTA's records are complete: count it (the march kernel compares the count with the grid size to learn that the queue is final)
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(&counters[kCntInitDone], 1u); }
}

// Every owned pixel OUTSIDE the tile box: constant background (one store pair per pixel, bandwidth-bound), by a grid sized to the
// machine instead of thousands of tiny CTAs in front of the march kernel.  Its first threads also ask for the hash table line
// by line (prefetch.global.L2): the march kernel is about to gather from it at random, and when it has left the L2 since the
// last frame a 4-byte gather would cost a DRAM round trip per level.
__global__ void __launch_bounds__(256) background_kernel(FrameParams P, FrameOut out, int rows_owned, TileBox box, const char* __restrict__ prefetch_base, uint32_t prefetch_lines) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    for (uint32_t g = tid; g < prefetch_lines; g += n_threads) asm volatile("prefetch.global.L2 [%0];" ::"l"(prefetch_base + (size_t)g * 128u));
    const int bx0 = box.x0 * 16, bx1 = (box.x0 + box.nx) * 16, by0 = box.y0 * 8, by1 = (box.y0 + box.ny) * 8;
    const uint32_t W = (uint32_t)P.width, n = W * (uint32_t)rows_owned;
    for (uint32_t i = tid; i < n; i += n_threads) {
        const int ly = (int)(i / W), x = (int)(i - (uint32_t)ly * W);
        if (box.nx > 0 && x >= bx0 && x < bx1 && ly >= by0 && ly < by1) continue;
        background_pixel(P, out, (uint32_t)x + W * (uint32_t)shard_row(P, ly));
    }
}

// Several NeRFs in one frame (NerfMeshRenderer::render_frame, S/nerf_mesh_renderer.cu:582-597): every NeRF renders into its own
// linear frame + depth buffers; the first one's are copied, the others z-merged with combineBuffersKernel's rule
// (S/nerf_mesh_renderer.cu:34-48): the nearer depth takes the pixel.
__global__ void combine_buffers_kernel(const float* __restrict__ in_depth, const float4* __restrict__ in_frame, float* __restrict__ out_depth, float4* __restrict__ out_frame, uint32_t n, int first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d = in_depth[i];
    if (first || d < out_depth[i]) { out_depth[i] = d; out_frame[i] = in_frame[i]; }
}
void launch_combine_buffers(const float* d_in_depth, const float4* d_in_frame, float* d_out_depth, float4* d_out_frame, uint32_t n, bool first, cudaStream_t s) {
    combine_buffers_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_in_depth, d_in_frame, d_out_depth, d_out_frame, n, first ? 1 : 0);
}
// accumulate + tonemap of a merged linear frame (the tail of finish_pixel: S/ngp/render_buffer.cu:232-267, 537-566)
__global__ void present_kernel(FrameParams P, const float4* __restrict__ frame, float4* __restrict__ accum, void* __restrict__ image, uint32_t n) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float4 fb = frame[idx];
    float4 acc = fb;
    if (P.spp_index != 0) {
        const float sc = (float)P.spp_index;
        const float4 prev = accum[idx];
        acc = make_float4((prev.x * sc + fb.x) / (sc + 1), (prev.y * sc + fb.y) / (sc + 1), (prev.z * sc + fb.z) / (sc + 1), (prev.w * sc + fb.w) / (sc + 1));
    }
    accum[idx] = acc;
    const float w = (1 - acc.w) * P.background[3];
    float cr = acc.x + P.background_linear[0] * w, cg = acc.y + P.background_linear[1] * w, cb = acc.z + P.background_linear[2] * w, ca = acc.w + w;
    tonemap_curve_apply(cr, cg, cb, P.tonemap_curve);
    if (P.to_srgb) {
        cr = fminf(fmaxf(linear_to_srgb(cr), 0.f), 1.f); cg = fminf(fmaxf(linear_to_srgb(cg), 0.f), 1.f);
        cb = fminf(fmaxf(linear_to_srgb(cb), 0.f), 1.f); ca = fminf(fmaxf(ca, 0.f), 1.f);
    }
    store_pixel(image, P.out_format, idx, cr, cg, cb, ca);
}
void launch_present(const FrameParams& P, const float4* d_frame, float4* d_accum, void* d_image, uint32_t n, cudaStream_t s) {
    present_kernel<<<(n + 255) / 256, 256, 0, s>>>(P, d_frame, d_accum, d_image, n);
}

// Shared frame target, destination rank: the constant background of every pixel outside both screen rectangles (the same
// rectangles on every rank: same camera, same scene), for ALL rows - so that the other ranks only send the pixels inside them.
__global__ void fill_background_kernel(FrameParams P, void* __restrict__ image) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.width) return;
    const int ms = P.mesh_scale;
    const bool in_occ = x >= P.occ_px[0] && x < P.occ_px[2] && y >= P.occ_px[1] && y < P.occ_px[3];
    const bool in_mesh = ms > 0 && P.zb_w > 0 && x * ms >= P.zb_x0 && x * ms < P.zb_x0 + P.zb_w && y * ms >= P.zb_y0 && y * ms < P.zb_y0 + P.zb_h;
    if (!in_occ && !in_mesh) store_pixel(image, P.out_format, (uint32_t)y * (uint32_t)P.width + (uint32_t)x, P.background_out[0], P.background_out[1], P.background_out[2], P.background_out[3]);
}
void launch_fill_background(const FrameParams& P, void* d_image, cudaStream_t s) {
    dim3 grid((P.width + 255) / 256, P.height);
    fill_background_kernel<<<grid, 256, 0, s>>>(P, d_image);
}
// ---- tile-sharded frames written straight into one rank's image (nmr_gather_*): sequence flags in that rank's memory ----
// flags[r] (r < 32) = last frame rank r has finished writing; flags[kGatherConsumed] = last frame the destination is done with;
// flags[kGatherError] != 0 after a wait that ran out of time.  Stores of a rank's render kernels precede its signal kernel in
// stream order; the fence + system-scope store below publish them to the peer that polls the flag.
__global__ void gather_signal_kernel(volatile uint32_t* flag, uint32_t seq) {
    __threadfence_system();
    *flag = seq;
    __threadfence_system();
}
__global__ void gather_wait_kernel(volatile uint32_t* flags, int first, int count, uint32_t seq, volatile uint32_t* err, unsigned long long timeout_ns) {
    const int i = threadIdx.x;
    if (i >= count) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int32_t)(flags[first + i] - seq) < 0) {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) { *err = 1u; break; }        // a rank never rendered this frame - give up instead of hanging the GPU; the destination's frame() fails
    }
    __threadfence_system();
}
void launch_gather_signal(uint32_t* d_flag, uint32_t seq, cudaStream_t s) { gather_signal_kernel<<<1, 1, 0, s>>>(d_flag, seq); }
void launch_gather_wait(uint32_t* d_flags, int first, int count, uint32_t seq, uint32_t* d_err, cudaStream_t s) {
    // NMR_GATHER_TIMEOUT_MS: how long a rank may take between frames (saving images, a debugger) before the others give up; 10 s
    static const unsigned long long timeout_ns = [] { const char* v = std::getenv("NMR_GATHER_TIMEOUT_MS"); const long long ms = v ? std::atoll(v) : 10000; return (unsigned long long)(ms > 0 ? ms : 10000) * 1000000ull; }();
    if (count > 0) gather_wait_kernel<<<1, 32, 0, s>>>(d_flags, first, count, seq, d_err, timeout_ns);
}

// owned rows of this context with image row < y (rows are dealt to ranks in bands, kernels.cu: shard_row)
static int local_rows_below(const FrameParams& P, int y) {
    if (y <= 0) return 0;
    if (P.shard_world <= 1) return std::max(0, y - P.row0);
    const int cycle = P.shard_world * P.shard_band, full = y / cycle, rem = y % cycle;
    return full * P.shard_band + std::min(std::max(rem - P.shard_rank * P.shard_band, 0), P.shard_band);
}

TileBox compute_tile_box(const FrameParams& P, int rows_owned) {
    TileBox box{0, 0, 0, 0, 0u, 0u};
    if (rows_owned <= 0) return box;
    // union of the two screen rectangles in pixels (image rows), then in 16 x 8 tiles of local rows
    int x0 = P.width, y0 = P.height, x1 = 0, y1 = 0;
    const bool have_occ = P.occ_px[2] > P.occ_px[0] && P.occ_px[3] > P.occ_px[1];
    if (have_occ) { x0 = std::min(x0, P.occ_px[0]); y0 = std::min(y0, P.occ_px[1]); x1 = std::max(x1, P.occ_px[2]); y1 = std::max(y1, P.occ_px[3]); }
    if (P.mesh_scale > 0 && P.zb_w > 0 && P.zb_h > 0) {
        const int ms = P.mesh_scale;
        x0 = std::min(x0, P.zb_x0 / ms); y0 = std::min(y0, P.zb_y0 / ms);
        x1 = std::max(x1, (P.zb_x0 + P.zb_w + ms - 1) / ms); y1 = std::max(y1, (P.zb_y0 + P.zb_h + ms - 1) / ms);
    }
    if (!(x1 > x0 && y1 > y0)) return box;
    const int ly0 = local_rows_below(P, y0), ly1 = std::min(rows_owned, local_rows_below(P, y1));
    x0 = std::max(0, x0); x1 = std::min(P.width, x1);
    if (!(ly1 > ly0 && x1 > x0)) return box;
    box.x0 = x0 / 16; box.y0 = ly0 / 8; box.nx = (x1 + 15) / 16 - box.x0; box.ny = (ly1 + 7) / 8 - box.y0;
    // first CTA column / row over the occupied cells; NMR_NO_BLOCK_ROTATION=1 for A/B runs
    static const bool no_rot = std::getenv("NMR_NO_BLOCK_ROTATION") != nullptr;
    if (!no_rot && have_occ) {
        const int rx = std::max(x0, P.occ_px[0]) / 16 - box.x0, ry = local_rows_below(P, std::max(y0, P.occ_px[1])) / 8 - box.y0;
        if (rx >= 0 && rx < box.nx && ry >= 0 && ry < box.ny) { box.rot_x = (unsigned)rx; box.rot_y = (unsigned)ry; }
    }
    return box;
}

void launch_background(const FrameParams& P, const DeviceModel& M, const FrameOut& out, int rows_owned, const TileBox& box, bool prefetch_table, int num_sms, cudaStream_t s) {
    if (rows_owned <= 0) return;
    // tables that fit the L2 comfortably are prefetched by the frame's first set-up pass; NMR_NO_PREFETCH=1 for A/B runs
    static const bool no_prefetch = std::getenv("NMR_NO_PREFETCH") != nullptr;
    const size_t table_bytes = ((size_t)M.level_offset[N_LEVELS - 1] + M.level_size[N_LEVELS - 1]) * sizeof(__half2);
    const uint32_t prefetch_lines = (prefetch_table && !no_prefetch && table_bytes <= ((size_t)48 << 20)) ? (uint32_t)(table_bytes / 128) : 0u;
    background_kernel<<<num_sms * 4, 256, 0, s>>>(P, out, rows_owned, box, reinterpret_cast<const char*>(M.grid), prefetch_lines);
}

int launch_init_rays(const FrameParams& P, const DeviceModel& M, const MeshDevice& mesh, const unsigned long long* d_zbuf, int rows_owned,
                     float4* d_queue, uint32_t* d_counters, const FrameOut& out, const TileBox& box, cudaStream_t s, uint32_t* d_surf_list) {
    if (rows_owned <= 0 || box.nx <= 0) return 0;
    launch_chained(init_rays_kernel, dim3((unsigned)box.nx, (unsigned)box.ny), dim3(128), s, P, M, mesh, d_zbuf, rows_owned, d_queue, d_counters, out, d_surf_list, box, make_uint2(box.rot_x, box.rot_y));
    return box.nx * box.ny;
}

// =================================================================================================================
// tcgen05 / tensor-memory plumbing (PTX ISA: tcgen05.*, mbarrier.*; descriptor bit layouts as in
// cute/arch/mma_sm100_desc.hpp: SmemDescriptor, InstrDescriptor)
// =================================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// K-major, no swizzle: 8-row x 16-byte core matrices; lbo = byte distance between K-adjacent core matrices,
// sbo = byte distance between M/N-adjacent core matrices (cute: ((8,n),2):((1,SBO),LBO) in uint128 units)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                               // descriptor version for sm_100
    return d;                                      // base_offset 0, layout_type 0 (SWIZZLE_NONE)
}
// kind::f16, A = B = F16, D = F32, both operands K-major, dense
__device__ __forceinline__ constexpr uint32_t umma_idesc(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t v[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// =================================================================================================================
// network evaluation on a tile of 128 samples (one per thread)
// NerfNetwork::inference_mixed_precision_impl (S/ngp/nerf_network.cuh:101-135):
//   enc(32) -> 64 ReLU -> 16 ; [16 | SH16] -> 64 ReLU -> 64 ReLU -> 16 (3 used); density = channel 0 of the first net.
// Layer semantics of T/src/fully_fused_mlp.cu:499-557: y = act(W x), fp16 activations between layers.
// =================================================================================================================
constexpr int kTile = 128;          // samples per MMA tile = threads per warpgroup
#ifndef NMR_GROUPS_TC
#define NMR_GROUPS_TC 2
#endif
constexpr int kGroupsTC = NMR_GROUPS_TC;   // warpgroups per CTA on the tensor path (they share the weights in shared memory); 6 with NMR_MARCH_CTAS=1: one 768-thread CTA per SM
constexpr uint32_t kTmemCols = 64 * kGroupsTC <= 128 ? 128u : (64 * kGroupsTC <= 256 ? 256u : 512u);   // tcgen05.alloc takes powers of two
// weight matrices in params order: [out][in] row-major halves
constexpr int kWD0 = 0, kWD1 = kWD0 + 64 * 32, kWR0 = kWD1 + 16 * 64, kWR1 = kWR0 + 64 * 32, kWR2 = kWR1 + 64 * 64, kWTotal = kWR2 + 16 * 64;   // 10240 halves

struct __align__(128) MarchSmem {   // CUDA-core path, one 128-thread group per CTA
    __half w[kWTotal];            // plain row-major weights
    __half act[kTile * 64];       // row-major, 64 halves per thread
    __half act2[kTile * 64];
    unsigned long long t_begin;   // globaltimer at the CTA's start (deadline of an overlapped frame's waits)
};
struct __align__(128) MarchSmemTC {
    __half w[kWTotal];                    // 20480 B, canonical K-major core-matrix layout, shared by the warpgroups
    __half act[kGroupsTC][kTile * 64];    // 2 x 16384 B, A operands: chunk-major (k/8)*2048 + row*16
    uint64_t mbar[kGroupsTC];
    uint64_t wbar;                        // completion of the bulk copy that brings the weights in
    uint32_t tmem_base;
    unsigned long long t_begin;           // globaltimer at the CTA's start (deadline of an overlapped frame's waits)
};

// ---- warpgroup-scoped barriers (named barriers 1..kGroupsTC, 128 threads each) -----------------------------------
__device__ __forceinline__ void group_sync(uint32_t bar_id) { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); }
__device__ __forceinline__ bool group_any(uint32_t bar_id, bool pred) {
    uint32_t r;
    asm volatile("{\n .reg .pred p, q;\n setp.ne.u32 p, %1, 0;\n bar.red.or.pred q, %2, 128, p;\n selp.u32 %0, 1, 0, q;\n}" : "=r"(r) : "r"((uint32_t)pred), "r"(bar_id) : "memory");
    return r != 0;
}

// ---- CUDA-core variant (bring-up / bisecting aid, selected with NMR_MLP=scalar): fp32 accumulation in k order ----
__device__ __forceinline__ void scalar_layer(const __half* __restrict__ W, int n_out, int n_in, const __half* __restrict__ x_row, __half* __restrict__ y_row, bool relu) {
    for (int j = 0; j < n_out; ++j) {
        float acc = 0.f;
        const __half* wr = W + j * n_in;
        for (int k = 0; k < n_in; ++k) acc = acc + __half2float(wr[k]) * __half2float(x_row[k]);
        if (relu) acc = acc > 0.f ? acc : 0.f;
        y_row[j] = __float2half_rn(acc);
    }
}

// the caller has written the 32 encoded features to the first 64 bytes of its row of S.act
__device__ __forceinline__ void network_scalar(MarchSmem& S, V3 dir01, float raw[4]) {
    __half* a = S.act + threadIdx.x * 64;
    __half* b = S.act2 + threadIdx.x * 64;
    scalar_layer(S.w + kWD0, 64, 32, a, b, true);
    scalar_layer(S.w + kWD1, 16, 64, b, a, false);
    const float density = __half2float(a[0]);
    __half2 sh[8];
    sh4(dir01, sh);
#pragma unroll
    for (int i = 0; i < 8; ++i) reinterpret_cast<__half2*>(a)[8 + i] = sh[i];
    scalar_layer(S.w + kWR0, 64, 32, a, b, true);
    scalar_layer(S.w + kWR1, 64, 64, b, a, true);
    scalar_layer(S.w + kWR2, 16, 64, a, b, false);
    raw[0] = __half2float(b[0]); raw[1] = __half2float(b[1]); raw[2] = __half2float(b[2]); raw[3] = density;
}

// ---- tensor-core variant ----------------------------------------------------------------------------------------
// smem operand layouts (K-major, SWIZZLE_NONE canonical layout, 16-byte "chunks" of 8 halves):
//   A (activations, M = 128 rows):  byte(row, k) = (k/8)*2048 + row*16 + (k%8)*2      LBO = 2048, SBO = 128
//   B (weights, N rows = outputs):  byte(n,   k) = (k/8)*(N*16) + n*16 + (k%8)*2      LBO = N*16, SBO = 128
// Thread r of a warpgroup owns row r of its A tile and lane r of its accumulator in tensor memory, so every hand-off is
// thread-private: tcgen05.ld 32x32b (lane = thread) -> ReLU/convert in registers -> st.shared of its own row
// (conflict-free, 512 B per warp store).
__device__ __forceinline__ void stage_weights_tc(__half* sw, const __half* __restrict__ gw) {
    const int off[5] = {kWD0, kWD1, kWR0, kWR1, kWR2};
    const int Ns[5] = {64, 16, 64, 64, 16};
    const int Ks[5] = {32, 64, 32, 64, 64};
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        const int N = Ns[m], K = Ks[m], chunks = K / 8;
        for (int i = threadIdx.x; i < N * chunks; i += blockDim.x) {
            const int n = i / chunks, c = i % chunks;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(gw + off[m] + n * K + c * 8));
            *reinterpret_cast<uint4*>(reinterpret_cast<char*>(sw + off[m]) + c * (N * 16) + n * 16) = v;
        }
    }
}

__global__ void weights_to_canonical_kernel(const __half* __restrict__ gw, __half* __restrict__ out) { stage_weights_tc(out, gw); }
void launch_weights_canonical(const uint16_t* d_mlp, uint16_t* d_out, cudaStream_t s) {
    weights_to_canonical_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const __half*>(d_mlp), reinterpret_cast<__half*>(d_out));
}

struct TcCtx {
    uint32_t tmem;        // base address of this warpgroup's 64 accumulator columns
    uint32_t a_addr;      // smem address of this warpgroup's A operand
    uint32_t w_addr;      // smem address of the weights
    uint64_t* mbar;
    uint32_t phase;
    uint32_t bar_id;      // named barrier of the warpgroup
    uint32_t row;         // thread index inside the warpgroup = tile row = accumulator lane
    bool swap;
};

// D[128 x N] = A[128 x K] * W^T  issued by one thread; K in {32, 64}
__device__ __forceinline__ void tc_issue_layer(const TcCtx& c, int w_off_halves, uint32_t N, uint32_t K) {
    const uint32_t idesc = umma_idesc(128, N);
    const uint32_t a_lbo = 2048, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
    const uint32_t b_addr = c.w_addr + (uint32_t)w_off_halves * 2u;
    for (uint32_t k = 0; k < K / 16; ++k) {
        const uint64_t ad = c.swap ? umma_desc(c.a_addr + k * 2 * a_lbo, a_sbo, a_lbo) : umma_desc(c.a_addr + k * 2 * a_lbo, a_lbo, a_sbo);
        const uint64_t bd = c.swap ? umma_desc(b_addr + k * 2 * b_lbo, b_sbo, b_lbo) : umma_desc(b_addr + k * 2 * b_lbo, b_lbo, b_sbo);
        umma_f16(c.tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(c.mbar);
}

// two fp32 accumulators -> packed fp16 (a in the low half), optionally through ReLU: one cvt instruction either way
__device__ __forceinline__ uint32_t pack_relu_h2(uint32_t a, uint32_t b, bool relu) {
    uint32_t r;
    if (relu) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(b)), "f"(__uint_as_float(a)));
    return r;
}

// one layer's barrier / issue / wait sequence: the warpgroup's rows are in shared memory -> accumulator is ready
__device__ __forceinline__ void tc_run_layer(TcCtx& c, int w_off_halves, uint32_t N, uint32_t K) {
    fence_proxy_async();          // generic-proxy st.shared -> visible to the tensor core's async proxy
    tc_fence_before();
    group_sync(c.bar_id);
    if (c.row == 0) { tc_fence_after(); tc_issue_layer(c, w_off_halves, N, K); }
    mbar_wait(c.mbar, c.phase); c.phase ^= 1u;
    tc_fence_after();
}

// accumulator columns [0, 64) of this thread's lane -> ReLU -> fp16 -> own row of A (8 chunks)
__device__ __forceinline__ void tc_hidden_to_smem(const TcCtx& c, char* a_row_base) {
    const uint32_t lane_addr = c.tmem + ((c.row & ~31u) << 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t v[16];
        tmem_ld16(lane_addr + q * 16, v);
        tmem_ld_wait();
        uint4 lo, hi;
        lo.x = pack_relu_h2(v[0], v[1], true); lo.y = pack_relu_h2(v[2], v[3], true); lo.z = pack_relu_h2(v[4], v[5], true); lo.w = pack_relu_h2(v[6], v[7], true);
        hi.x = pack_relu_h2(v[8], v[9], true); hi.y = pack_relu_h2(v[10], v[11], true); hi.z = pack_relu_h2(v[12], v[13], true); hi.w = pack_relu_h2(v[14], v[15], true);
        *reinterpret_cast<uint4*>(a_row_base + (2 * q) * 2048) = lo;
        *reinterpret_cast<uint4*>(a_row_base + (2 * q + 1) * 2048) = hi;
    }
}

// warpgroup-wide: all 128 threads must call (idle threads leave stale rows - rows are independent in the GEMMs).
// The caller has written the encoded features to chunks 0..3 of its row of A (encode_chunks with stride 2048).
__device__ __forceinline__ void network_tc(char* a_row, TcCtx& c, V3 dir01, float raw[4]) {
    // ---- density net: enc(32) -> 64 ReLU -> 16
    tc_run_layer(c, kWD0, 64, 32);
    tc_hidden_to_smem(c, a_row);
    tc_run_layer(c, kWD1, 16, 64);
    const uint32_t lane_addr = c.tmem + ((c.row & ~31u) << 16);
    {
        uint32_t v[16];
        tmem_ld16(lane_addr, v);
        tmem_ld_wait();
        uint4 lo, hi;
        lo.x = pack_relu_h2(v[0], v[1], false); lo.y = pack_relu_h2(v[2], v[3], false); lo.z = pack_relu_h2(v[4], v[5], false); lo.w = pack_relu_h2(v[6], v[7], false);
        hi.x = pack_relu_h2(v[8], v[9], false); hi.y = pack_relu_h2(v[10], v[11], false); hi.z = pack_relu_h2(v[12], v[13], false); hi.w = pack_relu_h2(v[14], v[15], false);
        raw[3] = __low2float(*reinterpret_cast<const __half2*>(&lo.x));   // density = fp16(channel 0), extract_density
        *reinterpret_cast<uint4*>(a_row + 0 * 2048) = lo;                 // rgb net input = [density net out (16) | SH (16)]
        *reinterpret_cast<uint4*>(a_row + 1 * 2048) = hi;
        __half2 sh[8];
        sh4(dir01, sh);
        uint4 s0, s1;
        s0.x = *reinterpret_cast<const uint32_t*>(&sh[0]); s0.y = *reinterpret_cast<const uint32_t*>(&sh[1]); s0.z = *reinterpret_cast<const uint32_t*>(&sh[2]); s0.w = *reinterpret_cast<const uint32_t*>(&sh[3]);
        s1.x = *reinterpret_cast<const uint32_t*>(&sh[4]); s1.y = *reinterpret_cast<const uint32_t*>(&sh[5]); s1.z = *reinterpret_cast<const uint32_t*>(&sh[6]); s1.w = *reinterpret_cast<const uint32_t*>(&sh[7]);
        *reinterpret_cast<uint4*>(a_row + 2 * 2048) = s0;
        *reinterpret_cast<uint4*>(a_row + 3 * 2048) = s1;
    }
    // ---- rgb net: 32 -> 64 ReLU -> 64 ReLU -> 16 (three channels used)
    tc_run_layer(c, kWR0, 64, 32);
    tc_hidden_to_smem(c, a_row);
    tc_run_layer(c, kWR1, 64, 64);
    tc_hidden_to_smem(c, a_row);
    tc_run_layer(c, kWR2, 16, 64);
    {
        uint32_t v[4];
        tmem_ld4(lane_addr, v);
        tmem_ld_wait();
        raw[0] = __half2float(__float2half_rn(__uint_as_float(v[0])));
        raw[1] = __half2float(__float2half_rn(__uint_as_float(v[1])));
        raw[2] = __half2float(__float2half_rn(__uint_as_float(v[2])));
    }
    tc_fence_before();   // the next tile's first MMA overwrites these accumulator columns after the next group barrier
}

// CTA prologue of the tensor path: TMEM allocation (64 columns per warpgroup), mbarriers, weights -> smem
__device__ __forceinline__ TcCtx tc_setup(MarchSmemTC& S, const DeviceModel& M, uint32_t debug_flags) {
    if (threadIdx.x < 32) tmem_alloc(&S.tmem_base, kTmemCols);
    if (threadIdx.x == 0) { for (int g = 0; g < kGroupsTC; ++g) mbar_init(&S.mbar[g], 1); fence_barrier_init(); }
    if (M.mlp_tc) {
        // weights already in the canonical layout (weights_to_canonical_kernel at load time): one 20 KB bulk copy per CTA
        // (cp.async.bulk, the TMA engine) instead of 1280 16-byte loads + re-indexed stores by the CTA's threads
        if (threadIdx.x == 0) {
            mbar_init(&S.wbar, 1);
            fence_barrier_init();
            const uint32_t bytes = kWTotal * (uint32_t)sizeof(__half);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&S.wbar)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(S.w)), "l"(M.mlp_tc), "r"(bytes), "r"(smem_u32(&S.wbar)) : "memory");
        }
        __syncthreads();                 // the barrier is initialised before anybody waits on it
        mbar_wait(&S.wbar, 0);
    } else {
        stage_weights_tc(S.w, M.mlp);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t g = threadIdx.x / kTile;
    TcCtx c;
    c.tmem = S.tmem_base + g * 64; c.a_addr = smem_u32(S.act[g]); c.w_addr = smem_u32(S.w); c.mbar = &S.mbar[g]; c.phase = 0;
    c.bar_id = 1 + g; c.row = threadIdx.x % kTile; c.swap = (debug_flags & kDebugSwapLboSbo) != 0;
    return c;
}
__device__ __forceinline__ void tc_teardown(MarchSmemTC& S) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(S.tmem_base, kTmemCols);
}

// =================================================================================================================
// march_kernel
// NerfTracer::trace's wavefront loop (S/ngp/testbed.cu:1938-2053) collapsed into one persistent kernel:
// generate_next_nerf_network_inputs -> network -> composite_kernel_nerf -> shade/accumulate/tonemap.
//
// Work layout: kRayLanes = 8 adjacent lanes form a RAY GROUP.  Each iteration the group generates the next 8 occupied
// samples of its ray (lane j keeps sample j), the 128 threads of a warpgroup evaluate their 128 samples as one network
// tile, and the group composites its (up to) 8 results strictly in order with the reference's per-sample rule, so pixels
// do not depend on the batching (the reference itself batches 1..8 samples per ray per iteration, S/ngp/testbed.cu:1996).
// Samples generated past a ray's termination are discarded, exactly like the reference's n_steps batches.
// Consecutive samples of one ray sit in adjacent lanes, so their hash-grid gathers share cache lines on the coarse and
// middle levels; all per-ray state is replicated in the 8 lanes (same arithmetic in every lane), so no state is exchanged.
// Groups pull new rays from the queue when their ray ends; tiles stay full until the queue drains.
// =================================================================================================================
#ifndef NMR_RAY_LANES
#define NMR_RAY_LANES 8
#endif
constexpr int kRayLanes = NMR_RAY_LANES;      // (4: measurement builds only - the 8-sample batch rule of the mesh surface needs 8)
#ifndef NMR_ENCODE_UNROLL
#define NMR_ENCODE_UNROLL 1
#endif
constexpr int kEncodeUnroll = NMR_ENCODE_UNROLL;   // hash-grid levels in flight per thread (8 gathers each)
constexpr int kWalkBudget = 6;      // empty voxels a ray may skip per tile iteration before it sits the iteration out

#ifndef NMR_MARCH_CTAS
#define NMR_MARCH_CTAS 3
#endif
// PASS2: the surface-ray pass of SchedArgs (variable batch sizes from the schedule); the main instantiation keeps the batch
// size a compile-time 8
template <bool TC, bool PASS2, bool OVERLAP = false>
__global__ void __launch_bounds__(TC ? kTile * kGroupsTC : kTile, TC ? NMR_MARCH_CTAS : 2) march_kernel(FrameParams P, DeviceModel M, const float4* __restrict__ queue,
                                                                                          uint32_t* __restrict__ counters, FrameOut out, uint32_t n_pixels, uint32_t debug_flags, const uint32_t* __restrict__ range_end, uint32_t* __restrict__ cursor, SchedArgs sched, int overlap_init_ctas) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = typename std::conditional<TC, MarchSmemTC, MarchSmem>::type;
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    // OVERLAPPED frame (overlap_init_ctas >= 0): this grid was launched with programmatic stream serialisation behind the set-up
    // kernel and runs next to it.  It never waits for that grid as a whole (no griddepcontrol.wait): every queue record carries a
    // ready word (kernels.cuh: kEmptyRecord) and counters[kCntInitDone] says when the queue is final.
    constexpr bool overlap = OVERLAP && !PASS2;       // (a compile-time variant: the serial kernel carries none of the extra state)
    if (threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMin(reinterpret_cast<unsigned long long*>(counters + kCntMarchStart), now);
        if (overlap) S.t_begin = now;          // (read only behind the tile barriers further down)
    }

    // this launch consumes the queue records [*cursor at launch, *range_end): the whole queue of a frame (range_end =
    // &counters[0], cursor = &counters[1]) or one band of it (nmr_render's copy-overlapped bands); the surface rule always
    // looks at the whole frame's live-ray count
    uint32_t n_rays = 0u, n_end = 0u;
    if (!overlap) { n_rays = counters[0]; n_end = *range_end; }
    // mesh surface insertion rule (SurfaceMode): the reference's 8-sample batches while <= 1/8 of the pixels are live
    const bool batch8 = P.surface_mode == kSurfaceBatch8 || (P.surface_mode == kSurfaceAuto && (unsigned long long)n_rays * 8ull <= (unsigned long long)n_pixels);
    // more than 1/8 live pixels under the auto rule: the reference's batch size varies per wavefront iteration (SchedArgs)
    const bool two_pass = sched.pass != 0 && !batch8 && P.surface_mode == kSurfaceAuto;
    constexpr bool pass2 = PASS2;
    if (pass2 && (!two_pass || counters[7] == 0u)) return;            // nothing to redo (whole grid, before any barrier)
    // The reference's wavefront loop (S/ngp/testbed.cu:1973-2047) replayed over the death histogram of pass 1: iteration k offers
    // n_k = clamp(pixels / live rays, 1, 8) samples to every live ray; a ray whose death index lies in [s_k, s_k + n_k) is
    // gone afterwards.  Every CTA of the surface-ray pass builds the boundaries s_0 .. s_K for itself (one thread, at most
    // ~1800 short iterations; close-ups only).
    uint16_t* sched_s = reinterpret_cast<uint16_t*>(smem_raw + sizeof(Smem));
    uint32_t sched_len = 0u;
    if (pass2) {
        uint32_t* len_s = reinterpret_cast<uint32_t*>(sched_s + kSchedMax + 2);
        if (threadIdx.x == 0) {
            uint32_t alive = n_rays - min(n_rays, counters[kCntLensRays]), s = 0, k = 0;      // the wavefront holds the ordinary rays only
            while (alive > 0u && k + 1u < kSchedMax && s < kSchedBins + 8u) {
                const uint32_t q = n_pixels / alive, n = q < 1u ? 1u : (q > 8u ? 8u : q);
                sched_s[k++] = (uint16_t)s;
                uint32_t dead = 0;
                for (uint32_t i = 0; i < n; ++i) dead += __ldg(sched.hist + min(s + i, kSchedBins - 1u));
                alive -= min(dead, alive);
                s += n;
            }
            sched_s[k] = (uint16_t)s;
            *len_s = k;
        }
        __syncthreads();
        sched_len = *len_s;
    }
    const bool batch_rule = batch8 || pass2;
    TcCtx tc;
    char* a_row;
    int enc_stride;
    if (TC) {
        MarchSmemTC& T = reinterpret_cast<MarchSmemTC&>(S);
        tc = tc_setup(T, M, debug_flags);
        a_row = reinterpret_cast<char*>(T.act[threadIdx.x / kTile]) + tc.row * 16;
        enc_stride = 2048;
    } else {
        MarchSmem& T = reinterpret_cast<MarchSmem&>(S);
        for (int i = threadIdx.x; i < kWTotal / 8; i += blockDim.x) reinterpret_cast<uint4*>(T.w)[i] = __ldg(reinterpret_cast<const uint4*>(M.mlp) + i);
        __syncthreads();
        a_row = reinterpret_cast<char*>(T.act) + threadIdx.x * 128;
        enc_stride = 16;
    }

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sub = lane & (kRayLanes - 1);             // which of the group's samples this lane evaluates
    const uint32_t gbase = lane & ~(uint32_t)(kRayLanes - 1);
    const uint32_t gmask = ((1u << kRayLanes) - 1u) << gbase;
    const V3 cam_origin = v3(P.ray_origin[0], P.ray_origin[1], P.ray_origin[2]);
    // lens rays park the state of their other segments here (one slot per ray group; all 8 lanes write the same values)
    float* __restrict__ stash = out.lens_scratch ? out.lens_scratch + ((size_t)blockIdx.x * (blockDim.x / kRayLanes) + threadIdx.x / kRayLanes) * kLensStash : nullptr;

    // per-ray state, identical in the 8 lanes of a group
    bool active = false, exhausted = false, pending_finish = false;
    bool waiting = false;        // overlapped frames: this group holds a queue slot whose record has not been written yet
    uint32_t my_slot = kEmptyRecord;
    V3 origin = cam_origin;      // primary rays start at the eye; the reflected segment of a lens ray starts on the lens
    uint32_t phase = 0;          // 0 ordinary ray; lens ray: 1 in front of the lens, 2 reflected segment, 3 behind the lens
    uint32_t kb = 0;             // surface-ray pass: schedule batch index of the ray's current batch
    bool sat = false;            // the ray's last batch ended on the opacity threshold
    V3 dir = v3(0.f, 0.f, 1.f);
    float t = 0.f, t_start = 0.f, t_surface = 0.f, t_limit = 0.f, max_weight = 0.f, depth = 0.f;
    float sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f;        // surface colour (mesh hand-off)
    float cr = 0.f, cg = 0.f, cb = 0.f, ca = 0.f;        // accumulated colour
    uint32_t idx = 0, n_samples = 0, evaluated = 0, n_batches = 0, n_passes = 0;

#ifdef NMR_PHASE_LOG_BUILD      // measurement build only (python nerf-glasses_b200/build.py -DNMR_PHASE_LOG_BUILD --out=...): FrameOut::phase_log
    uint32_t tile_iter = 0;
    unsigned long long* plog = nullptr;
    if (TC && !PASS2 && out.phase_log && (threadIdx.x % kTile) == 0)
        plog = out.phase_log + ((size_t)blockIdx.x * kGroupsTC + threadIdx.x / kTile) * kPhaseIters * kPhaseWords;
#define NMR_PLOG(k) do { if (plog && tile_iter < (uint32_t)kPhaseIters) { plog[tile_iter * kPhaseWords + (k)] = clock64(); if ((k) == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); plog[tile_iter * kPhaseWords + 5] = gt_; } } } while (0)
#define NMR_PLOG_NEXT() (++tile_iter)
#else
#define NMR_PLOG(k) do { } while (0)
#define NMR_PLOG_NEXT() do { } while (0)
#endif
    while (true) {
        NMR_PLOG(0);
        // ---- 1. next batch of up to 8 samples for this group's ray, pulling a new ray when the current one has ended ----
        V3 my_pos = v3(0.f, 0.f, 0.f);
        float my_dtw = 0.f, my_t_after = 0.f, t_batch_end = t;
        uint32_t n_valid = 0, batch_cap = kRayLanes;
        bool paused = false, ended = false;
        while (true) {
            if (pending_finish) {
                // composite_kernel_nerf's tail for a finished ray (S/ngp/testbed.cu:886-901), then shade/accumulate/tonemap
                if (sw > 0) { const float k = 1.f - ca; cr += sr * k; cg += sg * k; cb += sb * k; ca += sw * k; }
                pending_finish = false;
                if (phase == 1u || phase == 2u) {
                    // ---- lens ray: the segment in front of the lens (1) or the reflected segment (2) has ended ----
                    if (phase == 1u) {
                        // split at the lens: park what was accumulated in front of it and where the primary walk stands
                        const float4 l0 = __ldg(out.lens + (size_t)(idx & ~kLensRayFlag) * 2);
                        V3 refl;
                        const float F = lens_fresnel(P, dir, v3(l0.x, l0.y, l0.z), refl);
                        const float wl = __ldg(out.lens + (size_t)(idx & ~kLensRayFlag) * 2 + 1).x;
                        stash[0] = cr; stash[1] = cg; stash[2] = cb; stash[3] = ca; stash[14] = F; stash[15] = wl;
                        stash[4] = 0.f; stash[5] = 0.f; stash[6] = 0.f;
                        {   // where the transmitted segment starts (behind the pane, shifted sideways under the plate model)
                            V3 delta; float t_behind;
                            lens_plate_shift(P, dir, v3(l0.x, l0.y, l0.z), l0.w, delta, t_behind);
                            stash[7] = fmaxf(t, t_behind);
                            stash[20] = cam_origin.x + delta.x; stash[21] = cam_origin.y + delta.y; stash[22] = cam_origin.z + delta.z;
                        }
                        if (wl * F * (1.f - ca) >= 1.0f / 512.0f) {
                            // reflected segment: a fresh ray from the hit point along the mirror direction, NeRF only
                            stash[16] = dir.x; stash[17] = dir.y; stash[18] = dir.z;
                            origin = vadd(cam_origin, vmul(dir, l0.w));
                            dir = refl;
                            float t_in;
                            t_limit = occupied_exit(P, origin, dir, t_in);
                            t = 1e-3f; t_start = 0.f; t_surface = 0.f; sr = sg = sb = sw = 0.f;
                            cr = cg = cb = ca = 0.f;
                            phase = 2u;
                            continue;
                        }
                        stash[16] = dir.x; stash[17] = dir.y; stash[18] = dir.z;
                    } else {
                        // what the mirror direction sees: the march result over the background colour
                        const float k = (1.f - ca) * P.background[3];
                        stash[4] = cr + P.background[0] * k; stash[5] = cg + P.background[1] * k; stash[6] = cb + P.background[2] * k;
                    }
                    // transmitted segment: the primary ray carries on behind the lens, now with the opaque mesh surface (if any)
                    origin = v3(stash[20], stash[21], stash[22]);
                    dir = v3(stash[16], stash[17], stash[18]);
                    t = stash[7]; t_start = stash[8]; t_surface = stash[9]; sr = stash[10]; sg = stash[11]; sb = stash[12]; sw = stash[13]; t_limit = stash[19];
                    if (P.lens_model == 1) { float t_in; t_limit = occupied_exit(P, origin, dir, t_in); }      // (the shifted ray's own exit from the occupied box)
                    cr = cg = cb = ca = 0.f;
                    phase = 3u;
                    continue;
                }
                if (phase == 3u) {
                    // combine: front + T_front * ( w (F * reflected + (1 - F) k * behind) + (1 - w) * behind )
                    const float F = stash[14], wl = stash[15], Tf = 1.f - stash[3];
                    const float kr = (1.f - F) * P.lens_k[0], kg = (1.f - F) * P.lens_k[1], kb = (1.f - F) * P.lens_k[2];
                    const float a_behind = ca;
                    cr = stash[0] + Tf * (wl * (F * stash[4] + kr * cr) + (1.f - wl) * cr);
                    cg = stash[1] + Tf * (wl * (F * stash[5] + kg * cg) + (1.f - wl) * cg);
                    cb = stash[2] + Tf * (wl * (F * stash[6] + kb * cb) + (1.f - wl) * cb);
                    ca = stash[3] + Tf * (wl * (F + (1.f - F) * ((1.f - P.lens_kmean) + P.lens_kmean * a_behind)) + (1.f - wl) * a_behind);
                    phase = 0u;
                }
                if (two_pass && !pass2) {
                    // first of two passes: record where this ray dies in the reference's wavefront (sample index of the batch
                    // that kills it); rays that carry a mesh surface are composited again by the second pass
                    // (lens rays are marched on their own, outside the wavefront whose schedule is being replayed: oracle render_impl)
                    if (sub == 0 && !(idx & kLensRayFlag)) atomicAdd(sched.hist + min(sat ? n_samples - 1u : n_samples, kSchedBins - 1u), 1u);
                    if (idx & kSurfRayFlag) { active = false; continue; }
                }
                if (sub == 0) finish_pixel(P, out, idx & ~(kSurfRayFlag | kLensRayFlag), cr, cg, cb, ca, depth, n_samples);
                active = false;
            }
            if (!active) {
                if (exhausted) break;
                uint32_t slot = 0;
                if (!overlap || my_slot == kEmptyRecord) {
                    if (sub == 0) slot = atomicAdd(cursor, 1u);
                    slot = __shfl_sync(gmask, slot, gbase);
                    if (overlap) my_slot = slot;
                } else {
                    slot = my_slot;
                }
                NMR_DEVICE_CHECK(slot < (uint32_t)(P.width * P.height) + kQueueSlack);
                if (overlap) {
                    waiting = false;
                    // the record is there when its ready word is; when it is not and the set-up kernel has finished, re-read once (a
                    // record committed before the last CTA signed off is visible by now): still empty = beyond the end of the queue
                    const uint32_t* rw = reinterpret_cast<const uint32_t*>(queue + (size_t)slot * kRayRecordFloat4s + 1) + 2;
                    uint32_t w;
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(w) : "l"(rw) : "memory");
                    if (w == kEmptyRecord) {
                        uint32_t done;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(counters + kCntInitDone) : "memory");
                        const bool fin = done >= (uint32_t)overlap_init_ctas;
                        if (fin) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(w) : "l"(rw) : "memory");
                        if (w == kEmptyRecord) { exhausted = fin; waiting = !fin; break; }
                    }
                } else if (slot >= n_end) { exhausted = true; break; }
                if (overlap) my_slot = kEmptyRecord;
                if (pass2) slot = __ldg(sched.surf_list + slot);
                NMR_DEVICE_CHECK(slot < (uint32_t)(P.width * P.height));
                // (L2 loads: the records may have been written while this kernel was running)
                const float4 q0 = __ldcg(queue + (size_t)slot * kRayRecordFloat4s), q1 = __ldcg(queue + (size_t)slot * kRayRecordFloat4s + 1), q2 = __ldcg(queue + (size_t)slot * kRayRecordFloat4s + 2);
                dir = v3(q0.x, q0.y, q0.z); t = q0.w; t_start = q1.x; t_surface = q1.y; idx = __float_as_uint(q1.z); t_limit = q1.w;
                sr = q2.x; sg = q2.y; sb = q2.z; sw = q2.w;
                cr = cg = cb = ca = 0.f; max_weight = 0.f; depth = 0.f; n_samples = 0;
                origin = cam_origin;       // (the group's previous ray may have been a lens ray under the plate model: its transmitted segment starts beside the eye)
                active = true;
                phase = 0u; kb = 0u; sat = false;
                if (idx & kLensRayFlag) {
                    // lens ray, first segment: samples up to the lens only, the opaque mesh surface waits behind it
                    // (the flag stays in idx for the ray's lifetime: a lens ray is not part of the wavefront schedule)
                    stash[8] = t_start; stash[9] = t_surface; stash[10] = sr; stash[11] = sg; stash[12] = sb; stash[13] = sw; stash[19] = t_limit;
                    t_limit = fminf(t_limit, __ldg(out.lens + (size_t)(idx & ~kLensRayFlag) * 2).w);
                    t_surface = 0.f; sr = sg = sb = sw = 0.f;
                    phase = 1u;
                }
            }
            const V3 idir = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
            // Batch generation.  The reference walks samples one after the other (S/ngp/testbed.cu:596-629); inside an object
            // nearly every step lands in an occupied cell, so the 8 lanes first test 8 consecutive steps IN PARALLEL (lane j
            // replays the t += dt recurrence j times - same float operations, no memory access - and then does the one
            // occupancy load of its own sample).  The accepted prefix is exactly what the sequential walk produces; at the
            // first step that is not a plain "occupied, inside, not behind the mesh" step the group falls back to the
            // sequential rule for that one sample and then tries the parallel test again.
            float tt = t;
            n_valid = 0;
            paused = false; ended = false;
            uint32_t cap = kRayLanes;
            if (pass2 && sw > 0.f && kb < sched_len) cap = (uint32_t)sched_s[kb + 1] - (uint32_t)sched_s[kb];   // the reference's n_steps of this iteration
            batch_cap = cap;
#pragma unroll 1
            while (n_valid < cap) {
                ++n_passes;
                float tc = tt, dtc = 0.f;
                V3 pc = v3(0.f, 0.f, 0.f);
                bool ok = sub >= n_valid && sub < cap;
                if (ok) {
                    for (uint32_t k = n_valid; k < sub; ++k) tc += calc_dt(tc - t_start, P.cone_angle);
                    ok = !(t_surface != 0.0f && tc > t_surface && sw == 1.f) && !(tc > t_limit);
                    pc = vadd(origin, vmul(dir, tc));
                    ok = ok && box_contains(P.aabb_min, P.aabb_max, r2l_mul(P.r2l, pc));
                    if (ok) {
                        dtc = calc_dt(tc - t_start, P.cone_angle);
                        ok = occupied_at(pc, M.bitfield, (uint32_t)mip_from_dt(dtc, pc));
                    }
                }
                const uint32_t fail = (__ballot_sync(gmask, !ok) >> gbase) & ~((1u << n_valid) - 1u) & ((1u << kRayLanes) - 1u);
                const uint32_t first_fail = fail ? (uint32_t)(__ffs((int)fail) - 1) : (uint32_t)kRayLanes;
                if (sub >= n_valid && sub < first_fail) {
                    const V3 diag = v3(P.taabb_max[0] - P.taabb_min[0], P.taabb_max[1] - P.taabb_min[1], P.taabb_max[2] - P.taabb_min[2]);
                    my_pos = v3((pc.x - P.taabb_min[0]) / diag.x, (pc.y - P.taabb_min[1]) / diag.y, (pc.z - P.taabb_min[2]) / diag.z);
                    my_dtw = warp_dt(dtc);
                    my_t_after = tc + dtc;
                }
                if (first_fail > n_valid) {        // advance the ray past the accepted prefix
                    tt = __shfl_sync(gmask, tc + dtc, gbase + first_fail - 1);
                    n_valid = first_fail;
                }
                if (n_valid >= cap) break;
                // sample number n_valid needs the general rule (empty-space skip, box exit or opaque mesh surface).  A long walk
                // through empty cells is cut into slices of kWalkBudget voxels so that one ray cannot stall its tile: a paused
                // walk keeps its state in t and resumes in the next iteration.
                Sample smp;
                // (a ray that still carries a mesh surface under the batch rule never pauses: its batches must stay aligned to 8 samples)
                const int rc = next_sample(P, M.bitfield, origin, dir, idir, t_start, t_surface, sw, t_limit, false, (batch_rule && sw > 0.f) ? 0x7fffffff : kWalkBudget, tt, smp);
                if (rc != 1) { paused = rc == 2; ended = rc == 0; break; }
                if (sub == n_valid) { my_pos = smp.pos; my_dtw = smp.dt_warped; my_t_after = tt; }
                ++n_valid;
            }
            t_batch_end = tt;
            if (n_valid > 0 || paused) break;
            pending_finish = true;      // the batch came back empty: the ray has ended
            t = tt;                     // (where its walk stood - the transmitted segment of a lens ray resumes there)
        }
        const bool have = active && !pending_finish && sub < n_valid;
        if (have && sub == 0) ++n_batches;

        NMR_PLOG(1);
        // ---- 2. encode straight into this thread's row of the A operand ----
        // A full warp encodes 32 samples, one per lane, 16 dependent gather rounds each.  In the sparse iterations at the end of a
        // launch a warp holds few samples and the rounds are pure latency: its lanes then SHARE the samples - with n samples in the
        // warp, 32 / n' lanes (n' = n rounded up to a power of two) take the levels of one sample between them and write them
        // into that sample's row of the A tile (rows of a warp's lanes are 16 bytes apart).
        if (TC && kEncodeUnroll == 1 && !(debug_flags & kDebugNoSharedEncode)) {
            const uint32_t hm = __ballot_sync(0xffffffffu, have);
            const uint32_t n_have = (uint32_t)__popc(hm);
            const uint32_t share = n_have > 16u ? 1u : (n_have > 8u ? 2u : (n_have > 4u ? 4u : (n_have > 2u ? 8u : 16u)));   // lanes per sample
            uint32_t src = lane, first = 0u;
            bool work = have;
            if (share > 1u) {
                const uint32_t k = lane / share;                    // which of the warp's samples this lane helps with
                first = lane % share;
                work = k < n_have;
                src = work ? __fns(hm, 0u, (int)k + 1) : lane;
            }
            const V3 p = v3(__shfl_sync(0xffffffffu, my_pos.x, src), __shfl_sync(0xffffffffu, my_pos.y, src), __shfl_sync(0xffffffffu, my_pos.z, src));
            if (work) encode_levels_strided(M, p, a_row + ((int)src - (int)lane) * 16, enc_stride, first, share);
            if (share > 1u) __syncwarp();
            if (have) ++evaluated;
        } else if (have) { encode_chunks<kEncodeUnroll>(M, my_pos, a_row, enc_stride); ++evaluated; }
        // tile-wide decisions: run the network when any lane has a sample; leave only when no ray is left (a tile whose
        // rays are all in the middle of a paused empty-space walk has no sample this iteration but must keep going)
        const bool any_have = TC ? group_any(tc.bar_id, have) : (__syncthreads_or(have ? 1 : 0) != 0);
        if (!any_have) {
            const bool any_active = TC ? group_any(tc.bar_id, active || (overlap && waiting)) : (__syncthreads_or((active || (overlap && waiting)) ? 1 : 0) != 0);
            if (!any_active) break;
            if (active && !pending_finish) t = t_batch_end;
            if (overlap && waiting) {                      // nothing to do in this tile until the set-up kernel queues more rays
                __nanosleep(200);
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                // never hang the GPU on a producer that does not deliver (2 s: a frame takes milliseconds): give the slot up
                if (now - S.t_begin > 2000000000ull) { waiting = false; exhausted = true; }
            }
            continue;
        }

        NMR_PLOG(2);     // (after the tile-wide barrier: the slowest lane's encoding)
        // ---- 3. network ----
        float raw[4];
        const V3 dir01 = v3((dir.x + 1.0f) * 0.5f, (dir.y + 1.0f) * 0.5f, (dir.z + 1.0f) * 0.5f);   // warp_direction
        if (TC) network_tc(a_row, tc, dir01, raw);
        else network_scalar(reinterpret_cast<MarchSmem&>(S), dir01, raw);

        NMR_PLOG(3);
        // ---- 4. composite the batch in order (S/ngp/testbed.cu:830-884) ----
        // Every lane first turns ITS OWN sample's network outputs into alpha / colour / depth candidate (the expensive part:
        // an exponential, three logistics, a square root); the group then replays the reference's per-sample recurrence
        // over those values, strictly in order, so pixels do not depend on how samples were batched.
        float a_alpha = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f, a_depth = 0.f;
        if (have) {
            a_alpha = 1.f - __expf(-act_density(raw[3], P.density_activation) * unwarp_dt(my_dtw));
            a_r = act_rgb(raw[0], P.rgb_activation); a_g = act_rgb(raw[1], P.rgb_activation); a_b = act_rgb(raw[2], P.rgb_activation);
            const V3 tdiag = v3(P.taabb_max[0] - P.taabb_min[0], P.taabb_max[1] - P.taabb_min[1], P.taabb_max[2] - P.taabb_min[2]);
            const V3 pos = v3(P.taabb_min[0] + my_pos.x * tdiag.x, P.taabb_min[1] + my_pos.y * tdiag.y, P.taabb_min[2] + my_pos.z * tdiag.z);
            const V3 dd = vsub(pos, v3(P.cam[9], P.cam[10], P.cam[11]));   // NeRF-space sample vs world-space eye, as in the reference
            a_depth = sqrtf(edot(dd, dd));
        }
#ifdef NMR_CHECKED
        if (P.debug_pixel >= 0 && active && sub == 0 && (idx & ~(kSurfRayFlag | kLensRayFlag)) == (uint32_t)P.debug_pixel)
            printf("[gpu ray %d] batch: phase %u n_valid %u ended %d paused %d pending_finish %d t %.7f t_batch_end %.7f t_limit %.7f t_surface %.7f sw %.3f ca %.5f n_samples %u\n", P.debug_pixel, phase, n_valid, (int)ended, (int)paused, (int)pending_finish, t, t_batch_end, t_limit, t_surface, sw, ca, n_samples);
#endif
        if (active && !pending_finish) {
            bool done = false;
            // reference rule: the batch's end (payload.t after generate_next_nerf_network_inputs) decides, before its first sample
            // (payload.t is only advanced by a FULL batch, S/ngp/testbed.cu:620-633: a batch that came back short is judged by where
            // it started.  An ordinary ray never starts a batch behind its surface with the surface still pending; the transmitted
            // segment of a lens ray does when the opaque surface lies less than a step behind the lens.)
            const bool full_batch = n_valid == (pass2 ? batch_cap : (uint32_t)kRayLanes);
            const bool pre_blend = batch_rule && sw > 0.f && (full_batch ? t_batch_end : t) > t_surface;
#pragma unroll 1
            for (uint32_t j = 0; j < n_valid; ++j) {
                const uint32_t src = gbase + j;
                const float alpha = __shfl_sync(gmask, a_alpha, src);
                const float r0 = __shfl_sync(gmask, a_r, src), r1 = __shfl_sync(gmask, a_g, src), r2 = __shfl_sync(gmask, a_b, src);
                ++n_samples;
                float T = 1.f - ca;
                if (sw > 0.f) {        // same in all lanes of the group
                    const bool insert = batch_rule ? (j == 0 && pre_blend) : (__shfl_sync(gmask, my_t_after, src) > t_surface);
                    if (insert) {
                        cr += sr * sw * T; cg += sg * sw * T; cb += sb * sw * T; ca += sw * T;
                        sw = 0.f;
                        T = 1.f - ca;
                        if (ca > 0.99f) { const float a = ca; cr /= a; cg /= a; cb /= a; ca /= a; done = true; break; }
                    }
                }
                const float weight = alpha * T;
                cr += r0 * weight; cg += r1 * weight; cb += r2 * weight;
                ca += weight;
                if (weight > max_weight) { max_weight = weight; depth = __shfl_sync(gmask, a_depth, src); }
                if (ca > (1.0f - P.min_transmittance)) { const float a = ca; cr /= a; cg /= a; cb /= a; ca /= a; done = true; break; }
            }
            // a batch that came back short because the walk ended (box exit, opaque mesh surface) is the ray's last one: the
            // reference kills the ray when it produced fewer than n_steps samples (S/ngp/testbed.cu:886-901)
            if (done && phase == 1u) phase = 0u;      // saturated in front of the lens: an ordinary ray after all
            sat = done; if (pass2) ++kb;
            if (done || ended) pending_finish = true;
            t = t_batch_end;           // the walk's state: resume point of a paused empty-space walk or of a lens ray's next segment
        }
        NMR_PLOG(4);
        NMR_PLOG_NEXT();
    }
    NMR_PLOG(0);      // (exit stamp)

    // evaluated-sample counter: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) evaluated += __shfl_xor_sync(0xffffffffu, evaluated, o);
    if (lane == 0 && evaluated) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), (unsigned long long)evaluated);
    if (sub == 0 && n_batches) { atomicAdd(&counters[4], n_batches); atomicAdd(&counters[5], n_passes); }
    if (threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicMax(reinterpret_cast<unsigned long long*>(counters + kCntMarchEnd), now);
        // how many records this frame queued, for the next frame's clear kernel (every CTA stores the same final value; the
        // set-up kernel has finished by the time a CTA gets here)
        if (!PASS2) counters[kCntPrevCount] = overlap ? *reinterpret_cast<volatile uint32_t*>(counters) : max(n_rays, n_end);
    }
    if (TC) tc_teardown(reinterpret_cast<MarchSmemTC&>(S));
}

constexpr size_t kSchedSmem = (kSchedMax + 2) * sizeof(uint16_t) + 16;   // schedule boundaries + their count behind the surface-ray pass's tile memory
constexpr int kMarchCtasPerSm = NMR_MARCH_CTAS;   // __launch_bounds__(256, 3): 24 warps, 3 x 128 tensor-memory columns, 3 x 53 KB shared memory per SM

namespace {
// <<<>>> or, for an overlapped frame, cudaLaunchKernelEx with programmatic stream serialisation (the kernel may start while the
// set-up kernel in front of it is still running)
template <typename K, typename... Args>
void launch_march_variant(K kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool programmatic, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = programmatic ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}
}  // namespace

void launch_march(const FrameParams& P, const DeviceModel& M, const float4* d_queue, uint32_t* d_counters, const FrameOut& out,
                  uint32_t n_pixels, uint32_t debug_flags, int num_sms, cudaStream_t s, const uint32_t* d_range_end, uint32_t* d_cursor,
                  const SchedArgs* sched, int ctas_per_sm, int overlap_init_ctas) {
    if (!d_range_end) d_range_end = d_counters;
    if (!d_cursor) d_cursor = d_counters + 1;
    SchedArgs sa{};
    if (sched) sa = *sched;
    const int per_sm = ctas_per_sm > 0 ? ctas_per_sm : kMarchCtasPerSm;
    if (debug_flags & kDebugScalarMlp) {
        static std::atomic<bool> attr_set_dev[64];     // the opt-in to > 48 KB of dynamic shared memory is per device (setting it twice from two threads is harmless)
        int dev = 0; cudaGetDevice(&dev);
        std::atomic<bool>& attr_set = attr_set_dev[dev & 63];
        if (!attr_set.load(std::memory_order_acquire)) { cudaFuncSetAttribute(march_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem)); cudaFuncSetAttribute(march_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(MarchSmem) + kSchedSmem)); attr_set.store(true, std::memory_order_release); }
        if (sa.pass == 2) launch_march_variant(march_kernel<false, true>, num_sms, kTile, sizeof(MarchSmem) + kSchedSmem, s, false, P, M, d_queue, d_counters, out, n_pixels, debug_flags, d_range_end, d_cursor, sa, -1);
        else launch_march_variant(march_kernel<false, false>, num_sms * 2, kTile, sizeof(MarchSmem), s, false, P, M, d_queue, d_counters, out, n_pixels, debug_flags, d_range_end, d_cursor, sa, -1);
    } else {
        static std::atomic<bool> attr_set_dev[64];
        int dev = 0; cudaGetDevice(&dev);
        std::atomic<bool>& attr_set = attr_set_dev[dev & 63];
        if (!attr_set.load(std::memory_order_acquire)) { cudaFuncSetAttribute(march_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmemTC)); cudaFuncSetAttribute(march_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmemTC)); cudaFuncSetAttribute(march_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(MarchSmemTC) + kSchedSmem)); attr_set.store(true, std::memory_order_release); }
        if (sa.pass == 2) launch_march_variant(march_kernel<true, true>, num_sms * per_sm, kTile * kGroupsTC, sizeof(MarchSmemTC) + kSchedSmem, s, false, P, M, d_queue, d_counters, out, n_pixels, debug_flags, d_range_end, d_cursor, sa, -1);
        else if (overlap_init_ctas >= 0) launch_march_variant(march_kernel<true, false, true>, num_sms * per_sm, kTile * kGroupsTC, sizeof(MarchSmemTC), s, true, P, M, d_queue, d_counters, out, n_pixels, debug_flags, d_range_end, d_cursor, sa, overlap_init_ctas);
        else launch_march_variant(march_kernel<true, false>, num_sms * per_sm, kTile * kGroupsTC, sizeof(MarchSmemTC), s, false, P, M, d_queue, d_counters, out, n_pixels, debug_flags, d_range_end, d_cursor, sa, -1);
    }
}

// =================================================================================================================
// measurement helper: what the L2 delivers to the SMs on this GPU (bench.py's roofline denominators; not part of a render)
//   mode 0: coalesced 16-byte loads over a buffer that fits the L2 (the L2 -> SM bandwidth ceiling)
//   mode 1: independent 4-byte gathers at hashed indices, eight in flight per thread - the access pattern of the hashed
//           levels of the grid encoder with no arithmetic around it (each gather moves a 32-byte sector): the ceiling of
//           "algorithmic gather bytes per second" for that pattern
// =================================================================================================================
__global__ void __launch_bounds__(256) l2_probe_kernel(const uint4* __restrict__ buf, uint32_t n_vec, uint32_t loads_per_thread, int mode, uint32_t* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    uint32_t acc = 0;
    if (mode == 0) {
        uint32_t i = tid % n_vec;
        for (uint32_t k = 0; k < loads_per_thread; k += 4) {
            uint4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { v[j] = __ldcg(buf + i); i += n_threads; if (i >= n_vec) i -= n_vec; }
#pragma unroll
            for (int j = 0; j < 4; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
        }
    } else {
        const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>(buf);
        const uint32_t mask = n_vec * 4u - 1u;                     // n_vec is a power of two in this mode
        uint32_t x = tid * 2654435761u + 12345u;
        for (uint32_t k = 0; k < loads_per_thread; k += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { x = x * 1664525u + 1013904223u; v[j] = __ldg(w + ((x >> 4) & mask)); }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc ^= v[j];
        }
    }
    if (acc == 0x9E3779B9u) *sink = acc;      // keeps the loads alive
}
void launch_l2_probe(const void* d_buf, uint32_t n_vec, uint32_t loads_per_thread, int mode, uint32_t* d_sink, int num_sms, cudaStream_t s) {
    l2_probe_kernel<<<num_sms * 8, 256, 0, s>>>(reinterpret_cast<const uint4*>(d_buf), n_vec, loads_per_thread, mode, d_sink);
}

// =================================================================================================================
// parity probes
// =================================================================================================================
__global__ void debug_encode_kernel(DeviceModel M, const float* __restrict__ pos, int64_t n, uint16_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // the march kernel's own path (bricked coarse levels when the model has them; nmr_debug_set_flags bit 4 hides them for A/B checks)
    encode_levels_strided(M, v3(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]), reinterpret_cast<char*>(out + i * ENC_WIDTH), 16, 0u, 1u);
}
void launch_debug_encode(const DeviceModel& M, const float* d_pos, int64_t n, uint16_t* d_out, cudaStream_t s) {
    if (n <= 0) return;
    debug_encode_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(M, d_pos, n, d_out);
}

template <bool TC>
__global__ void __launch_bounds__(TC ? kTile * kGroupsTC : kTile) debug_network_kernel(DeviceModel M, const float* __restrict__ pos, const float* __restrict__ dir, int64_t n, uint16_t* __restrict__ out4, uint32_t debug_flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = typename std::conditional<TC, MarchSmemTC, MarchSmem>::type;
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    TcCtx tc;
    char* a_row;
    int enc_stride;
    if (TC) {
        MarchSmemTC& T = reinterpret_cast<MarchSmemTC&>(S);
        tc = tc_setup(T, M, debug_flags);
        a_row = reinterpret_cast<char*>(T.act[threadIdx.x / kTile]) + tc.row * 16;
        enc_stride = 2048;
    } else {
        MarchSmem& T = reinterpret_cast<MarchSmem&>(S);
        for (int i = threadIdx.x; i < kWTotal / 8; i += blockDim.x) reinterpret_cast<uint4*>(T.w)[i] = __ldg(reinterpret_cast<const uint4*>(M.mlp) + i);
        __syncthreads();
        a_row = reinterpret_cast<char*>(T.act) + threadIdx.x * 128;
        enc_stride = 16;
    }
    // every warpgroup walks the same number of tiles, so its barriers stay matched
    const int64_t per_cta = blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * per_cta; base < n; base += (int64_t)gridDim.x * per_cta) {
        const int64_t i = base + threadIdx.x;
        const bool have = i < n;
        V3 d01 = v3(0.5f, 0.5f, 0.5f);
        if (have) {
            encode_levels_strided(M, v3(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]), a_row, enc_stride, 0u, 1u);
            d01 = v3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]);
        }
        float raw[4];
        if (TC) network_tc(a_row, tc, d01, raw);
        else network_scalar(reinterpret_cast<MarchSmem&>(S), d01, raw);
        if (have) {
            for (int k = 0; k < 4; ++k) { const __half h = __float2half_rn(raw[k]); out4[i * 4 + k] = *reinterpret_cast<const uint16_t*>(&h); }
        }
    }
    if (TC) tc_teardown(reinterpret_cast<MarchSmemTC&>(S));
}
void launch_debug_network(const DeviceModel& M, const float* d_pos, const float* d_dir, int64_t n, uint16_t* d_out4, uint32_t debug_flags, cudaStream_t s) {
    if (n <= 0) return;
    if (debug_flags & kDebugScalarMlp) {
        const unsigned blocks = (unsigned)((n + kTile - 1) / kTile < 296 ? (n + kTile - 1) / kTile : 296);
        cudaFuncSetAttribute(debug_network_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmem));
        debug_network_kernel<false><<<blocks, kTile, sizeof(MarchSmem), s>>>(M, d_pos, d_dir, n, d_out4, debug_flags);
    } else {
        const int per = kTile * kGroupsTC;
        const unsigned blocks = (unsigned)((n + per - 1) / per < 296 ? (n + per - 1) / per : 296);
        cudaFuncSetAttribute(debug_network_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmemTC));
        debug_network_kernel<true><<<blocks, per, sizeof(MarchSmemTC), s>>>(M, d_pos, d_dir, n, d_out4, debug_flags);
    }
}

// ---- density probes of the collision tool ---------------------------------------------------------------------
// NerfTracer::intersects (mode 0, S/ngp/testbed.cu:1891-1935) and NerfTracer::collide + check_collision (mode 1, :1814-1888,
// :721-782) over the payloads NerfMeshRenderer::collide writes (origin = world point + 0.5, one direction for all,
// t = t_start = 0, S/nerf_mesh_renderer.cu:1564-1574).  One thread per point / ray; a warpgroup evaluates its 128 samples as
// one tcgen05 tile per step, exactly like the march kernel.  The reference marches in batches of 8 and checks the batch in
// order, so "first sample with alpha > 0" does not depend on the batching; a ray that leaves the render box keeps distance 0.
__global__ void __launch_bounds__(kTile * kGroupsTC) probe_kernel(FrameParams P, DeviceModel M, const float* __restrict__ points_world, float dx, float dy, float dz,
                                                                  int64_t n, int mode, float* __restrict__ out, uint32_t debug_flags) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MarchSmemTC& T = *reinterpret_cast<MarchSmemTC*>(smem_raw);
    TcCtx tc = tc_setup(T, M, debug_flags);
    char* a_row = reinterpret_cast<char*>(T.act[threadIdx.x / kTile]) + tc.row * 16;
    const V3 dir = v3(dx, dy, dz);
    const V3 idir = v3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
    const V3 dir01 = v3((dir.x + 1.0f) * 0.5f, (dir.y + 1.0f) * 0.5f, (dir.z + 1.0f) * 0.5f);
    const V3 diag = v3(P.taabb_max[0] - P.taabb_min[0], P.taabb_max[1] - P.taabb_min[1], P.taabb_max[2] - P.taabb_min[2]);
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    for (int64_t tile = (int64_t)blockIdx.x * kGroupsTC + threadIdx.x / kTile; tile < n_tiles; tile += (int64_t)gridDim.x * kGroupsTC) {
        const int64_t i = tile * kTile + tc.row;
        bool active = i < n;
        V3 origin = v3(0.5f, 0.5f, 0.5f);
        if (active) origin = v3(points_world[i * 3] + 0.5f, points_world[i * 3 + 1] + 0.5f, points_world[i * 3 + 2] + 0.5f);
        float t = 0.f, result = 0.f;
        while (group_any(tc.bar_id, active)) {
            bool have = false;
            V3 wpos = v3(0.f, 0.f, 0.f);
            float dtw = 0.f;
            if (active) {
                if (mode == 0) {
                    wpos = v3((origin.x - P.taabb_min[0]) / diag.x, (origin.y - P.taabb_min[1]) / diag.y, (origin.z - P.taabb_min[2]) / diag.z);
                    have = true;
                } else {
                    Sample smp;
                    if (next_sample(P, M.bitfield, origin, dir, idir, 0.f, 0.f, 0.f, 3.402823466e+38f, true, 0x7fffffff, t, smp) == 1) { wpos = smp.pos; dtw = smp.dt_warped; have = true; }
                    else active = false;       // left the render box without a collision
                }
            }
            if (!group_any(tc.bar_id, have)) continue;
            if (have) encode_levels_strided(M, wpos, a_row, 2048, 0u, 1u);
            float raw[4];
            network_tc(a_row, tc, dir01, raw);
            if (have) {
                const V3 pos = v3(P.taabb_min[0] + wpos.x * diag.x, P.taabb_min[1] + wpos.y * diag.y, P.taabb_min[2] + wpos.z * diag.z);   // unwarp_position
                if (mode == 0) {
                    const float dt = min_cone_stepsize();
                    const float alpha = 1.f - __expf(-act_density(raw[3], P.density_activation) * dt);
                    const int mip = max(0, mip_from_dt(dt, pos));
                    if (occupied_at(pos, M.bitfield, (uint32_t)mip)) result = alpha;
                    active = false;
                } else {
                    const float alpha = 1.f - __expf(-act_density(raw[3], P.density_activation) * unwarp_dt(dtw));
                    if (alpha > 0.f) { const V3 dd = vsub(pos, origin); result = sqrtf(edot(dd, dd)); active = false; }
                }
            }
        }
        if (i < n) out[i] = result;
    }
    tc_teardown(T);
}
void launch_probe(const FrameParams& P, const DeviceModel& M, const float* d_points_world, const float dir[3], int64_t n, int mode, float* d_out,
                  uint32_t debug_flags, int num_sms, cudaStream_t s) {
    if (n <= 0) return;
    static std::atomic<bool> attr_set_dev[64];
    int dev = 0; cudaGetDevice(&dev);
    std::atomic<bool>& attr_set = attr_set_dev[dev & 63];
    if (!attr_set.load(std::memory_order_acquire)) { cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MarchSmemTC)); attr_set.store(true, std::memory_order_release); }
    const int64_t per = kTile * kGroupsTC, want = (n + per - 1) / per, cap = (int64_t)num_sms * 2;
    probe_kernel<<<(unsigned)(want < cap ? want : cap), kTile * kGroupsTC, sizeof(MarchSmemTC), s>>>(P, M, d_points_world, dir[0], dir[1], dir[2], n, mode, d_out, debug_flags);
}

__global__ void debug_trace_kernel(FrameParams P, DeviceModel M, const uint32_t* __restrict__ pixels, int64_t n_pix, uint32_t max_samples,
                                   float* __restrict__ o_t, uint32_t* __restrict__ o_cell, uint32_t* __restrict__ o_mip, float* __restrict__ o_pos,
                                   uint32_t* __restrict__ o_count, float* __restrict__ o_ray) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    const uint32_t pix = pixels[i];
    const uint32_t x = pix % (uint32_t)P.width, y = pix / (uint32_t)P.width;
    RayInit r = init_ray(P, x, y);
    float t = r.t, t_start;
    const bool alive = advance_pos(P, M.bitfield, M.coarse, r.origin, r.dir, pix, 0.f, r.t_occ_in, r.t_limit, r.alive, t, t_start);
    float* rr = o_ray + i * 8;
    rr[0] = r.origin.x; rr[1] = r.origin.y; rr[2] = r.origin.z; rr[3] = r.dir.x; rr[4] = r.dir.y; rr[5] = r.dir.z; rr[6] = t; rr[7] = alive ? 1.f : 0.f;
    const V3 idir = v3(1.0f / r.dir.x, 1.0f / r.dir.y, 1.0f / r.dir.z);
    uint32_t cnt = 0;
    while (alive && cnt < max_samples) {
        Sample s;
        if (next_sample(P, M.bitfield, r.origin, r.dir, idir, t_start, 0.f, 0.f, r.t_limit, true, 0x7fffffff, t, s) != 1) break;
        const int64_t o = i * max_samples + cnt;
        o_t[o] = s.t; o_cell[o] = s.cell; o_mip[o] = s.mip;
        o_pos[o * 3] = s.pos.x; o_pos[o * 3 + 1] = s.pos.y; o_pos[o * 3 + 2] = s.pos.z;
        ++cnt;
    }
    o_count[i] = cnt;
}
void launch_debug_trace(const FrameParams& P, const DeviceModel& M, const uint32_t* d_pixels, int64_t n_pix, uint32_t max_samples,
                        float* d_t, uint32_t* d_cell, uint32_t* d_mip, float* d_pos, uint32_t* d_count, float* d_ray, cudaStream_t s) {
    if (n_pix <= 0) return;
    debug_trace_kernel<<<(unsigned)((n_pix + 63) / 64), 64, 0, s>>>(P, M, d_pixels, n_pix, max_samples, d_t, d_cell, d_mip, d_pos, d_count, d_ray);
}

__global__ void debug_mesh_kernel(MeshDevice mesh, FrameParams P, const unsigned long long* __restrict__ zbuf, float* __restrict__ rgba2, float* __restrict__ depth2,
                                  int32_t* __restrict__ tri2, float* __restrict__ surf, float* __restrict__ tsurf, float* __restrict__ lens5) {
    const int ms = P.mesh_scale, W2 = P.width * ms, H2 = P.height * ms;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < W2 * H2 && (rgba2 || depth2 || tri2)) {
        const int x = i % W2, y = i / W2;
        float c[4] = {0.f, 0.f, 0.f, 0.f}; float ht = __uint_as_float(0xFFFFFFFFu); int32_t tri;
        mesh_tap(mesh, P, zbuf, x, y, W2, H2, c, ht, &tri);
        if (rgba2) { rgba2[i * 4] = c[0]; rgba2[i * 4 + 1] = c[1]; rgba2[i * 4 + 2] = c[2]; rgba2[i * 4 + 3] = c[3]; }
        if (depth2) depth2[i] = ht;
        if (tri2) tri2[i] = tri;
    }
    if (i < P.width * P.height && (surf || tsurf || lens5)) {
        float s4[4]; float ts;
        mesh_resolve(mesh, P, zbuf, i % P.width, i / P.width, s4, ts);
        if (surf) { surf[i * 4] = s4[0]; surf[i * 4 + 1] = s4[1]; surf[i * 4 + 2] = s4[2]; surf[i * 4 + 3] = s4[3]; }
        if (tsurf) tsurf[i] = ts;
        if (lens5) {   // same rule as init_one_ray
            LensHit L; L.w = 0.f; L.t = 0.f; L.n = v3(0.f, 0.f, 1.f);
            if (P.lens_on) { lens_resolve(mesh, P, zbuf, i % P.width, i / P.width, L); if (L.w > 0.f && ts != 0.0f && ts < L.t) L.w = 0.f; }
            lens5[i * 5] = L.w; lens5[i * 5 + 1] = L.w > 0.f ? L.t : 0.f; lens5[i * 5 + 2] = L.n.x; lens5[i * 5 + 3] = L.n.y; lens5[i * 5 + 4] = L.n.z;
        }
    }
}
void launch_debug_mesh(const MeshDevice& mesh, const FrameParams& P, const unsigned long long* d_zbuf, float* d_rgba2, float* d_depth2, int32_t* d_tri2,
                       float* d_surf, float* d_tsurf, float* d_lens5, cudaStream_t s) {
    const int n = P.width * P.mesh_scale * P.height * P.mesh_scale;
    debug_mesh_kernel<<<(n + 127) / 128, 128, 0, s>>>(mesh, P, d_zbuf, d_rgba2, d_depth2, d_tri2, d_surf, d_tsurf, d_lens5);
}

}  // namespace nmr
