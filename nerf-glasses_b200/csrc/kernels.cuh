// kernels.cuh - launch interface between the C-ABI layer (api.cpp) and the CUDA kernels (kernels.cu, floaties.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_common.cuh"

namespace nmr {

// per-frame output surfaces (all device pointers; optional ones may be null)
struct FrameOut {
    void* image;          // displayed image (tonemapped), W*H pixels of FrameParams::out_format (float4 / half4 / uchar4)
    float4* accum;        // running mean over spp (linear, premultiplied), W*H
    float4* frame;        // this sample's linear premultiplied frame buffer (parity probe), W*H or null
    float* depth;         // W*H or null
    uint32_t* n_samples;  // network evaluations per ray, W*H or null
    float4* lens;         // lens hand-off per pixel, 2 x W*H: (normal.xyz, t_lens) (coverage, -, -, -); null when the frame has no lens
    float* lens_scratch;  // per ray group of the march kernel: state parked across the segments of a lens ray (kLensStash floats each)
    // measurement aid (NMR_PHASE_LOG): per warpgroup of the march kernel and tile iteration, clock64 at the iteration's start, after
    // batch generation, after the encoding, after the network, after compositing, and globaltimer (ns) at the start; the warpgroup's last entry is its exit from the loop (start stamps only): [n_warpgroups][kPhaseIters][kPhaseWords]; null normally
    unsigned long long* phase_log;
};
constexpr int kPhaseIters = 32, kPhaseWords = 6;
// The reference's n_steps schedule for frames with MORE than 1/8 live pixels (SurfaceMode auto): pass 1 marches every ray and
// histograms the sample index at which each ray dies; pass 2 replays clamp(pixels / live rays, 1, 8) per wavefront iteration
// over that histogram and re-marches the rays that carry a mesh surface with those batch sizes.
struct SchedArgs {
    uint32_t* hist;             // [kSchedBins] death histogram (pass 1 adds, pass 2 reads); null: no schedule machinery
    const uint32_t* surf_list;  // queue slots of the rays that carry a mesh surface (pass 2 consumes these)
    int pass;                   // 0: plain single pass, 1: first pass of a possible two, 2: surface-ray pass
};
constexpr uint32_t kSchedBins = 2048, kSchedMax = 2048;
constexpr uint32_t kSurfRayFlag = 0x40000000u;   // bit 30 of the pixel index in a ray record: the ray carries a mesh surface
constexpr uint32_t kLensRayFlag = 0x80000000u;   // bit 31 of the pixel index in a ray record: the pixel has a lens hand-off entry
constexpr int kLensStash = 24;

// device counters of one render: [0] rays queued by the init kernel, [1] queue cursor of the march kernel,
// [2..3] total network evaluations (64-bit), [4] ray batches, [5] batch generation passes, [6] cursor of the surface-ray pass,
// [7] rays that carry a mesh surface (length of the surface list), [8] CTAs of the set-up kernel that have finished (the march
// kernel of an overlapped frame learns from it that the queue is complete).  These nine are zeroed at the head of a frame.
// [9] queue records the PREVIOUS frame wrote (its march kernel stores it; the next frame's clear kernel resets their ready words),
// [10..11] / [12..13] globaltimer of the first march CTA's start / the last one's end (64-bit min / max) of this frame,
// [14] lens rays among [0] (zeroed with the frame counters).
constexpr int kNumCounters = 16, kFrameCounters = 9;
constexpr int kCntInitDone = 8, kCntPrevCount = 9, kCntMarchStart = 10, kCntMarchEnd = 12, kCntLensRays = 14;
constexpr int kRayRecordFloat4s = 3;   // queue record: (dir.xyz, t) (t_start, t_surface, idx, t_limit) (surface rgba)
// The idx word of a record is its READY word: the set-up kernel stores it last, with release semantics, and the march kernel
// reads it first, with acquire semantics - so the march kernel can consume the queue while the set-up kernel is still appending to
// it (overlapped frames).  Records that are not (yet) written hold kEmptyRecord there: the queue is filled with it when it is
// allocated, and every frame's clear kernel resets the records of the frame before.
constexpr uint32_t kQueueSlack = 32768;   // records allocated past width x height: >= ray groups of a march launch (SMs x CTAs per SM x 32), each of which may hold one slot past the last record in an overlapped frame
constexpr uint32_t kEmptyRecord = 0xFFFFFFFFu;

enum DebugFlags : uint32_t {
    kDebugScalarMlp = 1u,     // run the CUDA-core MLP instead of tcgen05 (NMR_MLP=scalar)
    kDebugSwapLboSbo = 2u,    // swap the UMMA descriptor offsets (bring-up aid)
    kDebugKeepProbes = 4u,
    kDebugNoSharedEncode = 8u, // every lane encodes its own sample even when the warp holds few (A/B of the shared encoding)
    kDebugNoBricks = 16u,      // gather every level from the plain hash-table layout (A/B of the brick layout of the coarse levels)
};

void launch_occupancy_build(const uint16_t* d_density_grid_fp16, int n_cascades_present, uint8_t* d_bitfield, float* d_scratch, cudaStream_t s);
// per cascade: min xyz / max xyz (inclusive, in cells) of the set cells; min = 1 << 20 and max = -1 when the cascade is empty
void launch_occupancy_bounds(const uint8_t* d_bitfield, int* d_out48, cudaStream_t s);
// d_occ_scratch: kCoarseRes^3 bytes; d_near_bits: kCoarseRes^2 words (DeviceModel::coarse)
void launch_coarse_build(const uint8_t* d_bitfield, uint8_t* d_occ_scratch, uint32_t* d_near_bits, cudaStream_t s);
// one level of the hash grid re-laid out as 2x2x2 bricks (res^3 cells x 32 bytes), see DeviceModel::brick
void launch_brick_build(const DeviceModel& M, int level, uint32_t res, void* d_out, cudaStream_t s);
// d_out: 10 240 halves; the MLP weights re-laid out as the tcgen05 B operands the march kernel keeps in shared memory
void launch_weights_canonical(const uint16_t* d_mlp, uint16_t* d_out, cudaStream_t s);
void launch_mesh_raster(const MeshDevice& mesh, const FrameParams& P, int rows_owned, unsigned long long* d_zbuf, cudaStream_t s, bool clear = true);
// counters + (optional) schedule histogram + (optional) mesh visibility window in one launch; zbuf_window_words: words of the window (both layers)
size_t zbuf_window_words(const MeshDevice& mesh, const FrameParams& P);
void launch_frame_clear(uint32_t* d_counters, uint32_t* d_hist, unsigned long long* d_zbuf, size_t zbuf_words, float4* d_queue, cudaStream_t s);
// sequence flags of a shared frame target (nmr_gather_*): one word per rank + "consumed" + "error", behind the image
constexpr int kGatherMaxRanks = 32, kGatherConsumed = 32, kGatherError = 33, kGatherFlagWords = 64;
// destination rank of a shared frame target: constant background of every pixel outside both screen rectangles, all rows
void launch_fill_background(const FrameParams& P, void* d_image, cudaStream_t s);
// bytes of one pixel of an image in FrameParams::out_format
inline size_t pixel_bytes(int out_format) { return out_format == kPixelU8 ? 4 : (out_format == kPixelF16 ? 8 : 16); }
void launch_gather_signal(uint32_t* d_flag, uint32_t seq, cudaStream_t s);
void launch_gather_wait(uint32_t* d_flags, int first, int count, uint32_t seq, uint32_t* d_err, cudaStream_t s);
// The set-up kernel only covers the TILE BOX: the 16 x 8 pixel tiles [x0, x0 + nx) x [y0, y0 + ny) of local (owned) rows that the
// union of the two screen rectangles touches; everything else is background, written by the background kernel.  nx == 0: no
// rectangle in view, no set-up kernel at all.  (rot_x, rot_y): first CTA column / row over the occupied cells, relative to the box.
struct TileBox { int x0, y0, nx, ny; unsigned rot_x, rot_y; };
TileBox compute_tile_box(const FrameParams& P, int rows_owned);
// every owned pixel outside the tile box (+ the L2 prefetch of the hash table when prefetch_table); independent of the mesh stage
// and of the set-up kernel, so a frame runs it on a side stream
void launch_background(const FrameParams& P, const DeviceModel& M, const FrameOut& out, int rows_owned, const TileBox& box, bool prefetch_table, int num_sms, cudaStream_t s);
// ray set-up of the tile box.  Returns the number of CTAs launched (what an overlapped march launch has to wait for; 0: not launched).
int launch_init_rays(const FrameParams& P, const DeviceModel& M, const MeshDevice& mesh, const unsigned long long* d_zbuf, int rows_owned,
                     float4* d_queue, uint32_t* d_counters, const FrameOut& out, const TileBox& box, cudaStream_t s, uint32_t* d_surf_list = nullptr);
// n_pixels: pixels traced by this context in this pass (the reference's m_n_rays_initialized), for SurfaceMode auto
void launch_march(const FrameParams& P, const DeviceModel& M, const float4* d_queue, uint32_t* d_counters, const FrameOut& out,
                  uint32_t n_pixels, uint32_t debug_flags, int num_sms, cudaStream_t s, const uint32_t* d_range_end = nullptr, uint32_t* d_cursor = nullptr,
                  const SchedArgs* sched = nullptr, int ctas_per_sm = 0, int overlap_init_ctas = -1);
// (overlap_init_ctas >= 0: OVERLAPPED frame - the kernel is launched with programmatic stream serialisation right behind the
//  set-up kernel of `overlap_init_ctas` CTAs and consumes the whole queue while that kernel is still filling it; d_range_end is
//  ignored.  The caller has resolved the mesh-surface rule on the host: P.surface_mode is not `auto` when a mesh is in view.)

// (d_range_end / d_cursor: consume only the queue records [*d_cursor, *d_range_end) - default: the whole queue)
// density probes of the collision tool: mode 0 = NerfTracer::intersects (alpha at points), 1 = NerfTracer::collide (distance along dir)
void launch_probe(const FrameParams& P, const DeviceModel& M, const float* d_points_world, const float dir[3], int64_t n, int mode, float* d_out,
                  uint32_t debug_flags, int num_sms, cudaStream_t s);
// several NeRFs in one frame: z-merge of per-NeRF linear frame / depth buffers (first = plain copy), then accumulate + tonemap
void launch_combine_buffers(const float* d_in_depth, const float4* d_in_frame, float* d_out_depth, float4* d_out_frame, uint32_t n, bool first, cudaStream_t s);
void launch_present(const FrameParams& P, const float4* d_frame, float4* d_accum, void* d_image, uint32_t n, cudaStream_t s);
// measurement helper (bench.py): L2 -> SM throughput probe, see kernels.cu
void launch_l2_probe(const void* d_buf, uint32_t n_vec, uint32_t loads_per_thread, int mode, uint32_t* d_sink, int num_sms, cudaStream_t s);
// parity probes
void launch_debug_encode(const DeviceModel& M, const float* d_pos, int64_t n, uint16_t* d_out, cudaStream_t s);
void launch_debug_network(const DeviceModel& M, const float* d_pos, const float* d_dir, int64_t n, uint16_t* d_out4, uint32_t debug_flags, cudaStream_t s);
void launch_debug_trace(const FrameParams& P, const DeviceModel& M, const uint32_t* d_pixels, int64_t n_pix, uint32_t max_samples,
                        float* d_t, uint32_t* d_cell, uint32_t* d_mip, float* d_pos, uint32_t* d_count, float* d_ray, cudaStream_t s);
void launch_debug_mesh(const MeshDevice& mesh, const FrameParams& P, const unsigned long long* d_zbuf, float* d_rgba2, float* d_depth2, int32_t* d_tri2,
                       float* d_surf, float* d_tsurf, float* d_lens5, cudaStream_t s);
// floatie pruning on the 2 MiB bitfield (floaties.cu); results[0] = clusters, results[1] = kept cells (host-visible after sync)
void launch_remove_floaties(uint8_t* d_bitfield, int max_cascade, uint32_t* d_labels, unsigned long long* d_scratch, cudaStream_t s);
size_t floaties_label_bytes();
size_t floaties_scratch_bytes();

int rows_owned_by(int height, int rank, int world, int band);

}  // namespace nmr
