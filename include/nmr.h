/*
 * nmr.h - C ABI of libnmr.so, the B200-native hybrid NeRF + mesh renderer.
 *
 * This is the drop-in boundary for the one hot path of arnerak/nerf-glasses' nerf_mesh_renderer:
 * "render a trained iNGP snapshot composited with a glasses mesh".  The reference exposes that path
 * through the pybind11 module `pynmr` (S/python_api.cu:156-623, S = nerf_mesh_renderer/src/); every
 * entry point below names the reference binding / method it replaces.  A thin `pynmr` shim
 * (nerf-glasses_b200/pynmr) reproduces the Python names on top of this ABI, see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns NMR_OK (0) or a negative nmr_status; nothing throws across the ABI;
 *     nmr_last_error() returns a human-readable description of the last failure on that context
 *     (or of the last failed nmr_create when ctx == NULL);
 *   - the caller owns every host buffer it passes; the context owns all device memory and streams;
 *   - a context is bound to one CUDA device and is NOT thread-safe (calls are serialised by an internal
 *     mutex, like the single-threaded reference whose frame() merely drops the GIL);
 *   - matrices are 3x4 float32, COLUMN-major (12 floats: col0 = right*uLen, col1 = up*vLen, col2 = forward,
 *     col3 = eye), i.e. Eigen::Matrix<float,3,4>::data() of NerfMeshRenderer::viewProjectionMat;
 *   - images are float32 RGBA, row 0 = BOTTOM row of the picture, exactly what Testbed.render returns
 *     (S/python_api.cu:83-111);
 *   - there is no CPU fallback: every compute entry point fails with NMR_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef NMR_H
#define NMR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define NMR_API __attribute__((visibility("default")))
#else
#define NMR_API
#endif

typedef struct nmr_ctx nmr_ctx;

typedef enum nmr_status {
    NMR_OK = 0,
    NMR_ERR_INVALID = -1,     /* bad argument / unknown id */
    NMR_ERR_IO = -2,          /* file missing or unreadable */
    NMR_ERR_FORMAT = -3,      /* malformed msgpack / glTF / PNG, or a snapshot this build does not support */
    NMR_ERR_CUDA = -4,        /* CUDA runtime failure, or no usable device */
    NMR_ERR_STATE = -5        /* call needs something that was not loaded yet */
} nmr_status;

/* Counters of the most recent render on this context (device counters read back after the frame). */
typedef struct nmr_stats {
    uint64_t rays;                 /* primary NeRF rays launched (pixels of this shard) */
    uint64_t rays_alive;           /* rays that reached an occupied cell or a mesh surface */
    uint64_t samples;              /* network evaluations (encoding + both MLPs) consumed */
    uint64_t mesh_rays;            /* mesh-stage primary rays (mesh_scale^2 per pixel) */
    uint64_t kernel_launches;      /* kernels launched by libnmr for that render */
    float    gpu_ms;               /* device time of that render (CUDA events on the context's stream) */
    float    march_ms;             /* device time of the fused march/encode/MLP/composite kernel alone */
} nmr_stats;

/* ---- life cycle -------------------------------------------------------------------------------------------- */

/* NerfMeshRenderer(width, height)  [S/python_api.cu:286 -> S/nerf_mesh_renderer.cu:365-452].
 * Fixes the frame() resolution (render_size_factor 1) and the mesh supersampling factor 2; headless (no GLFW
 * window).  device < 0 selects the current CUDA device. */
NMR_API int nmr_create(int width, int height, int device, nmr_ctx** out_ctx);
NMR_API void nmr_destroy(nmr_ctx* ctx);
NMR_API const char* nmr_last_error(const nmr_ctx* ctx);

/* ---- scene ------------------------------------------------------------------------------------------------- */

/* NerfMeshRenderer.load_nerf(path)  [S/python_api.cu:294 -> S/nerf_mesh_renderer.cu:967-1000 ->
 * Testbed::load_snapshot S/ngp/testbed.cu:939-1002].  On success *out_id is the NeRF handle (0, 1, ...). */
NMR_API int nmr_load_nerf(nmr_ctx* ctx, const char* msgpack_path, int* out_id);

/* NerfMeshRenderer.load_mesh(path, t, s, r)  [S/python_api.cu:288-293 -> S/nerf_mesh_renderer.cu:941-965].
 * r_wxyz is the quaternion in (w, x, y, z) order, which is what the reference's Vector4f argument means
 * (glm::quat{r.x(), r.y(), r.z(), r.w()}); the transform T*R*S applies to node 0 of the glTF scene. */
NMR_API int nmr_load_mesh(nmr_ctx* ctx, const char* gltf_path, const float t[3], const float s[3], const float r_wxyz[4], int* out_id);

/* GltfScene.nodes[0].translation / .scale written from Python  [S/python_api.cu:263-282]; re-transforms the mesh. */
NMR_API int nmr_set_mesh_transform(nmr_ctx* ctx, int mesh_id, const float t[3], const float s[3], const float r_wxyz[4]);
NMR_API int nmr_get_mesh_transform(nmr_ctx* ctx, int mesh_id, float t[3], float s[3], float r_wxyz[4]);

/* renderer.envmap(path): called by V/render.py:228 but absent from the reference module; accepted and ignored. */
NMR_API int nmr_set_envmap(nmr_ctx* ctx, const char* path);

/* NerfMeshRenderer.remove_floaties()  [S/python_api.cu:295 -> S/nerf_mesh_renderer.cu:901-917, S/floatyremover.h].
 * Acts on the most recently loaded NeRF.  out_clusters / out_kept_cells are optional. */
NMR_API int nmr_remove_floaties(nmr_ctx* ctx, int* out_clusters, int64_t* out_kept_cells);

/* Testbed.render_aabb (BoundingBox with writable min/max)  [S/python_api.cu:410, 242-261]. */
NMR_API int nmr_get_render_aabb(nmr_ctx* ctx, int nerf_id, float mn[3], float mx[3]);
NMR_API int nmr_set_render_aabb(nmr_ctx* ctx, int nerf_id, const float mn[3], const float mx[3]);
/* Testbed.aabb (read-only training box). */
NMR_API int nmr_get_aabb(nmr_ctx* ctx, int nerf_id, float mn[3], float mx[3]);
/* Testbed.background_color  [S/python_api.cu:402]. */
NMR_API int nmr_get_background(nmr_ctx* ctx, int nerf_id, float rgba[4]);
NMR_API int nmr_set_background(nmr_ctx* ctx, int nerf_id, const float rgba[4]);
/* Testbed.nerf.render_min_transmittance  [S/python_api.cu:470-496]. */
NMR_API int nmr_set_min_transmittance(nmr_ctx* ctx, int nerf_id, float v);

/* ---- camera ------------------------------------------------------------------------------------------------ */

/* NerfMeshRenderer.orbit(delta_azimuth, delta_polar, delta_zoom)  [S/python_api.cu:296 -> S/nerf_mesh_renderer.cu:896-899,
 * S/orbit_camera.h:7-77]. */
NMR_API int nmr_orbit(nmr_ctx* ctx, float delta_azimuth, float delta_polar, float delta_zoom);
/* NerfMeshRenderer.view_projection_mat getter / setter  [S/python_api.cu:297].  The setter also moves the eye used by the
 * mesh stage (the reference leaves cam_pos stale). */
NMR_API int nmr_get_camera(nmr_ctx* ctx, float out12[12]);
NMR_API int nmr_set_camera(nmr_ctx* ctx, const float in12[12]);

/* ---- rendering --------------------------------------------------------------------------------------------- */

/* NerfMeshRenderer.frame()  [S/python_api.cu:287 -> S/nerf_mesh_renderer.cu:499-598]: one hybrid frame (mesh stage at 2x,
 * hand-off, NeRF march, accumulate, tonemap to sRGB) of the first NeRF and all meshes at the constructor resolution.
 * The image stays in device memory; *keep_running is always 1 (there is no window to close). */
NMR_API int nmr_frame(nmr_ctx* ctx, int* keep_running);
/* Copies the image of the last nmr_frame() to host memory (width*height*4 floats).  Headless replacement for the GL blit. */
NMR_API int nmr_read_frame(nmr_ctx* ctx, float* out_rgba);

/* Testbed.render(width, height, spp, linear) -> float32[H, W, 4]  [S/python_api.cu:326-331 -> 83-111].
 * Renders the NeRF `nerf_id` with the current camera; loaded meshes are composited (the reference only shows them
 * through stale payload memory; here the hybrid image is coherent).  Includes the device->host copy. */
NMR_API int nmr_render(nmr_ctx* ctx, int nerf_id, int width, int height, int spp, int linear, float* out_rgba);

/* New: n_views cameras rendered back to back without host synchronisation in between.
 * cams12: n_views x 12 floats; out_rgba: n_views x height x width x 4 floats. */
NMR_API int nmr_render_views(nmr_ctx* ctx, int nerf_id, int n_views, const float* cams12, int width, int height, int linear, float* out_rgba);

/* Data-parallel sharding (no reference equivalent; the reference is single-GPU).  Rows are dealt to ranks in bands of
 * `band` rows: row y belongs to rank (y / band) % world.  A sharded context renders only its own rows; the other rows of
 * the output image are left untouched.  world = 1 restores full-frame rendering. */
NMR_API int nmr_set_shard(nmr_ctx* ctx, int rank, int world, int band);

/* Device pointer of the float4 image of the last render (valid until the next render on this context) - lets a host
 * language gather shards with NCCL without staging through the CPU. */
NMR_API int nmr_get_device_image(nmr_ctx* ctx, void** out_dev_ptr, int* out_width, int* out_height);

/* nmr_frame() without the trailing synchronisation (pipelined callers / throughput measurement); wait with
 * nmr_synchronize(), nmr_get_stats() or nmr_read_frame(). */
NMR_API int nmr_frame_async(nmr_ctx* ctx);

/* Page-locked host buffers for image outputs: device->host copies into these run at full PCIe rate and asynchronously. */
NMR_API void* nmr_host_alloc(size_t bytes);
NMR_API void nmr_host_free(void* p);

/* Copies the float4 image of the last render into caller-owned DEVICE memory (cudaMemcpyAsync device->device on the
 * context's stream, then a stream synchronise): hand-off point for NCCL gathers done by the host language. */
NMR_API int nmr_copy_device_image(nmr_ctx* ctx, void* dst_device_ptr);

/* Measurement helpers (bench.py): evict the L2 between timed steps by overwriting a 256 MiB scratch buffer on the
 * context's stream; not part of any render. */
NMR_API int nmr_flush_l2(nmr_ctx* ctx);

NMR_API int nmr_get_stats(nmr_ctx* ctx, nmr_stats* out);
NMR_API int nmr_synchronize(nmr_ctx* ctx);

/* ---- occupancy grid ---------------------------------------------------------------------------------------- */
/* 8 cascades x 128^3 bits = 2 MiB, Morton order within a cascade (S/ngp/testbed.cu:115-117, 234-264). */
NMR_API int nmr_get_density_bitfield(nmr_ctx* ctx, int nerf_id, uint8_t* out_2mib);
NMR_API int nmr_set_density_bitfield(nmr_ctx* ctx, int nerf_id, const uint8_t* in_2mib);

/* ---- parity probes (run the SAME device functions as the fused render kernel on caller-supplied inputs) ----- */

/* Hash-grid encoding of n positions in [0,1]^3 -> n x 32 fp16 bit patterns (T/.../encodings/grid.h:219-349). */
NMR_API int nmr_debug_encode(nmr_ctx* ctx, int nerf_id, const float* pos_xyz, int64_t n, uint16_t* out_n_by_32);
/* Full network: positions + directions in [0,1]^3 -> n x 4 fp16 (r, g, b raw, density raw)
 * (S/ngp/nerf_network.cuh:101-135). */
NMR_API int nmr_debug_network(nmr_ctx* ctx, int nerf_id, const float* pos_xyz, const float* dir01_xyz, int64_t n, uint16_t* out_n_by_4);
/* Ray set-up + first-hit DDA + the first max_samples occupied samples of the given pixels, network and mesh ignored
 * (S/ngp/testbed.cu:355-537, 564-633).  Outputs: t/cell/mip [n_pix][max_samples], pos [n_pix][max_samples][3],
 * count [n_pix], ray [n_pix][8] = origin3, dir3, t_first, alive. */
NMR_API int nmr_debug_trace(nmr_ctx* ctx, int nerf_id, int width, int height, const uint32_t* pixels, int64_t n_pix, uint32_t max_samples,
                            float* out_t, uint32_t* out_cell, uint32_t* out_mip, float* out_pos, uint32_t* out_count, float* out_ray);
/* Mesh stage alone at (mesh_scale*width) x (mesh_scale*height): RGBA, hitT (NaN bit pattern on miss), triangle id (-1 miss)
 * (S/optix/optix_scene.cu:120-325), followed by the 2x2 resolve into surface colour / t_surface
 * (S/nerf_mesh_renderer.cu:64-100).  Any output pointer may be NULL. */
NMR_API int nmr_debug_mesh(nmr_ctx* ctx, int width, int height, float* out_rgba2, float* out_depth2, int32_t* out_tri2,
                           float* out_surf_rgba, float* out_t_surface);
/* Per-ray network-evaluation counts and the linear premultiplied frame buffer of the last render (before tonemap). */
NMR_API int nmr_debug_last_frame(nmr_ctx* ctx, float* out_frame_rgba, float* out_depth, uint32_t* out_n_samples);
/* Bring-up knobs: bit 0 = evaluate the MLPs on CUDA cores instead of tcgen05 (also NMR_MLP=scalar), bit 1 = swap the
 * UMMA shared-memory descriptor offsets (also NMR_UMMA_SWAP=1). */
NMR_API int nmr_debug_set_flags(nmr_ctx* ctx, uint32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* NMR_H */
