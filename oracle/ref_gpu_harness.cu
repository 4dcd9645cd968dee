// ref_gpu_harness.cu - TEST INFRASTRUCTURE ONLY.
//
// C entry points over the REFERENCE's own NeRF renderer (ngp::Testbed, S/ngp/testbed.{cuh,cu}, and the vendored
// tiny-cuda-nn it calls), compiled for sm_100 from the sources where they lie by oracle/Makefile.refgpu into
// oracle/_ref/libnmr_refgpu.so.  Nothing of the reference is copied: this file only CALLS its classes and kernels.
// The -m gpu tests use it to pin libnmr's kernels and the C oracle against what the reference itself computes on the
// same B200; bench.py times it as the "reference kernels recompiled for sm_100" baseline.  The product never loads it.
//
// What runs here is exactly the reference's code path below NerfMeshRenderer::render_frame:
//   Testbed::load_snapshot            S/ngp/testbed.cu:939-1002
//   Testbed::render_frame             S/ngp/testbed.cu:1481-1509  (init rays, advance_pos_nerf, NerfTracer::trace loop with
//                                     compact / generate / network / composite, shade, accumulate, tonemap)
//   mesh hand-off                     the harness writes payload.t_surface / payload.surface_color into the tracer's ray
//                                     buffer between frames, which is precisely what copyRaytracingBuffersToNerfRays does
//                                     (S/nerf_mesh_renderer.cu:64-100, 554-558) - the OptiX stage that produces those two
//                                     buffers needs the OptiX SDK and is not built; the caller supplies them.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "testbed.cuh"
#include "nerf_network.cuh"

#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/gpu_memory.h>

#define REF_API extern "C" __attribute__((visibility("default")))

using namespace Eigen;
using namespace ngp;
using namespace tcnn;

// the reference's kernel, defined in S/ngp/testbed.cu:564-633 (external linkage); declared here to launch it directly
NGP_NAMESPACE_BEGIN
__global__ void generate_next_nerf_network_inputs(
        const uint32_t n_elements, BoundingBox render_aabb, Matrix3f render_aabb_to_local, BoundingBox train_aabb, Vector2f focal_length,
        Vector3f camera_fwd, NerfPayload* __restrict__ payloads, PitchedPtr<NerfCoordinate> network_input, uint32_t n_steps,
        const uint8_t* __restrict__ density_grid, uint32_t min_mip, float cone_angle_constant, const float* extra_dims);
NGP_NAMESPACE_END

namespace {

struct RefCtx {
    std::unique_ptr<Testbed> tb;
    std::string err;
    int last_w = 0, last_h = 0;
};
thread_local std::string g_err;

Matrix<float, 3, 4> cam_from12(const float* c) {
    Matrix<float, 3, 4> m;
    for (int col = 0; col < 4; ++col) for (int r = 0; r < 3; ++r) m(r, col) = c[col * 3 + r];
    return m;
}

__global__ void write_surface_kernel(uint32_t n, NerfPayload* payloads, const float4* surf, const float* ts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = surf ? surf[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    payloads[i].surface_color = Vector4f{s.x, s.y, s.z, s.w};
    payloads[i].t_surface = ts ? ts[i] : 0.f;
}

__global__ void read_payload_kernel(uint32_t n, const NerfPayload* payloads, float* out10) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const NerfPayload& p = payloads[i];
    float* o = out10 + (size_t)i * 10;
    o[0] = p.origin.x(); o[1] = p.origin.y(); o[2] = p.origin.z();
    o[3] = p.dir.x(); o[4] = p.dir.y(); o[5] = p.dir.z();
    o[6] = p.t; o[7] = p.alive ? 1.f : 0.f; o[8] = p.t_start; o[9] = (float)p.n_steps;
}

// first call at a resolution: run the tracer's ray set-up once so that the ray buffer exists (the reference's first frame)
void init_rays(RefCtx* r, const Matrix<float, 3, 4>& cam, int w, int h, uint32_t spp_index, CudaRenderBuffer& rb) {
    Testbed& t = *r->tb;
    cudaStream_t stream = t.m_stream.get();
    Vector2f focal_length = t.calc_focal_length(rb.in_resolution(), t.m_fov_axis, t.m_zoom);
    t.m_nerf.tracer.init_rays_from_camera(
        spp_index, t.m_network->padded_output_width(), t.m_nerf_network->n_extra_dims(), rb.in_resolution(), focal_length, cam, cam,
        Vector4f::Zero(), t.render_screen_center(), t.m_parallax_shift, t.m_quilting_dims, t.m_snap_to_pixel_centers, t.m_render_aabb,
        t.m_render_aabb_to_local, Matrix4f::Identity(), t.m_render_near_distance, t.m_slice_plane_z + t.m_scale, t.m_aperture_size, Lens{},
        t.m_envmap.envmap->params_inference(), t.m_envmap.resolution, nullptr, t.m_distortion.resolution, rb.frame_buffer(), rb.depth_buffer(),
        t.m_nerf.density_grid_bitfield.data(), t.m_nerf.show_accel, t.m_nerf.cone_angle_constant, stream);
}

template <typename F>
int guarded(RefCtx* r, F&& f) {
    try { f(); return 0; }
    catch (const std::exception& e) { (r ? r->err : g_err) = e.what(); return -1; }
}

}  // namespace

REF_API void* refgpu_create() {
    try {
        auto* r = new RefCtx;
        r->tb.reset(new Testbed("ref"));
        return r;
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
REF_API void refgpu_destroy(void* h) { delete static_cast<RefCtx*>(h); }
REF_API const char* refgpu_last_error(void* h) { return h ? static_cast<RefCtx*>(h)->err.c_str() : g_err.c_str(); }

// NerfMeshRenderer::loadNerf (S/nerf_mesh_renderer.cu:967-1000) minus the GL render texture
REF_API int refgpu_load_snapshot(void* h, const char* path) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        r->tb->load_snapshot(path);
        r->tb->set_fov(45.f);
        CUDA_CHECK_THROW(cudaDeviceSynchronize());
    });
}

REF_API int refgpu_set_render_aabb(void* h, const float* mn, const float* mx) {
    RefCtx* r = static_cast<RefCtx*>(h);
    r->tb->m_render_aabb = BoundingBox{Vector3f{mn[0], mn[1], mn[2]}, Vector3f{mx[0], mx[1], mx[2]}};
    return 0;
}
REF_API int refgpu_get_render_aabb(void* h, float* mn, float* mx) {
    RefCtx* r = static_cast<RefCtx*>(h);
    for (int k = 0; k < 3; ++k) { mn[k] = r->tb->m_render_aabb.min[k]; mx[k] = r->tb->m_render_aabb.max[k]; }
    return 0;
}
// Testbed::set_crop_box / crop_box / crop_box_corners (S/ngp/testbed.cu:1421-1477), matrices row-major 3x4.  set: what the call leaves
// in m_render_aabb_to_local (3x3 row-major) and m_render_aabb; get: the matrix and its eight corners
REF_API int refgpu_set_crop_box(void* h, const float* m12, int nerf_space, float* r2l9, float* mn3, float* mx3) {
    Testbed& t = *static_cast<RefCtx*>(h)->tb;
    Eigen::Matrix<float, 3, 4> m;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) m(r, c) = m12[r * 4 + c];
    t.set_crop_box(m, nerf_space != 0);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) r2l9[r * 3 + c] = t.m_render_aabb_to_local(r, c);
    for (int k = 0; k < 3; ++k) { mn3[k] = t.m_render_aabb.min[k]; mx3[k] = t.m_render_aabb.max[k]; }
    return 0;
}
REF_API int refgpu_crop_box(void* h, int nerf_space, float* out12, float* corners24) {
    Testbed& t = *static_cast<RefCtx*>(h)->tb;
    const Eigen::Matrix<float, 3, 4> m = t.crop_box(nerf_space != 0);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) out12[r * 4 + c] = m(r, c);
    const auto corners = t.crop_box_corners(nerf_space != 0);
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) corners24[i * 3 + k] = corners[i][k];
    return 0;
}
// The camera helpers of Testbed (S/ngp/testbed.cu:1328-1349) applied to a given camera in this order: set_scale, set_look_at,
// set_view_dir; -> the camera afterwards, look_at() and scale().  m_scale starts at the constructor's value.
REF_API int refgpu_camera_ops(void* h, const float* cam12, float new_scale, const float* look_at3, const float* view_dir3, const float* up3,
                              float* cam_out12, float* look_at_out3, float* scale_before) {
    Testbed& t = *static_cast<RefCtx*>(h)->tb;
    const auto saved = t.m_camera; const float saved_scale = t.scale(); const auto saved_up = t.m_up_dir;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) t.m_camera(r, c) = cam12[r * 4 + c];
    t.m_up_dir = Vector3f{up3[0], up3[1], up3[2]};
    *scale_before = t.scale();
    t.set_scale(new_scale);
    t.set_look_at(Vector3f{look_at3[0], look_at3[1], look_at3[2]});
    t.set_view_dir(Vector3f{view_dir3[0], view_dir3[1], view_dir3[2]});
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) cam_out12[r * 4 + c] = t.m_camera(r, c);
    const Vector3f la = t.look_at();
    for (int k = 0; k < 3; ++k) look_at_out3[k] = la[k];
    t.m_camera = saved; t.set_scale(saved_scale); t.m_camera = saved; t.m_up_dir = saved_up;
    return 0;
}
// Testbed.tonemap_curve (S/python_api.cu:448): ETonemapCurve 0 Identity, 1 ACES, 2 Hable, 3 Reinhard
REF_API int refgpu_set_tonemap_curve(void* h, int curve) {
    static_cast<RefCtx*>(h)->tb->m_tonemap_curve = (ETonemapCurve)curve;
    return 0;
}
// Testbed::m_model_translation / m_model_rotation (S/ngp/testbed.cuh:508-509; the GUI's "Position" / "Rotation" sliders)
REF_API int refgpu_set_model_transform(void* h, const float* translation3, const float* rotation_pi3) {
    Testbed& t = *static_cast<RefCtx*>(h)->tb;
    for (int k = 0; k < 3; ++k) { t.m_model_translation[k] = translation3[k]; t.m_model_rotation[k] = rotation_pi3[k]; }
    return 0;
}
// what NerfMeshRenderer::render_frame copies / z-merges after every NeRF's render_frame (S/nerf_mesh_renderer.cu:582-597): the
// render surface's linear frame buffer [h][w][4] and depth buffer [h][w] of the last refgpu_render call
REF_API int refgpu_get_buffers(void* h, int w, int hh, float* frame, float* depth) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        CudaRenderBuffer& rb = r->tb->m_windowless_render_surface;
        const size_t n = (size_t)w * (size_t)hh;
        if (rb.in_resolution().x() != w || rb.in_resolution().y() != hh) throw std::runtime_error{"render surface has another resolution"};
        if (frame) CUDA_CHECK_THROW(cudaMemcpy(frame, rb.frame_buffer(), n * 16, cudaMemcpyDeviceToHost));
        if (depth) CUDA_CHECK_THROW(cudaMemcpy(depth, rb.depth_buffer(), n * 4, cudaMemcpyDeviceToHost));
    });
}
REF_API int refgpu_set_background(void* h, const float* rgba) {
    static_cast<RefCtx*>(h)->tb->m_background_color = Array4f{rgba[0], rgba[1], rgba[2], rgba[3]};
    return 0;
}

// the 2 MiB occupancy bitfield built by update_density_grid_mean_and_bitfield (S/ngp/testbed.cu:1120-1135)
REF_API int refgpu_get_bitfield(void* h, uint8_t* out) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        auto& b = r->tb->m_nerf.density_grid_bitfield;
        CUDA_CHECK_THROW(cudaMemcpy(out, b.data(), b.size(), cudaMemcpyDeviceToHost));
    });
}
REF_API int64_t refgpu_bitfield_bytes(void* h) { return (int64_t)static_cast<RefCtx*>(h)->tb->m_nerf.density_grid_bitfield.size(); }
REF_API int refgpu_set_bitfield(void* h, const uint8_t* in) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        auto& b = r->tb->m_nerf.density_grid_bitfield;
        CUDA_CHECK_THROW(cudaMemcpy(b.data(), in, b.size(), cudaMemcpyHostToDevice));
    });
}

// NerfNetwork::inference_mixed_precision (S/ngp/nerf_network.cuh:101-135) on caller-supplied samples.
// in: pos[n][3] in [0,1]^3 (already warped), dir01[n][3]; out: n x 16 halves (r, g, b, density, 12 pad) as the reference lays them out.
REF_API int refgpu_network(void* h, const float* pos, const float* dir01, int64_t n, uint16_t* out_n_by_16) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        Testbed& t = *r->tb;
        cudaStream_t stream = t.m_stream.get();
        const uint32_t B = next_multiple((uint32_t)n, tcnn::batch_size_granularity);
        std::vector<float> in((size_t)B * 7, 0.5f);
        for (int64_t i = 0; i < n; ++i) {
            float* c = in.data() + i * 7;
            c[0] = pos[i * 3]; c[1] = pos[i * 3 + 1]; c[2] = pos[i * 3 + 2]; c[3] = 0.f;
            c[4] = dir01[i * 3]; c[5] = dir01[i * 3 + 1]; c[6] = dir01[i * 3 + 2];
        }
        GPUMemory<float> d_in(in.size());
        d_in.copy_from_host(in);
        const uint32_t ow = t.m_nerf_network->padded_output_width();
        GPUMemory<network_precision_t> d_out((size_t)B * ow);
        GPUMatrix<float> positions_matrix(d_in.data(), 7, B);
        GPUMatrix<network_precision_t, RM> rgbsigma_matrix(d_out.data(), ow, B);
        t.m_nerf_network->inference_mixed_precision(stream, positions_matrix, rgbsigma_matrix);
        CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
        std::vector<network_precision_t> host((size_t)B * ow);
        d_out.copy_to_host(host);
        // row-major (ow x B): element (channel c, sample i) at c * B + i
        for (int64_t i = 0; i < n; ++i)
            for (uint32_t c = 0; c < 16 && c < ow; ++c)
                std::memcpy(out_n_by_16 + i * 16 + c, &host[(size_t)c * B + i], 2);
    });
}

// the position encoding alone (kernel_grid, T/include/tiny-cuda-nn/encodings/grid.h:219-349): n x 32 halves
REF_API int refgpu_encode(void* h, const float* pos, int64_t n, uint16_t* out_n_by_32) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        Testbed& t = *r->tb;
        cudaStream_t stream = t.m_stream.get();
        const uint32_t B = next_multiple((uint32_t)n, tcnn::batch_size_granularity);
        std::vector<float> in((size_t)B * 3, 0.5f);
        std::memcpy(in.data(), pos, (size_t)n * 12);
        GPUMemory<float> d_in(in.size());
        d_in.copy_from_host(in);
        auto enc = t.m_nerf_network->encoding();
        const uint32_t ow = enc->padded_output_width();
        GPUMatrixDynamic<float> in_m(d_in.data(), 3, B, CM);
        GPUMatrixDynamic<network_precision_t> out_m(ow, B, stream, enc->preferred_output_layout());
        enc->inference_mixed_precision(stream, in_m, out_m);
        CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
        std::vector<network_precision_t> host((size_t)B * ow);
        CUDA_CHECK_THROW(cudaMemcpy(host.data(), out_m.data(), host.size() * sizeof(network_precision_t), cudaMemcpyDeviceToHost));
        const bool soa = out_m.layout() == RM;   // SoA: element (feature f, sample i) at f * B + i
        for (int64_t i = 0; i < n; ++i)
            for (uint32_t f = 0; f < 32 && f < ow; ++f)
                std::memcpy(out_n_by_32 + i * 32 + f, soa ? &host[(size_t)f * B + i] : &host[(size_t)i * ow + f], 2);
    });
}

// The collision tool's density probes exactly as NerfMeshRenderer::collide drives them (S/nerf_mesh_renderer.cu:1548-1574,
// 1657-1694): payloads memset to zero, alive, one direction, origin = world point + 0.5, written into the tracer's ray buffer;
// mode 0 = NerfTracer::intersects (S/ngp/testbed.cu:1891-1935), mode 1 = NerfTracer::collide (:1814-1888).  The reference
// relies on a previous frame having sized the tracer's buffers; the harness calls its enlarge() for n rays instead.
REF_API int refgpu_probe(void* h, int mode, const float* points_world, const float* dir, int64_t n, float* out) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        Testbed& t = *r->tb;
        cudaStream_t stream = t.m_stream.get();
        t.m_nerf.tracer.enlarge((size_t)n, t.m_network->padded_output_width(), t.m_nerf_network->n_extra_dims(), stream);
        std::vector<NerfPayload> payloads((size_t)n);
        std::memset((void*)payloads.data(), 0, sizeof(NerfPayload) * (size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            payloads[(size_t)i].alive = true;
            payloads[(size_t)i].dir = {dir[0], dir[1], dir[2]};
            payloads[(size_t)i].idx = (uint32_t)i;
            payloads[(size_t)i].origin = {points_world[i * 3] + 0.5f, points_world[i * 3 + 1] + 0.5f, points_world[i * 3 + 2] + 0.5f};
        }
        CUDA_CHECK_THROW(cudaMemcpy(t.m_nerf.tracer.rays_init().payload, payloads.data(), sizeof(NerfPayload) * (size_t)n, cudaMemcpyHostToDevice));
        GPUMemory<float> dist((size_t)n);
        dist.memset(0);
        if (mode == 0)
            t.m_nerf.tracer.intersects((int)n, *t.m_nerf_network, t.m_aabb, t.get_inference_extra_dims(stream), t.m_nerf.density_activation,
                                       t.m_nerf.density_grid_bitfield.data(), dist.data(), stream);
        else
            t.m_nerf.tracer.collide((int)n, *t.m_nerf_network, t.m_render_aabb, t.m_render_aabb_to_local, t.m_aabb, t.m_nerf.cone_angle_constant,
                                    t.m_nerf.density_grid_bitfield.data(), t.get_inference_extra_dims(stream), t.m_nerf.density_activation, dist.data(), stream);
        CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
        std::vector<float> hst((size_t)n);
        dist.copy_to_host(hst);
        std::memcpy(out, hst.data(), sizeof(float) * (size_t)n);
        t.m_nerf.tracer.clear();
    });
}

// Ray set-up + first-hit DDA (init_rays_with_payload_kernel_nerf + advance_pos_nerf, S/ngp/testbed.cu:355-537) for every
// pixel, then generate_next_nerf_network_inputs (S/ngp/testbed.cu:564-633) with n_steps = 1 repeated max_samples times:
// the occupied-sample sequence of each ray with the network out of the loop.
// Optional surf / ts (host, [h][w][4] / [h][w]) are written into the payloads first (mesh hand-off), 0 = none.
// out_ray [w*h][10] = origin3, dir3, t after advance_pos, alive, t_start, 0
// out_pos [w*h][max_samples][3] warped sample positions, out_dt [w*h][max_samples] warped dt, out_t_after same shape = payload.t after the sample
// out_count [w*h]
REF_API int refgpu_trace(void* h, const float* cam12, int w, int hh, uint32_t spp_index, const float* surf, const float* ts, uint32_t max_samples,
                         float* out_ray, float* out_pos, float* out_dt, float* out_t_after, uint32_t* out_count) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        Testbed& t = *r->tb;
        cudaStream_t stream = t.m_stream.get();
        const uint32_t n = (uint32_t)w * (uint32_t)hh;
        CudaRenderBuffer& rb = t.m_windowless_render_surface;
        rb.resize({w, hh});
        rb.reset_accumulation();
        const auto cam = cam_from12(cam12);
        GPUMemory<float4> d_surf; GPUMemory<float> d_ts;
        if (surf) { d_surf.resize(n); CUDA_CHECK_THROW(cudaMemcpy(d_surf.data(), surf, (size_t)n * 16, cudaMemcpyHostToDevice)); }
        if (ts) { d_ts.resize(n); CUDA_CHECK_THROW(cudaMemcpy(d_ts.data(), ts, (size_t)n * 4, cudaMemcpyHostToDevice)); }
        // The reference writes the hand-off into the ray buffer LEFT OVER from the previous frame (freed back to the arena by
        // render_nerf's scope guard, S/ngp/testbed.cu:1526-1528) and relies on the next frame's enlarge() returning the same
        // block.  Same sequence here: trace once, free, write, trace again - and check the block did not move.
        init_rays(r, cam, w, hh, spp_index, rb);
        NerfPayload* stale = t.m_nerf.tracer.rays_init().payload;
        t.m_nerf.tracer.clear();
        linear_kernel(write_surface_kernel, 0, stream, n, stale, surf ? d_surf.data() : nullptr, ts ? d_ts.data() : nullptr);
        init_rays(r, cam, w, hh, spp_index, rb);
        if (t.m_nerf.tracer.rays_init().payload != stale) throw std::runtime_error{"ray buffer moved between frames: hand-off went to stale memory"};
        NerfPayload* pl = t.m_nerf.tracer.rays_init().payload;
        GPUMemory<float> d_ray((size_t)n * 10);
        linear_kernel(read_payload_kernel, 0, stream, n, pl, d_ray.data());
        CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
        std::vector<float> ray((size_t)n * 10);
        d_ray.copy_to_host(ray);
        for (uint32_t i = 0; i < n; ++i) { std::memcpy(out_ray + (size_t)i * 10, ray.data() + (size_t)i * 10, 40); out_count[i] = 0; }

        GPUMemory<float> d_in((size_t)n * 7);
        std::vector<float> in((size_t)n * 7), after((size_t)n * 10);
        std::vector<uint8_t> live(n);
        for (uint32_t i = 0; i < n; ++i) live[i] = ray[(size_t)i * 10 + 7] != 0.f;
        Vector2f focal_length = t.calc_focal_length(rb.in_resolution(), t.m_fov_axis, t.m_zoom);
        for (uint32_t s = 0; s < max_samples; ++s) {
            PitchedPtr<NerfCoordinate> input_data((NerfCoordinate*)d_in.data(), 1, 0, 0);
            linear_kernel(generate_next_nerf_network_inputs, 0, stream, n, t.m_render_aabb, t.m_render_aabb_to_local, t.m_aabb, focal_length,
                          Vector3f{cam.col(2)}, pl, input_data, 1u, (const uint8_t*)t.m_nerf.density_grid_bitfield.data(), 0u, t.m_nerf.cone_angle_constant,
                          (const float*)nullptr);
            linear_kernel(read_payload_kernel, 0, stream, n, pl, d_ray.data());
            CUDA_CHECK_THROW(cudaStreamSynchronize(stream));
            d_in.copy_to_host(in);
            d_ray.copy_to_host(after);
            bool any = false;
            for (uint32_t i = 0; i < n; ++i) {
                if (!live[i]) continue;
                if (after[(size_t)i * 10 + 9] < 1.f) { live[i] = 0; continue; }    // n_steps == 0: the ray produced no sample
                const size_t o = (size_t)i * max_samples + s;
                out_pos[o * 3] = in[(size_t)i * 7]; out_pos[o * 3 + 1] = in[(size_t)i * 7 + 1]; out_pos[o * 3 + 2] = in[(size_t)i * 7 + 2];
                out_dt[o] = in[(size_t)i * 7 + 3];
                out_t_after[o] = after[(size_t)i * 10 + 6];
                out_count[i] = s + 1;
                any = true;
            }
            if (!any) break;
        }
        t.m_nerf.tracer.clear();
    });
}

// Testbed::render_to_cpu (S/python_api.cu:83-111) with the mesh hand-off of NerfMeshRenderer::render_frame
// (S/nerf_mesh_renderer.cu:554-558) in front of it.  surf / ts: host [h][w][4] / [h][w] or NULL.  out: [h][w][4] float.
// ms_out (optional): device time of Testbed::render_frame alone, CUDA events on the testbed's stream, best of `repeat`.
REF_API int refgpu_render(void* h, const float* cam12, int w, int hh, int spp, int linear, const float* surf, const float* ts, float* out,
                          int repeat, float* ms_out) {
    RefCtx* r = static_cast<RefCtx*>(h);
    return guarded(r, [&] {
        Testbed& t = *r->tb;
        cudaStream_t stream = t.m_stream.get();
        const uint32_t n = (uint32_t)w * (uint32_t)hh;
        CudaRenderBuffer& rb = t.m_windowless_render_surface;
        rb.resize({w, hh});
        const auto cam = cam_from12(cam12);
        t.m_camera = cam; t.m_smoothed_camera = cam;
        GPUMemory<float4> d_surf; GPUMemory<float> d_ts;
        if (surf) { d_surf.resize(n); CUDA_CHECK_THROW(cudaMemcpy(d_surf.data(), surf, (size_t)n * 16, cudaMemcpyHostToDevice)); }
        if (ts) { d_ts.resize(n); CUDA_CHECK_THROW(cudaMemcpy(d_ts.data(), ts, (size_t)n * 4, cudaMemcpyHostToDevice)); }
        // a frame at this resolution must have been traced before the hand-off has a ray buffer to write into
        // (NerfMeshRenderer::render_frame's `if (... rays_init().payload)`); its image is discarded
        rb.reset_accumulation();
        t.render_frame(cam, cam, Vector4f::Zero(), rb, !linear);
        cudaEvent_t e0, e1;
        CUDA_CHECK_THROW(cudaEventCreate(&e0)); CUDA_CHECK_THROW(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int it = 0; it < (repeat < 1 ? 1 : repeat); ++it) {
            rb.reset_accumulation();
            for (int s = 0; s < spp; ++s) {
                NerfPayload* pl = t.m_nerf.tracer.rays_init().payload;
                if (!pl) throw std::runtime_error{"no ray buffer after the first frame"};
                linear_kernel(write_surface_kernel, 0, stream, n, pl, surf ? d_surf.data() : nullptr, ts ? d_ts.data() : nullptr);
                CUDA_CHECK_THROW(cudaEventRecord(e0, stream));
                t.render_frame(cam, cam, Vector4f::Zero(), rb, !linear);
                CUDA_CHECK_THROW(cudaEventRecord(e1, stream));
                CUDA_CHECK_THROW(cudaEventSynchronize(e1));
                if (t.m_nerf.tracer.rays_init().payload != pl) throw std::runtime_error{"ray buffer moved between frames: hand-off went to stale memory"};
                float ms = 0.f; CUDA_CHECK_THROW(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (ms_out) *ms_out = best;
        CUDA_CHECK_THROW(cudaMemcpy2DFromArray(out, (size_t)w * 16, rb.surface_provider().array(), 0, 0, (size_t)w * 16, hh, cudaMemcpyDeviceToHost));
    });
}
