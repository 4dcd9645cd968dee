// ref_harness.cu - links the REFERENCE's own headers (included from /root/reference where they lie, never
// copied) into oracle/_ref/libnmr_ref.so so that the C oracle can be pinned against them on the CPU.
// TEST INFRASTRUCTURE ONLY; built only when /root/reference is present (see oracle/build_ref.py).
//
// Exposed (host side of the reference's own code):
//   S/floatyremover.h            NgpGrid ctor / cluster / point_set_importance / to_ngp_grid
//   S/orbit_camera.h + R/dependencies/flythrough_camera.h   orbitcam, flythrough_camera_update/look_to
//   S/ngp/random_val.cuh         ld_random_val
//   S/ngp/bounding_box.cuh       BoundingBox::ray_intersect / contains
//   S/ngp/ngp_common.cuh         pixel_to_ray, linear_to_srgb, srgb_to_linear
//   T/include/tiny-cuda-nn/common_device.h   morton3D, morton3D_invert
//   T/include/tiny-cuda-nn/encodings/grid.h  grid_scale, grid_resolution
//   S/ngp/nerf_loader.cuh        NerfDataset::nerf_matrix_to_ngp / ngp_matrix_to_nerf (behind Testbed.crop_box(nerf_space=True))
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ngp/ngp_common.cuh"
#include "ngp/bounding_box.cuh"
#include "ngp/random_val.cuh"
#include <tiny-cuda-nn/common_device.h>
#include <tiny-cuda-nn/encodings/grid.h>

#include "ngp/nerf_loader.cuh"
#include "floatyremover.h"
#define FLYTHROUGH_CAMERA_IMPLEMENTATION
#include "orbit_camera.h"

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int ref_remove_floaties(uint8_t* cells, int64_t* best_size, int64_t* best_score) {
    std::vector<uint8_t> grid(cells, cells + 8 * 128 * 128 * 128);
    NgpGrid g{grid};
    auto clusters = g.cluster();
    if (clusters.empty()) return -1;
    const auto& largest = *std::max_element(clusters.begin(), clusters.end(),
        [](const NgpGrid::DensityPointSet& a, const NgpGrid::DensityPointSet& b) {
            return NgpGrid::point_set_importance(a) < NgpGrid::point_set_importance(b);
        });
    if (best_size) *best_size = (int64_t)largest.size();
    if (best_score) *best_score = NgpGrid::point_set_importance(largest);
    NgpGrid::to_ngp_grid(cells, largest);
    return (int)clusters.size();
}

struct RefCamera { float view[16]; float eye[3]; float look[3]; float pivot[3]; float up[3]; };

REF_API void ref_camera_init(RefCamera* c) {
    c->eye[0] = 0.f; c->eye[1] = 0.f; c->eye[2] = 2.f;
    c->look[0] = 0.f; c->look[1] = -0.000001f; c->look[2] = -0.999999f;
    c->up[0] = 0; c->up[1] = 1; c->up[2] = 0;
    c->pivot[0] = c->pivot[1] = c->pivot[2] = 0.f;
    flythrough_camera_update(c->eye, c->look, c->up, c->view, 0.016f, 0, 0, 90.f, 0, 0, 0, 0, 0, 0, 0, 0, 0);
}
REF_API void ref_camera_orbit(RefCamera* c, float delta_azimuth, float delta_polar, float delta_zoom) {
    // NerfMeshRenderer::orbit (S/nerf_mesh_renderer.cu:896-899)
    orbitcam(c->eye, c->pivot, c->up, c->look, c->view, delta_polar, delta_azimuth, delta_zoom);
}

REF_API float ref_ld_random_val(uint32_t index, uint32_t seed) { return ngp::ld_random_val(index, seed); }
REF_API uint32_t ref_morton3D(uint32_t x, uint32_t y, uint32_t z) { return tcnn::morton3D(x, y, z); }
REF_API uint32_t ref_morton3D_invert(uint32_t x) { return tcnn::morton3D_invert(x); }
REF_API float ref_linear_to_srgb(float x) { return ngp::linear_to_srgb(x); }
REF_API float ref_srgb_to_linear(float x) { return ngp::srgb_to_linear(x); }
REF_API float ref_grid_scale(uint32_t level, float log2_pls, uint32_t base) { return tcnn::grid_scale(level, log2_pls, base); }
REF_API uint32_t ref_grid_resolution(float scale) { return tcnn::grid_resolution(scale); }

REF_API void ref_aabb_ray_intersect(const float* bmin, const float* bmax, const float* pos, const float* dir, float* out2) {
    ngp::BoundingBox b{Eigen::Vector3f{bmin[0], bmin[1], bmin[2]}, Eigen::Vector3f{bmax[0], bmax[1], bmax[2]}};
    Eigen::Vector2f r = b.ray_intersect(Eigen::Vector3f{pos[0], pos[1], pos[2]}, Eigen::Vector3f{dir[0], dir[1], dir[2]});
    out2[0] = r.x(); out2[1] = r.y();
}
REF_API int ref_aabb_contains(const float* bmin, const float* bmax, const float* p) {
    ngp::BoundingBox b{Eigen::Vector3f{bmin[0], bmin[1], bmin[2]}, Eigen::Vector3f{bmax[0], bmax[1], bmax[2]}};
    return b.contains(Eigen::Vector3f{p[0], p[1], p[2]}) ? 1 : 0;
}

// pixel_to_ray + the normalisation / +0.5 shift of init_rays_with_payload_kernel_nerf (S/ngp/testbed.cu:435-446,
// identity model matrix); out = origin3, dir3 (unnormalised), ndir3
REF_API void ref_pixel_to_ray(uint32_t spp, int px, int py, int W, int H, const float* cam12_colmajor, float* out9) {
    Eigen::Matrix<float, 3, 4> cam;
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 3; ++r) cam(r, c) = cam12_colmajor[c * 3 + r];
    ngp::Ray ray = ngp::pixel_to_ray(spp, {px, py}, {W, H}, Eigen::Vector2f{1.f, 1.f}, cam, Eigen::Vector2f{0.5f, 0.5f}, Eigen::Vector3f{0.f, 0.f, 0.f});
    Eigen::Vector3f nd = ray.d.normalized();
    out9[0] = ray.o.x(); out9[1] = ray.o.y(); out9[2] = ray.o.z();
    out9[3] = ray.d.x(); out9[4] = ray.d.y(); out9[5] = ray.d.z();
    out9[6] = nd.x(); out9[7] = nd.y(); out9[8] = nd.z();
}

// ---- the GUI's trajectory tool (S/nerf_mesh_renderer.cu:630-659): where it puts the camera for an angle of its arc, and the text
// it writes to transform_N.  The reference's own flythrough_camera_look_to, glm::normalize and Eigen::IOFormat. ----
#include <glm/glm.hpp>
#include <sstream>
REF_API void ref_trajectory_camera(float angle, float distance, float height, const float* lookat3, float* eye3, float* look3, float* view16) {
    float cam_pos[3], cam_look[3], up[3] = {0.f, 1.f, 0.f};
    cam_pos[0] = cosf(angle) * distance;
    cam_pos[1] = height;
    cam_pos[2] = sinf(angle) * distance;
    glm::vec3 look{lookat3[0] - cam_pos[0], lookat3[1] - cam_pos[1], lookat3[2] - cam_pos[2]};
    look = glm::normalize(look);
    cam_look[0] = look.x; cam_look[1] = look.y; cam_look[2] = look.z;
    flythrough_camera_look_to(cam_pos, cam_look, up, view16, 0);
    for (int k = 0; k < 3; ++k) { eye3[k] = cam_pos[k]; look3[k] = cam_look[k]; }
}
REF_API int ref_format_transform(const float* m12_rowmajor, char* out, int cap) {
    Eigen::Matrix<float, 3, 4> M;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) M(r, c) = m12_rowmajor[r * 4 + c];
    Eigen::IOFormat json_format(Eigen::FullPrecision, 0, ", ", ",\n", "[", "]", "[", "]");
    std::ostringstream ss;
    ss << M.format(json_format);
    const std::string t = ss.str();
    if ((int)t.size() + 1 > cap) return -1;
    std::memcpy(out, t.c_str(), t.size() + 1);
    return (int)t.size();
}

// NerfDataset's coordinate conversions (S/ngp/nerf_loader.cuh:115-153) on a 3x4 matrix given and returned row-major
REF_API void ref_dataset_matrix(int to_ngp, int scale_columns, float scale, const float* offset3, int from_mitsuba, const float* m12_rowmajor, float* out12_rowmajor) {
    ngp::NerfDataset d;
    d.scale = scale; d.offset = Eigen::Vector3f(offset3[0], offset3[1], offset3[2]); d.from_mitsuba = from_mitsuba != 0;
    Eigen::Matrix<float, 3, 4> m;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) m(r, c) = m12_rowmajor[r * 4 + c];
    const Eigen::Matrix<float, 3, 4> o = to_ngp ? d.nerf_matrix_to_ngp(m, scale_columns != 0) : d.ngp_matrix_to_nerf(m, scale_columns != 0);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) out12_rowmajor[r * 4 + c] = o(r, c);
}
