// host.h - host-side model/mesh/camera structures and loaders of libnmr.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace nmr {

constexpr int kMaxLevels = 16;
constexpr uint32_t kGridSize = 128;
constexpr uint32_t kCascades = 8;
constexpr uint32_t kGridCells = kGridSize * kGridSize * kGridSize;
constexpr uint32_t kBitfieldBytes = kCascades * kGridCells / 8;   // 2 MiB

// What Testbed::load_snapshot + reset_network + Trainer::deserialize extract from a snapshot
// (S/ngp/testbed.cu:939-1002, 1098-1115, 1158-1236; T/include/tiny-cuda-nn/trainer.h:285-310).
struct HostModel {
    // hash grid (T/include/tiny-cuda-nn/encodings/grid.h:959-1025)
    int n_levels = 16, n_features_per_level = 2, log2_hashmap_size = 19, base_resolution = 16;
    int hash_type = 1;                       // 0 Prime, 1 CoherentPrime, 2 ReversedPrime
    float per_level_scale = 0.f;
    uint32_t offsets[kMaxLevels + 1] = {};   // in entries (x n_features_per_level halves)
    float scales[kMaxLevels] = {};
    uint32_t resolutions[kMaxLevels] = {};
    uint32_t stride_y[kMaxLevels] = {}, stride_z[kMaxLevels] = {};
    int dense[kMaxLevels] = {};
    // MLPs (FullyFusedMLP, 64 neurons)
    int density_hidden = 1, rgb_hidden = 2;
    // parameters, params_binary order: density net | rgb net | hash grid, as fp16 bit patterns
    std::vector<uint16_t> params;
    size_t mlp_params = 0;                   // number of MLP halves before the grid
    // occupancy
    int aabb_scale = 1, max_cascade = 0;
    std::vector<uint16_t> density_grid;      // fp16, 128^3 * (max_cascade+1), Morton order
    float cone_angle_constant = 0.f;
    // boxes
    float aabb_min[3] = {0, 0, 0}, aabb_max[3] = {1, 1, 1};               // m_aabb
    float render_aabb_min[3] = {0, 0, 0}, render_aabb_max[3] = {1, 1, 1}; // m_render_aabb
    float render_aabb_to_local[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};          // row-major
    int rgb_activation = 2;                  // Logistic unless the dataset is HDR (S/ngp/testbed.cu:1028)
    int density_activation = 3;              // Exponential (S/ngp/testbed.cuh:469)
    int64_t training_step = 0;               // snapshot["training_step"] (Testbed.training_step)
    float loss = 0.f;                        // snapshot["loss"] (Testbed.loss)
    // what the secondary Python properties read (S/ngp/testbed.cu:952, 1100, 1117; S/ngp/json_binding.h:188-197)
    float bounding_radius = 1.f;             // snapshot["bounding_radius"] (Testbed.bounding_radius, translate_camera)
    float dataset_scale = 1.f, dataset_offset[3] = {0, 0, 0}, dataset_up[3] = {0, 1, 0};   // NerfDataset::scale / offset / up
    int from_mitsuba = 0;
};

// Throws std::runtime_error with a description on malformed / unsupported snapshots.
HostModel load_snapshot(const std::string& path);
// 2 MiB occupancy bitfield from the fp16 density grid (S/ngp/testbed.cu:119-166, 1120-1135) - host version used when the
// device path is not wanted (tests of the loader); the product builds the bitfield on the device (occupancy kernels).
void build_level_table(HostModel& m);

struct HostMesh {
    std::vector<float> positions;   // object space, xyz
    std::vector<float> normals;
    std::vector<float> texcoords;
    std::vector<uint32_t> indices;
    float base_color[4] = {1, 1, 1, 1};
    float emissive[3] = {0, 0, 0};
    float metallic = 1.f, roughness = 1.f;
    int tex_w = 0, tex_h = 0;
    std::vector<uint8_t> tex_rgba8;  // base colour texture (sRGB), row-major, may be empty
    // the other textures of the reference's closest-hit program (S/optix/optix_scene.cu:234-258; S/gltf_scene.cpp:160-215):
    // emissive (sRGB), metallicRoughness (linear; G = roughness, B = metallic), normal (linear, tangent space), occlusion (linear, R)
    struct Texture { int w = 0, h = 0; std::vector<uint8_t> rgba8; };
    Texture tex_emissive, tex_metallic_roughness, tex_normal, tex_occlusion;
    float normal_scale = 1.f, occlusion_strength = 1.f;
    // per-vertex tangents (xyz + handedness w): the file's TANGENT attribute, else generated from the UV derivatives
    // (per-triangle tangents accumulated per vertex, Gram-Schmidt against the normal; the reference runs MikkTSpace here,
    // S/gltf_mikktspace_handler.cpp, whose seam splitting and angle weighting are not reproduced)
    std::vector<float> tangents;
    // node 0 TRS as loaded from the file (overwritten by load_mesh's t/s/r, S/nerf_mesh_renderer.cu:952-954)
    float t[3] = {0, 0, 0}, s[3] = {1, 1, 1}, r_wxyz[4] = {1, 0, 0, 0};
    std::string warning;            // e.g. "texture could not be decoded, using constant colour"
    // Lens surfaces (no reference equivalent - SURVEY.md 8f.1): triangles of primitives whose material is transmissive
    // (KHR_materials_transmission.transmissionFactor > 0, or alphaMode BLEND with baseColorFactor alpha < 1).  They are
    // not shaded as opaque surfaces; a ray meeting one splits into a reflected and a transmitted ray.
    std::vector<uint8_t> tri_lens;  // one flag per triangle (indices.size() / 3), all zero when the file has no such material
    bool has_lens = false;
    float lens_ior = 1.5f;          // KHR_materials_ior.ior of the first lens material
    float lens_transmission = 1.f;  // transmissionFactor, or 1 - alpha
    float lens_tint[3] = {1, 1, 1}; // baseColorFactor rgb of the first lens material
};

// glTF 2.0 (.gltf + external/embedded buffers, or .glb): all primitives of all nodes of the default scene are
// concatenated like GltfScene::getMeshPrimitives (S/gltf_scene.h:195-224); material of the first primitive.
HostMesh load_gltf(const std::string& path);
// Tangents (xyz + handedness, 4 floats per vertex) of an indexed triangle list without a TANGENT attribute: Mikkelsen's method as the
// reference runs it per mesh primitive (S/gltf_scene.cpp:150-155, S/gltf_mikktspace_handler.cpp:14-66).  mikk.cpp.  Vertices no
// triangle uses keep what `tangents` held.
void mikk_tangents(const float* positions, const float* normals, const float* texcoords, size_t n_vert,
                   const uint32_t* indices, size_t n_idx, float* tangents);
// 8-bit PNG (grey / grey+alpha / RGB / RGBA / palette, non-interlaced) -> RGBA8.  Throws on anything else.
void decode_png(const uint8_t* data, size_t size, int& w, int& h, std::vector<uint8_t>& rgba);

// world-space vertices / normals for T*R*S (S/gltf_scene.h:122-127): out arrays sized like the inputs
void transform_mesh(const HostMesh& m, const float t[3], const float s[3], const float r_wxyz[4],
                    std::vector<float>& world_pos, std::vector<float>& world_nrm);
// what the normal map's TBN matrix needs (computeTbnMatrix, S/optix/optix_scene.cu:92-98): per vertex M3 n, M3 t (M3 = R S) and
// the tangent's handedness, 8 floats each; and the normal matrix R S^-1 (row-major)
void transform_tangent_frames(const HostMesh& m, const float s[3], const float r_wxyz[4], std::vector<float>& wtbn, float nmat[9]);

// The NerfMeshRenderer camera (S/nerf_mesh_renderer.cuh:88-95; S/orbit_camera.h; flythrough_camera.h)
struct OrbitCamera {
    float view[16];
    float eye[3] = {0.f, 0.f, 2.f};
    float look[3] = {0.f, -0.000001f, -0.999999f};
    float pivot[3] = {0.f, 0.f, 0.f};
    float up[3] = {0.f, 1.f, 0.f};
    OrbitCamera();
    void orbit(float delta_azimuth, float delta_polar, float delta_scroll);
    void look_from(const float eye3[3], const float look3[3]);       // flythrough_camera_look_to(cam_pos, cam_look, up, viewMat, 0)
    void trajectory_pose(float angle, float distance, float height, const float lookat3[3]);   // the GUI's trajectory tool, S/nerf_mesh_renderer.cu:649-658
    void matrix(int screen_w, int screen_h, float out12[12]) const;   // updateModelViewProj
};

}  // namespace nmr
