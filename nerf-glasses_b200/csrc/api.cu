// api.cu - the C ABI of libnmr.so (include/nmr.h): context, scene management and frame orchestration.
// Host-side mirror of NerfMeshRenderer / Testbed as far as the render path needs them
// (S/nerf_mesh_renderer.cu:365-452, 499-598, 896-1000; S/ngp/testbed.cu:939-1135, 1481-1612; S/python_api.cu:83-111).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nmr.h"
#include "host.h"
#include "kernels.cuh"

using namespace nmr;

namespace {

std::string g_create_error;

struct CudaError : std::runtime_error { explicit CudaError(const std::string& m) : std::runtime_error(m) {} };
#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) throw CudaError(std::string(#call) + ": " + cudaGetErrorString(e_));                \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;              // owns device memory
    DevBuf& operator=(const DevBuf&) = delete;
    bool ensure(size_t count) {      // true when the buffer was (re)allocated: its contents are undefined
        if (count <= n) return false;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        CK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

struct Nerf {
    HostModel host;                 // parameters are dropped after upload
    DevBuf<uint16_t> d_params;
    DevBuf<uint8_t> d_bitfield;
    DevBuf<uint16_t> d_mlp_tc;                      // MLP weights in the tensor-core operand layout
    DevBuf<uint32_t> d_coarse;                      // coarse "near" bits of cascade 0 (kernels.cuh: launch_coarse_build)
    DevBuf<uint4> d_bricks;                         // coarse levels of the hash grid as 2x2x2 bricks (DeviceModel::brick)
    DeviceModel dev{};
    float render_aabb_min[3], render_aabb_max[3];
    float occ_min[3], occ_max[3];                   // box around every occupied cell (see update_occupied_box)
    uint64_t n_params = 0;
    float background[4] = {1.f, 1.f, 1.f, 1.f};     // S/ngp/testbed.cuh:525
    float min_transmittance = 0.01f;                // S/ngp/testbed.cuh:484
    int tonemap_curve = 0;                          // Testbed.tonemap_curve (ETonemapCurve), Identity by default
    // Testbed::m_model_translation / m_model_rotation (S/ngp/testbed.cuh:508-509): translation in world units, rotation as three
    // angles in units of pi about X, Y, Z; model_rot = AngleAxis(rx pi, X) * AngleAxis(ry pi, Y) * AngleAxis(rz pi, Z), row-major
    float model_translation[3] = {0.f, 0.f, 0.f}, model_rotation_pi[3] = {0.f, 0.f, 0.f};
    float model_rot[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
};

struct Mesh {
    HostMesh host;
    float t[3], s[3], r[4];
};


struct Surfaces {
    int w = 0, h = 0;
    DevBuf<float4> image, accum, frame;
    DevBuf<float4> image_alt;        // second image buffer of nmr_render_views (a view renders while the previous one is copied out)
    DevBuf<float4> bg_image;         // nmr_render: an image of nothing but the background colour (rows known to be background leave from here at once)
    float bg_value[4] = {-1.f, -1.f, -1.f, -1.f};   // what bg_image holds
    size_t bg_pixels = 0;
    int bg_format = -1;
    int image_format = 0;            // PixelFormat of what `image` holds
    bool image_is_frame = false;     // `image` holds the last frame() (constructor resolution, float32), not a render() / render_views() result
    DevBuf<float> depth;
    DevBuf<uint32_t> n_samples;
    DevBuf<float4> queue;
    DevBuf<unsigned long long> zbuf;       // opaque window, then (scenes with lens surfaces) the lens window
    DevBuf<float4> comb_frame, comb_accum; // several NeRFs in one frame: merged linear frame and its running mean
    DevBuf<float> comb_depth;
    bool comb_valid = false;               // comb_frame / comb_depth hold the last frame()'s merged buffers
    DevBuf<float4> lens;                   // FrameOut::lens
    DevBuf<float> lens_scratch;            // FrameOut::lens_scratch
    // the reference's n_steps schedule for close-ups (SchedArgs): death histogram, batch boundaries + their count, surface rays
    DevBuf<uint32_t> hist, surf_list;
    uint32_t spp = 0;
    void resize(int W, int H, int mesh_scale) {
        const size_t n = (size_t)W * H;
        if (W != w || H != h) spp = 0;
        image.ensure(n); accum.ensure(n); frame.ensure(n); depth.ensure(n); n_samples.ensure(n);
        // a fresh ray queue holds kEmptyRecord (all ones) in every ready word, see kernels.cuh
        // (+ kQueueSlack records: in an overlapped frame every ray group of the march kernel may hold one slot past the last record)
        if (queue.ensure((n + kQueueSlack) * kRayRecordFloat4s)) { CK(cudaMemset(queue.p, 0xFF, queue.n * sizeof(float4))); CK(cudaDeviceSynchronize()); }
        zbuf.ensure(n * (size_t)mesh_scale * mesh_scale * 2);
        w = W; h = H;
    }
};

}  // namespace

struct nmr_ctx {
    std::mutex mu;
    std::string last_error;
    int device = 0, num_sms = 148;
    int width = 0, height = 0;
    int mesh_scale = 2;                                   // mesh_render_size_factor, S/nerf_mesh_renderer.cuh:112
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;                   // device->host copies of nmr_render_views
    cudaStream_t aux_stream = nullptr;                    // a frame's background kernel runs here, next to the mesh stage and the ray set-up
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_view[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // per image buffer: rendered, copied
    // nmr_render_views: views are independent frames, and a single frame leaves most of the GPU idle (both its kernels wait on
    // dependent chains).  Helper contexts ("lanes": own stream, ray queue, counters, visibility buffer, image) render several
    // views at once; they borrow the parent's model and mesh buffers.
    std::vector<nmr_ctx*> lanes;
    int march_ctas = 0;                                   // CTAs per SM of this context's march kernel (0 = default); lanes share the SMs
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_tl[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // NMR_TIMELINE=1 (measurement aid): events between a frame's kernels
    bool tl_pending = false;
    OrbitCamera camera;
    float cam12[12];
    float light[3] = {1.f, 1.f, 1.f};                     // S/nerf_mesh_renderer.cuh:95
    std::vector<std::unique_ptr<Nerf>> nerfs;
    std::vector<std::unique_ptr<Mesh>> meshes;
    // concatenated world-space mesh on the device
    bool mesh_dirty = true;
    DevBuf<float> d_wpos, d_wnrm, d_uv, d_tex, d_wtbn, d_xtex[4];
    DevBuf<uint32_t> d_idx;
    DevBuf<uint8_t> d_tri_lens;
    bool scene_has_lens = false;                          // some loaded mesh has a transmissive material
    int lens_enabled = 1;                                 // nmr_set_lens
    float lens_ior = 1.5f, lens_transmission = 1.f, lens_tint[3] = {1.f, 1.f, 1.f};
    bool lens_params_set = false;                         // nmr_set_lens gave explicit parameters
    int lens_model = 0; float lens_thickness = 0.f;      // nmr_set_lens_model
    MeshDevice mesh_dev{};
    float mesh_wmin[3] = {0.f, 0.f, 0.f}, mesh_wmax[3] = {0.f, 0.f, 0.f};   // world-space box of the concatenated mesh
    Surfaces surf;
    // shared frame target of tile-sharded rendering (nmr_gather_*): every rank's frame() writes its rows straight into ONE
    // image in the destination rank's memory - its own kernels' stores travel over NVLink, there is no gather afterwards
    DevBuf<float4> gather_image;                          // destination rank: the shared image (exported through CUDA IPC)
    float4* frame_target = nullptr;                       // where frame() writes pixels: gather_image / a peer's, or null = own image
    void* ipc_mapped = nullptr;                           // peer image opened with cudaIpcOpenMemHandle
    uint32_t* gather_flags = nullptr;                     // sequence flags behind the shared image (kernels.cuh: kGather*)
    uint32_t gather_seq = 0;                              // frames rendered into the shared image so far (same on every rank)
    bool gather_is_dst = false;
    DevBuf<uint32_t> d_counters;
    cudaEvent_t ev_rows = nullptr;                        // nmr_render: the frame is complete, its rows may be copied out
    uint32_t* h_counters = nullptr;                       // pinned
    DevBuf<float> d_scratch;
    int shard_rank = 0, shard_world = 1, shard_band = 8;
    int surface_mode = 0;                                 // nmr_surface_mode
    int overlap = 1;                                      // nmr_set_overlap: march kernel consumes the ray queue while the set-up kernel fills it
    bool last_overlapped = false;                         // the last timed pass ran overlapped (march time comes from device timestamps)
    uint32_t debug_flags = 0;
    nmr_stats stats{};
    bool stats_pending = false;
    std::string envmap_path;
    const void* l2_window_ptr = nullptr;                  // NMR_L2_PERSIST: what the stream's access-policy window covers
    // nmr_render_update: host images this context has filled and may update in place - the pixels outside `rect` hold the background
    struct HostImage { void* p; int w, h, fmt; float bg[4]; int rect[4]; };
    std::vector<HostImage> host_images;
    DevBuf<uint8_t> d_flush;
    DevBuf<unsigned long long> d_phase_log;               // NMR_PHASE_LOG (measurement aid)
};

namespace {

int fail(nmr_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->last_error = msg; else g_create_error = msg;
    return code;
}

template <typename F>
int guarded(nmr_ctx* ctx, F&& f) {
    if (!ctx) return fail(nullptr, NMR_ERR_INVALID, "null context");
    std::lock_guard<std::mutex> lock(ctx->mu);
    try {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess) return fail(ctx, NMR_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
        return f();
    } catch (const CudaError& e) {
        return fail(ctx, NMR_ERR_CUDA, e.what());
    } catch (const std::bad_alloc&) {
        return fail(ctx, NMR_ERR_INVALID, "out of host memory");
    } catch (const std::exception& e) {
        const std::string m = e.what();
        const int code = m.rfind("cannot open", 0) == 0 ? NMR_ERR_IO : NMR_ERR_FORMAT;
        return fail(ctx, code, m);
    }
}

// the model as the kernels see it under the context's debug flags
DeviceModel model_for(const nmr_ctx* ctx, const Nerf& n) {
    DeviceModel d = n.dev;
    if (ctx->debug_flags & kDebugNoBricks) d.n_brick = 0;
    return d;
}

Nerf* get_nerf(nmr_ctx* ctx, int id) {
    if (id < 0 || id >= (int)ctx->nerfs.size()) throw std::invalid_argument("unknown NeRF id");
    return ctx->nerfs[(size_t)id].get();
}

void invert3(const float* cam12, float out[9]) {
    // inverse of the matrix whose columns are U, V, W
    const double a = cam12[0], b = cam12[3], c = cam12[6], d = cam12[1], e = cam12[4], f = cam12[7], g = cam12[2], h = cam12[5], i = cam12[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double inv = det != 0.0 ? 1.0 / det : 0.0;
    out[0] = (float)((e * i - f * h) * inv); out[1] = (float)((c * h - b * i) * inv); out[2] = (float)((b * f - c * e) * inv);
    out[3] = (float)((f * g - d * i) * inv); out[4] = (float)((a * i - c * g) * inv); out[5] = (float)((c * d - a * f) * inv);
    out[6] = (float)((d * h - e * g) * inv); out[7] = (float)((b * g - a * h) * inv); out[8] = (float)((a * e - b * d) * inv);
}

void upload_mesh_if_dirty(nmr_ctx* ctx) {
    if (!ctx->mesh_dirty) return;
    std::vector<float> wpos, wnrm, uv, wtbn;
    std::vector<uint32_t> idx;
    float nmat0[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    std::vector<uint8_t> tri_lens;
    bool any_lens = false;
    for (const auto& m : ctx->meshes) {
        std::vector<float> p, n;
        transform_mesh(m->host, m->t, m->s, m->r, p, n);
        const uint32_t base = (uint32_t)(wpos.size() / 3);
        wpos.insert(wpos.end(), p.begin(), p.end());
        wnrm.insert(wnrm.end(), n.begin(), n.end());
        uv.insert(uv.end(), m->host.texcoords.begin(), m->host.texcoords.end());
        {
            std::vector<float> f; float nm[9];
            transform_tangent_frames(m->host, m->s, m->r, f, nm);
            wtbn.insert(wtbn.end(), f.begin(), f.end());
            if (&m == &ctx->meshes.front()) std::memcpy(nmat0, nm, sizeof(nm));
        }
        for (uint32_t i : m->host.indices) idx.push_back(base + i);
        tri_lens.insert(tri_lens.end(), m->host.tri_lens.begin(), m->host.tri_lens.end());
        tri_lens.resize(idx.size() / 3, 0);
        if (m->host.has_lens && !any_lens) {   // lens parameters of the first lens material in the scene (nmr_set_lens overrides)
            any_lens = true;
            if (!ctx->lens_params_set) { ctx->lens_ior = m->host.lens_ior; ctx->lens_transmission = m->host.lens_transmission; std::memcpy(ctx->lens_tint, m->host.lens_tint, 12); }
        }
    }
    ctx->scene_has_lens = any_lens;
    MeshDevice d{};
    d.n_tris = (uint32_t)(idx.size() / 3);
    for (int k = 0; k < 3; ++k) { ctx->mesh_wmin[k] = 1e30f; ctx->mesh_wmax[k] = -1e30f; }
    for (size_t i = 0; i < wpos.size(); ++i) {
        ctx->mesh_wmin[i % 3] = std::min(ctx->mesh_wmin[i % 3], wpos[i]);
        ctx->mesh_wmax[i % 3] = std::max(ctx->mesh_wmax[i % 3], wpos[i]);
    }
    if (d.n_tris) {
        ctx->d_wpos.ensure(wpos.size()); ctx->d_wnrm.ensure(wnrm.size()); ctx->d_uv.ensure(uv.size()); ctx->d_idx.ensure(idx.size());
        CK(cudaMemcpyAsync(ctx->d_wpos.p, wpos.data(), wpos.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_wnrm.p, wnrm.data(), wnrm.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_uv.p, uv.data(), uv.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_idx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (any_lens) {
            ctx->d_tri_lens.ensure(tri_lens.size());
            CK(cudaMemcpyAsync(ctx->d_tri_lens.p, tri_lens.data(), tri_lens.size(), cudaMemcpyHostToDevice, ctx->stream));
            d.tri_lens = ctx->d_tri_lens.p;
        }
        const HostMesh& m0 = ctx->meshes[0]->host;   // one material for the whole scene (first mesh), see DESIGN.md
        std::memcpy(d.base_color, m0.base_color, 16); std::memcpy(d.emissive, m0.emissive, 12);
        d.metallic = m0.metallic; d.roughness = m0.roughness;
        std::vector<float> tex;
        if (!m0.tex_rgba8.empty()) {
            tex.resize((size_t)m0.tex_w * m0.tex_h * 4);
            for (size_t i = 0; i < (size_t)m0.tex_w * m0.tex_h; ++i) {
                for (int k = 0; k < 3; ++k) {
                    const float sv = (float)m0.tex_rgba8[i * 4 + k] / 255.0f;
                    tex[i * 4 + k] = sv <= 0.04045f ? sv / 12.92f : powf((sv + 0.055f) / 1.055f, 2.4f);
                }
                tex[i * 4 + 3] = (float)m0.tex_rgba8[i * 4 + 3] / 255.0f;
            }
            ctx->d_tex.ensure(tex.size());
            CK(cudaMemcpyAsync(ctx->d_tex.p, tex.data(), tex.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            d.tex_lin = ctx->d_tex.p; d.tex_w = m0.tex_w; d.tex_h = m0.tex_h;
        }
        // emissive (sRGB), metallic-roughness, normal, occlusion (linear) textures of the first mesh's material
        const HostMesh::Texture* xt[4] = {&m0.tex_emissive, &m0.tex_metallic_roughness, &m0.tex_normal, &m0.tex_occlusion};
        TexDev* xd[4] = {&d.tex_emissive, &d.tex_mr, &d.tex_normal, &d.tex_occ};
        std::vector<float> xbuf[4];
        for (int k = 0; k < 4; ++k) {
            *xd[k] = TexDev{nullptr, 0, 0};
            if (xt[k]->rgba8.empty()) continue;
            const size_t px = (size_t)xt[k]->w * xt[k]->h;
            xbuf[k].resize(px * 4);
            for (size_t i = 0; i < px; ++i) {
                for (int c = 0; c < 3; ++c) {
                    const float sv = (float)xt[k]->rgba8[i * 4 + c] / 255.0f;
                    xbuf[k][i * 4 + c] = k == 0 ? (sv <= 0.04045f ? sv / 12.92f : powf((sv + 0.055f) / 1.055f, 2.4f)) : sv;
                }
                xbuf[k][i * 4 + 3] = (float)xt[k]->rgba8[i * 4 + 3] / 255.0f;
            }
            ctx->d_xtex[k].ensure(xbuf[k].size());
            CK(cudaMemcpyAsync(ctx->d_xtex[k].p, xbuf[k].data(), xbuf[k].size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            *xd[k] = TexDev{ctx->d_xtex[k].p, xt[k]->w, xt[k]->h};
        }
        d.normal_scale = m0.normal_scale; d.occlusion_strength = m0.occlusion_strength;
        std::memcpy(d.nmat, nmat0, sizeof(nmat0));
        d.wtbn = nullptr;
        if (d.tex_normal.p) {
            ctx->d_wtbn.ensure(wtbn.size());
            CK(cudaMemcpyAsync(ctx->d_wtbn.p, wtbn.data(), wtbn.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            d.wtbn = ctx->d_wtbn.p;
        }
        CK(cudaStreamSynchronize(ctx->stream));   // host vectors go out of scope
        d.wpos = ctx->d_wpos.p; d.wnrm = ctx->d_wnrm.p; d.uv = ctx->d_uv.p; d.idx = ctx->d_idx.p;
    }
    ctx->mesh_dev = d;
    ctx->mesh_dirty = false;
}

// World-space box around every occupied cell a sample can test, inflated by one cell of the coarsest cascade involved.
// A position p tests cascade mip_from_pos(p) (S/ngp/testbed.cu:188-193): inside the unit cube that is cascade 0 (cascade 1
// exactly on its boundary), in general cascades 0..max_cascade+1; cascade c cells are 2^c / 128 wide and centred on 0.5.
void update_occupied_box(nmr_ctx* ctx, Nerf& n) {
    {   // the coarse view of cascade 0 that the first-hit walk jumps through (NMR_NO_COARSE=1: plain walk, for A/B runs)
        n.d_coarse.ensure((size_t)kCoarseRes * kCoarseRes * kCoarseRowWords);
        DevBuf<uint8_t> d_occ; d_occ.ensure((size_t)kCoarseRes * kCoarseRes * kCoarseRes);
        launch_coarse_build(n.d_bitfield.p, d_occ.p, n.d_coarse.p, ctx->stream);
        CK(cudaStreamSynchronize(ctx->stream));
        const char* off = std::getenv("NMR_NO_COARSE");
        n.dev.coarse = (off && off[0] == '1') ? nullptr : n.d_coarse.p;
    }
    DevBuf<int> d_b; d_b.ensure(48);
    launch_occupancy_bounds(n.d_bitfield.p, d_b.p, ctx->stream);
    int b[48];
    CK(cudaMemcpyAsync(b, d_b.p, sizeof(b), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    const int top = std::min(n.host.max_cascade + 1, (int)kCascades - 1);
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int c = 0; c <= top; ++c) {
        const float cell = std::ldexp(1.0f, c) / 128.0f;
        for (int k = 0; k < 3; ++k) {
            if (b[c * 6 + 3 + k] < 0) continue;
            lo[k] = std::min(lo[k], ((float)b[c * 6 + k] / 128.0f - 0.5f) * std::ldexp(1.0f, c) + 0.5f);
            hi[k] = std::max(hi[k], ((float)(b[c * 6 + 3 + k] + 1) / 128.0f - 0.5f) * std::ldexp(1.0f, c) + 0.5f);
        }
        (void)cell;
    }
    const float margin = std::ldexp(1.0f, top) / 128.0f;
    for (int k = 0; k < 3; ++k) {
        if (hi[k] < lo[k]) { n.occ_min[k] = 1.f; n.occ_max[k] = -1.f; }     // nothing occupied: every ray misses
        else { n.occ_min[k] = lo[k] - margin; n.occ_max[k] = hi[k] + margin; }
    }
}

// Screen bounding box of the mesh's world-space box in the supersampled mesh frame, padded by two sub-pixels and aligned
// to whole pixels; the whole frame when a corner lies behind the eye.  Only that window of the visibility buffer is
// cleared, rasterised into and read back.
void mesh_screen_box(const nmr_ctx* ctx, FrameParams& P) {
    P.zb_x0 = P.zb_y0 = P.zb_w = P.zb_h = 0;
    if (P.mesh_scale <= 0) return;
    const int ms = P.mesh_scale, W2 = P.width * ms, H2 = P.height * ms;
    float minx = 1e30f, miny = 1e30f, maxx = -1e30f, maxy = -1e30f;
    bool behind = false;
    for (int c = 0; c < 8 && !behind; ++c) {
        const float q[3] = {((c & 1) ? ctx->mesh_wmax[0] : ctx->mesh_wmin[0]) - P.cam[9], ((c & 2) ? ctx->mesh_wmax[1] : ctx->mesh_wmin[1]) - P.cam[10],
                            ((c & 4) ? ctx->mesh_wmax[2] : ctx->mesh_wmin[2]) - P.cam[11]};
        const float a = P.cam_inv[0] * q[0] + P.cam_inv[1] * q[1] + P.cam_inv[2] * q[2];
        const float b = P.cam_inv[3] * q[0] + P.cam_inv[4] * q[1] + P.cam_inv[5] * q[2];
        const float w = P.cam_inv[6] * q[0] + P.cam_inv[7] * q[1] + P.cam_inv[8] * q[2];
        if (!(w > 1e-3f)) { behind = true; break; }
        const float px = (a / w + 1.0f) * 0.5f * (float)W2 - 0.5f, py = (b / w + 1.0f) * 0.5f * (float)H2 - 0.5f;
        minx = std::min(minx, px); maxx = std::max(maxx, px); miny = std::min(miny, py); maxy = std::max(maxy, py);
    }
    int x0 = 0, y0 = 0, x1 = W2, y1 = H2;    // [x0, x1) x [y0, y1)
    if (!behind) {
        if (!(maxx >= -4.f && maxy >= -4.f && minx <= (float)W2 + 4.f && miny <= (float)H2 + 4.f)) return;   // off screen
        x0 = std::max(0, ((int)std::floor(minx) - 2) / ms * ms); y0 = std::max(0, ((int)std::floor(miny) - 2) / ms * ms);
        x1 = std::min(W2, (((int)std::ceil(maxx) + 3 + ms - 1) / ms) * ms); y1 = std::min(H2, (((int)std::ceil(maxy) + 3 + ms - 1) / ms) * ms);
        if (x1 <= x0 || y1 <= y0) return;
    }
    P.zb_x0 = x0; P.zb_y0 = y0; P.zb_w = x1 - x0; P.zb_h = y1 - y0;
}

// Screen rectangle (pixels) of the box around the occupied cells: a ray through a pixel outside it misses the box, so
// advance_pos_nerf could only walk it out of the render box - a background pixel unless the mesh covers it.  Conservative:
// the whole frame when a corner of the box is not safely in front of the eye; padded by two pixels.
void occupied_screen_box(FrameParams& P) {
    P.occ_px[0] = P.occ_px[1] = 0; P.occ_px[2] = P.width; P.occ_px[3] = P.height;
    if (!(P.occ_min[0] <= P.occ_max[0])) { P.occ_px[2] = P.occ_px[3] = 0; return; }       // nothing occupied
    float minx = 1e30f, miny = 1e30f, maxx = -1e30f, maxy = -1e30f;
    for (int c = 0; c < 8; ++c) {
        // NeRF-space corner relative to the NeRF-space eye (cam + 0.5)
        // a NeRF-space point p lies on the ray of the pixel whose camera direction d satisfies R d || p - origin, i.e. d || R^T (p - origin)
        const float v[3] = {((c & 1) ? P.occ_max[0] : P.occ_min[0]) - P.ray_origin[0], ((c & 2) ? P.occ_max[1] : P.occ_min[1]) - P.ray_origin[1],
                            ((c & 4) ? P.occ_max[2] : P.occ_min[2]) - P.ray_origin[2]};
        const float* R = P.model_rot;
        const float q[3] = {R[0] * v[0] + R[3] * v[1] + R[6] * v[2], R[1] * v[0] + R[4] * v[1] + R[7] * v[2], R[2] * v[0] + R[5] * v[1] + R[8] * v[2]};
        const float a = P.cam_inv[0] * q[0] + P.cam_inv[1] * q[1] + P.cam_inv[2] * q[2];
        const float b = P.cam_inv[3] * q[0] + P.cam_inv[4] * q[1] + P.cam_inv[5] * q[2];
        const float w = P.cam_inv[6] * q[0] + P.cam_inv[7] * q[1] + P.cam_inv[8] * q[2];
        if (!(w > 1e-3f)) return;                                                           // whole frame
        const float px = (a / w + 1.0f) * 0.5f * (float)P.width - 0.5f, py = (b / w + 1.0f) * 0.5f * (float)P.height - 0.5f;
        minx = std::min(minx, px); maxx = std::max(maxx, px); miny = std::min(miny, py); maxy = std::max(maxy, py);
    }
    if (!(maxx >= -4.f && maxy >= -4.f && minx <= (float)P.width + 4.f && miny <= (float)P.height + 4.f)) { P.occ_px[2] = P.occ_px[3] = 0; return; }   // off screen
    P.occ_px[0] = std::max(0, (int)std::floor(minx) - 2); P.occ_px[1] = std::max(0, (int)std::floor(miny) - 2);
    P.occ_px[2] = std::min(P.width, (int)std::ceil(maxx) + 3); P.occ_px[3] = std::min(P.height, (int)std::ceil(maxy) + 3);
    if (P.occ_px[2] <= P.occ_px[0] || P.occ_px[3] <= P.occ_px[1]) { P.occ_px[2] = P.occ_px[3] = 0; P.occ_px[0] = P.occ_px[1] = 0; }
}

FrameParams make_params(nmr_ctx* ctx, const Nerf& n, int W, int H, const float* cam12, uint32_t spp_index, bool to_srgb, bool with_mesh) {
    FrameParams P{};
    P.width = W; P.height = H;
    std::memcpy(P.cam, cam12, sizeof(P.cam));
    std::memcpy(P.aabb_min, n.render_aabb_min, 12); std::memcpy(P.aabb_max, n.render_aabb_max, 12);
    std::memcpy(P.r2l, n.host.render_aabb_to_local, sizeof(P.r2l));
    std::memcpy(P.taabb_min, n.host.aabb_min, 12); std::memcpy(P.taabb_max, n.host.aabb_max, 12);
    P.cone_angle = n.host.cone_angle_constant;
    P.spp_index = spp_index;
    static const int debug_pixel = [] { const char* v = std::getenv("NMR_DEBUG_PIXEL"); return v ? std::atoi(v) : -1; }();
    P.debug_pixel = debug_pixel;
    P.min_transmittance = n.min_transmittance;
    P.rgb_activation = n.host.rgb_activation; P.density_activation = n.host.density_activation;
    std::memcpy(P.background, n.background, 16);
    for (int k = 0; k < 3; ++k) {   // tonemap_kernel linearises the sRGB background colour (S/ngp/render_buffer.cu:548-550)
        const float sv = n.background[k];
        P.background_linear[k] = sv <= 0.04045f ? sv / 12.92f : powf((sv + 0.055f) / 1.055f, 2.4f);
    }
    P.to_srgb = to_srgb ? 1 : 0;
    P.tonemap_curve = n.tonemap_curve;
    {   // accumulate_kernel + tonemap_kernel of an empty pixel (S/ngp/render_buffer.cu:232-267, 537-566)
        const float w = (1.f - 0.f) * n.background[3];
        float c3[3];
        for (int k = 0; k < 3; ++k) c3[k] = 0.f + P.background_linear[k] * w;
        tonemap_curve_apply(c3[0], c3[1], c3[2], n.tonemap_curve);
        for (int k = 0; k < 3; ++k) {
            float c = c3[k];
            if (to_srgb) { c = c < 0.0031308f ? 12.92f * c : 1.055f * powf(c, 0.41666f) - 0.055f; c = std::min(std::max(c, 0.f), 1.f); }
            P.background_out[k] = c;
        }
        P.background_out[3] = to_srgb ? std::min(std::max(0.f + w, 0.f), 1.f) : 0.f + w;
    }
    P.shard_rank = ctx->shard_rank; P.shard_world = ctx->shard_world; P.shard_band = ctx->shard_band;
    P.mesh_scale = (with_mesh && ctx->mesh_dev.n_tris > 0) ? ctx->mesh_scale : 0;
    std::memcpy(P.light, ctx->light, 12);
    invert3(cam12, P.cam_inv);
    std::memcpy(P.occ_min, n.occ_min, 12); std::memcpy(P.occ_max, n.occ_max, 12);
    P.surface_mode = ctx->surface_mode;
    {   // model matrix: ray origin = R eye + 0.5 + R t_model in fp32, sums left to right (the oracle restates the same expression)
        std::memcpy(P.model_rot, n.model_rot, sizeof(P.model_rot));
        const float* R = n.model_rot; const float* e = cam12 + 9; const float* tm = n.model_translation;
        const bool identity = R[0] == 1.f && R[4] == 1.f && R[8] == 1.f && R[1] == 0.f && R[2] == 0.f && R[3] == 0.f && R[5] == 0.f && R[6] == 0.f && R[7] == 0.f;
        for (int k = 0; k < 3; ++k) {
            if (identity) P.ray_origin[k] = (tm[k] != 0.f) ? (e[k] + 0.5f) + tm[k] : e[k] + 0.5f;
            else {
                const float ro = (R[k * 3] * e[0] + R[k * 3 + 1] * e[1]) + R[k * 3 + 2] * e[2];
                const float rt = (R[k * 3] * tm[0] + R[k * 3 + 1] * tm[1]) + R[k * 3 + 2] * tm[2];
                P.ray_origin[k] = (ro + 0.5f) + rt;
            }
        }
    }
    occupied_screen_box(P);
    mesh_screen_box(ctx, P);
    P.lens_on = (P.mesh_scale > 0 && ctx->scene_has_lens && ctx->lens_enabled) ? 1 : 0;
    {
        const float r0 = (ctx->lens_ior - 1.f) / (ctx->lens_ior + 1.f);
        P.lens_f0 = r0 * r0;
        for (int k = 0; k < 3; ++k) P.lens_k[k] = ctx->lens_transmission * ctx->lens_tint[k];
        P.lens_kmean = (P.lens_k[0] + P.lens_k[1] + P.lens_k[2]) / 3.f;
        P.lens_model = ctx->lens_model; P.lens_thickness = ctx->lens_thickness; P.lens_ior = ctx->lens_ior;
    }
    return P;
}

// Frames that can need the reference's per-iteration n_steps schedule: a mesh is in view and the surface rule is `auto`.
// Returns false (and leaves `sa` zeroed) otherwise.  The histogram is cleared on the stream.
bool prepare_schedule(nmr_ctx* ctx, const FrameParams& P, SchedArgs& sa, bool clear = true, bool* proved_batch8 = nullptr) {
    sa = SchedArgs{};
    if (proved_batch8) *proved_batch8 = false;
    if (!(P.mesh_scale > 0 && P.zb_w > 0 && P.surface_mode == kSurfaceAuto)) return false;
    {   // Only pixels inside the two screen rectangles can queue a ray.  When even ALL of them together are at most 1/8 of the pixels
        // this pass traces, the kernel's own rule (live rays * 8 <= pixels) must come out as "8-sample batches": the schedule replay -
        // histogram, surface list, second march launch - cannot be needed, whatever the device finds.
        const long long ms = P.mesh_scale;
        const long long occ = (long long)std::max(0, P.occ_px[2] - P.occ_px[0]) * std::max(0, P.occ_px[3] - P.occ_px[1]);
        const long long mesh_px = ((long long)P.zb_w / ms + 2) * ((long long)P.zb_h / ms + 2);
        const long long traced = (long long)P.width * rows_owned_by(P.height, P.shard_rank, P.shard_world, P.shard_band);
        if ((occ + mesh_px) * 8 <= traced) { if (proved_batch8) *proved_batch8 = true; return false; }
    }
    Surfaces& S = ctx->surf;
    S.hist.ensure(kSchedBins); S.surf_list.ensure((size_t)P.width * P.height);
    if (clear) CK(cudaMemsetAsync(S.hist.p, 0, sizeof(uint32_t) * kSchedBins, ctx->stream));
    sa.hist = S.hist.p; sa.surf_list = S.surf_list.p; sa.pass = 1;
    return true;
}
// second pass over the rays that carry a mesh surface (exits at once unless more than 1/8 of the pixels were live)
void enqueue_surface_pass(nmr_ctx* ctx, Nerf& n, const FrameParams& P, const FrameOut& out, uint32_t n_pixels, SchedArgs sa) {
    Surfaces& S = ctx->surf;
    sa.pass = 2;
    launch_march(P, model_for(ctx, n), S.queue.p, ctx->d_counters.p, out, n_pixels, ctx->debug_flags, ctx->num_sms, ctx->stream, ctx->d_counters.p + 7, ctx->d_counters.p + 6, &sa, 1);
}

// one sample-per-pixel pass: mesh stage -> init -> march.  Enqueues only; no host synchronisation.
void enqueue_pass(nmr_ctx* ctx, Nerf& n, const FrameParams& P, bool timed, void* image_target = nullptr, bool force_probes = false) {
    Surfaces& S = ctx->surf;
    const int rows = rows_owned_by(P.height, P.shard_rank, P.shard_world, P.shard_band);
    // the linear frame, depth and per-ray sample counts are parity probes (nmr_debug_last_frame): written only on request
    // (or when several NeRFs are merged by their depth buffers)
    const bool probes = force_probes || (ctx->debug_flags & kDebugKeepProbes) != 0;
    FrameOut out{image_target ? image_target : static_cast<void*>(S.image.p), S.accum.p, probes ? S.frame.p : nullptr, probes ? S.depth.p : nullptr, probes ? S.n_samples.p : nullptr, nullptr, nullptr};
    if (!image_target) { S.image_format = P.out_format; S.image_is_frame = false; }
    // Experiment (NMR_L2_PERSIST=1, profiles/r2_experiments.md): the hash table of the NeRF being rendered inside a persisting L2
    // access-policy window of the frame's stream, instead of / next to the prefetch of launch_background.
    static const bool l2_persist = std::getenv("NMR_L2_PERSIST") != nullptr;
    if (l2_persist && ctx->l2_window_ptr != static_cast<const void*>(n.dev.grid)) {
        cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
        const size_t table_bytes = ((size_t)n.dev.level_offset[N_LEVELS - 1] + n.dev.level_size[N_LEVELS - 1]) * sizeof(__half2);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min((size_t)prop.persistingL2CacheMaxSize, table_bytes)));
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = const_cast<__half2*>(n.dev.grid);
        attr.accessPolicyWindow.num_bytes = std::min(table_bytes, (size_t)prop.accessPolicyMaxWindowSize);
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
        ctx->l2_window_ptr = n.dev.grid;
        std::fprintf(stderr, "libnmr: persisting L2 window over %zu MiB of hash table (device limits: %d MiB persisting, %d MiB window)\n", table_bytes >> 20, prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
    }
    static const char* phase_log_path = std::getenv("NMR_PHASE_LOG");      // measurement aid: see FrameOut::phase_log
    if (phase_log_path && timed) {
        const size_t words = (size_t)ctx->num_sms * 4 * 2 * kPhaseIters * kPhaseWords;
        ctx->d_phase_log.ensure(words);
        CK(cudaMemsetAsync(ctx->d_phase_log.p, 0, words * 8, ctx->stream));
        out.phase_log = ctx->d_phase_log.p;
    }
    MeshDevice mesh = ctx->mesh_dev;
    if (!P.lens_on) mesh.tri_lens = nullptr;     // lenses off: their triangles are ordinary opaque surfaces
    else {
        S.lens.ensure((size_t)P.width * P.height * 2);
        S.lens_scratch.ensure((size_t)ctx->num_sms * 4 * 32 * kLensStash);   // launch_march: num_sms x 3 CTAs x 32 ray groups
        out.lens = S.lens.p; out.lens_scratch = S.lens_scratch.p;
    }
    // (the schedule's buffers are allocated here, in front of the frame's opening event: a first allocation is not frame time)
    SchedArgs sa;
    bool proved_batch8 = false;
    const bool sched = prepare_schedule(ctx, P, sa, false, &proved_batch8);
    if (timed) CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    uint64_t launches = 0;
    // OVERLAPPED frame: the march kernel starts while the set-up kernel is still running and consumes the queue as it fills
    // (kernels.cu: march_kernel).  Possible whenever the march kernel does not need the frame's final live-ray count up front,
    // i.e. whenever the mesh-surface rule is known on the host: no mesh in view, a pinned rule, or the screen rectangles prove the
    // 8-sample batches (every BASELINE config); close-ups under the auto rule keep the serial order (they need the count and the
    // death histogram).  The CUDA-core bring-up variant stays serial as well.
    static const bool no_overlap_env = std::getenv("NMR_NO_OVERLAP") != nullptr;
    const bool overlap = ctx->overlap && !no_overlap_env && !sched && !(ctx->debug_flags & kDebugScalarMlp);
    FrameParams Pm = P;
    if (overlap && proved_batch8) Pm.surface_mode = kSurfaceBatch8;
    // counters, schedule histogram, the mesh visibility window and the ready words of the last frame's queue records are
    // cleared by one kernel (not four memset nodes)
    // NMR_TIMELINE=1: where a frame's time goes kernel by kernel (events between the kernels - they serialise the frame, so this
    // is for serial frames, NMR_NO_OVERLAP=1); printed by the next nmr_get_stats
    static const bool timeline = std::getenv("NMR_TIMELINE") != nullptr;
    const bool tl = timeline && timed && !overlap;
    if (tl) { for (auto& e : ctx->ev_tl) if (!e) CK(cudaEventCreate(&e)); CK(cudaEventRecord(ctx->ev_tl[0], ctx->stream)); }
    launch_frame_clear(ctx->d_counters.p, sched ? S.hist.p : nullptr, S.zbuf.p, zbuf_window_words(mesh, P), S.queue.p, ctx->stream);
    launches += 1;
    if (tl) CK(cudaEventRecord(ctx->ev_tl[1], ctx->stream));
    // The background pixels (everything outside the tile box of the two screen rectangles) depend on nothing else in the frame:
    // they are written on a side stream while the mesh stage and the ray set-up run (fork behind the clear kernel, which is
    // behind the previous frame; join in front of the frame's closing event).  NMR_NO_AUX_STREAM=1: in line, for A/B runs.
    const TileBox box = compute_tile_box(P, rows);
    static const bool no_aux = std::getenv("NMR_NO_AUX_STREAM") != nullptr;
    const bool side = !no_aux && ctx->aux_stream != nullptr && rows > 0;
    if (side) {
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
        launch_background(P, n.dev, out, rows, box, true, ctx->num_sms, ctx->aux_stream);
        CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    } else {
        launch_background(P, n.dev, out, rows, box, true, ctx->num_sms, ctx->stream);
    }
    launches += rows > 0 ? 1 : 0;
    if (P.mesh_scale > 0) { launch_mesh_raster(mesh, P, rows, S.zbuf.p, ctx->stream, false); launches += 1; }
    if (tl) CK(cudaEventRecord(ctx->ev_tl[2], ctx->stream));
    const int init_ctas = launch_init_rays(P, n.dev, mesh, S.zbuf.p, rows, S.queue.p, ctx->d_counters.p, out, box, ctx->stream, sched ? S.surf_list.p : nullptr);
    launches += init_ctas > 0 ? 1 : 0;
    if (timed && !overlap) CK(cudaEventRecord(ctx->ev[1], ctx->stream));      // (an event between the two kernels would serialise them)
    if (tl) CK(cudaEventRecord(ctx->ev_tl[3], ctx->stream));
    const uint32_t n_pixels = (uint32_t)P.width * (uint32_t)rows;
    launch_march(Pm, model_for(ctx, n), S.queue.p, ctx->d_counters.p, out, n_pixels, ctx->debug_flags, ctx->num_sms, ctx->stream, nullptr, nullptr, sched ? &sa : nullptr, ctx->march_ctas,
                 overlap ? init_ctas : -1);
    launches += 1;
    if (tl) { CK(cudaEventRecord(ctx->ev_tl[4], ctx->stream)); ctx->tl_pending = true; }
    if (side) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    if (timed) ctx->last_overlapped = overlap;
    if (sched) { enqueue_surface_pass(ctx, n, P, out, n_pixels, sa); launches += 1; }
    if (timed) {
        CK(cudaEventRecord(ctx->ev[2], ctx->stream));
        CK(cudaMemcpyAsync(ctx->h_counters, ctx->d_counters.p, sizeof(uint32_t) * kNumCounters, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.rays = (uint64_t)P.width * rows;
        ctx->stats.mesh_rays = P.mesh_scale > 0 ? (uint64_t)P.width * rows * P.mesh_scale * P.mesh_scale : 0;
        ctx->stats.kernel_launches = launches;
        ctx->stats_pending = true;
    }
    CK(cudaGetLastError());
}

// nmr_render's single-sample path: the image crosses PCIe underneath the rendering instead of after it.
// Rows above and below both screen rectangles (occupied box, mesh) are background whatever the GPU does - the host knows them
// before anything is launched, and in render.py's framing they are most of the picture; so are the columns left and right of the
// rectangles in the rows in between.  They leave at once on the copy stream from an image of nothing but the background colour,
// so the copy engine, which bounds a float32 call, starts at time zero; only the rectangle itself (a 2-D copy) waits for the
// frame (one overlapped pass, exactly as in frame()) - for the compact formats that is what is left on the critical path.
// `known` (nmr_render_update): the destination already holds a complete image of this size / format / background whose pixels
// outside known->rect are background (it is the buffer of the previous call and nobody else wrote to it); then only the bounding
// box of the old and the new rectangle is touched at all.  rect_out receives this frame's rectangle, copied the bytes moved.
void enqueue_pass_with_copy(nmr_ctx* ctx, Nerf& n, const FrameParams& P0, void* out_host, const int* known_rect = nullptr, int* rect_out = nullptr, size_t* copied = nullptr) {
    Surfaces& S = ctx->surf;
    const size_t bpp = pixel_bytes(P0.out_format);
    size_t moved = 0;
    auto copy_rows = [&](int y0, int y1, const void* src) {
        if (y1 <= y0) return;
        const size_t off = (size_t)y0 * P0.width * bpp;
        CK(cudaMemcpyAsync(static_cast<char*>(out_host) + off, static_cast<const char*>(src) + off, (size_t)(y1 - y0) * P0.width * bpp, cudaMemcpyDeviceToHost, ctx->copy_stream));
        moved += (size_t)(y1 - y0) * P0.width * bpp;
    };
    auto copy_rect = [&](int y0, int y1, int x0, int x1, const void* src) {      // columns [x0, x1) of rows [y0, y1)
        if (y1 <= y0 || x1 <= x0) return;
        if (x0 == 0 && x1 == P0.width) { copy_rows(y0, y1, src); return; }
        const size_t pitch = (size_t)P0.width * bpp, off = (size_t)y0 * pitch + (size_t)x0 * bpp;
        CK(cudaMemcpy2DAsync(static_cast<char*>(out_host) + off, pitch, static_cast<const char*>(src) + off, pitch, (size_t)(x1 - x0) * bpp, (size_t)(y1 - y0), cudaMemcpyDeviceToHost, ctx->copy_stream));
        moved += (size_t)(x1 - x0) * bpp * (size_t)(y1 - y0);
    };
    int known_top = 0, known_bot = P0.height;          // rows [0, known_top) and [known_bot, H) are known background
    int known_left = 0, known_right = P0.width;        // and so are columns [0, known_left) and [known_right, W) of the rows in between
    {
        int r0 = P0.height, r1 = 0, c0 = P0.width, c1 = 0;   // pixels that may hold anything else: union of the rectangles
        if (P0.occ_px[2] > P0.occ_px[0] && P0.occ_px[3] > P0.occ_px[1]) {
            r0 = std::min(r0, P0.occ_px[1]); r1 = std::max(r1, P0.occ_px[3]); c0 = std::min(c0, P0.occ_px[0]); c1 = std::max(c1, P0.occ_px[2]);
        }
        if (P0.mesh_scale > 0 && P0.zb_w > 0 && P0.zb_h > 0) {
            r0 = std::min(r0, P0.zb_y0 / P0.mesh_scale); r1 = std::max(r1, (P0.zb_y0 + P0.zb_h + P0.mesh_scale - 1) / P0.mesh_scale);
            c0 = std::min(c0, P0.zb_x0 / P0.mesh_scale); c1 = std::max(c1, (P0.zb_x0 + P0.zb_w + P0.mesh_scale - 1) / P0.mesh_scale);
        }
        if (r1 <= r0 || c1 <= c0) { known_top = P0.height; known_bot = P0.height; }
        else {
            known_top = std::max(0, r0); known_bot = std::min(P0.height, r1);
            known_left = std::max(0, c0) & ~15; known_right = std::min(P0.width, (c1 + 15) & ~15);      // (whole 64-byte runs of the narrowest format)
        }
        static const bool off = std::getenv("NMR_NO_KNOWN_ROWS") != nullptr;      // A/B aid
        if (off) { known_top = 0; known_bot = P0.height; }
        static const bool off_cols = std::getenv("NMR_NO_KNOWN_COLS") != nullptr;
        // Strips cost three more (2-D) copies: a win when the rectangle's copy is what the call waits for after the frame (sRGB8:
        // 3230 -> 3650 calls/s at 1080p), a loss when the call is bound by PCIe as a whole (float32: 1500 -> 1440;
        // profiles/r2_e2e_paths.txt).  Images up to 12 MiB take the strips.
        if (!known_rect && (off || off_cols || (size_t)P0.width * P0.height * bpp > ((size_t)12 << 20))) { known_left = 0; known_right = P0.width; }
        const size_t px = (size_t)P0.width * P0.height;
        if (S.bg_pixels != px || S.bg_format != P0.out_format || std::memcmp(S.bg_value, P0.background_out, 16) != 0) {
            S.bg_image.ensure(px);      // (float4 capacity: enough for every format)
            FrameParams Pb = P0;        // every pixel "outside both rectangles": the fill kernel writes the constant in the frame's format
            Pb.occ_px[0] = Pb.occ_px[1] = Pb.occ_px[2] = Pb.occ_px[3] = 0; Pb.mesh_scale = 0; Pb.zb_w = Pb.zb_h = 0;
            launch_fill_background(Pb, S.bg_image.p, ctx->stream);                                       // once per resolution / background colour / format
            CK(cudaStreamSynchronize(ctx->stream));
            std::memcpy(S.bg_value, P0.background_out, 16); S.bg_pixels = px; S.bg_format = P0.out_format;
        }
    }
    // The frame's kernels are submitted FIRST: copies submitted to the copy stream ahead of them hold them back until the copies
    // are done (measured: 0.87 ms per float32 call against 0.65 ms this way round, profiles/r2_experiments.md).
    enqueue_pass(ctx, n, P0, true);
    if (known_top >= known_bot) { known_left = 0; known_right = 0; }                 // (nothing but background: an empty rectangle)
    if (rect_out) { rect_out[0] = known_left; rect_out[1] = known_top; rect_out[2] = known_right; rect_out[3] = known_bot; }
    // what leaves at once from the background image: everything outside this frame's rectangle - or, when the destination is
    // known to hold background outside the previous frame's rectangle, just the part of that rectangle this frame no longer covers
    int bx0 = 0, by0 = 0, bx1 = P0.width, by1 = P0.height;
    if (known_rect && known_rect[2] > known_rect[0] && known_rect[3] > known_rect[1]) {
        if (known_top < known_bot) { bx0 = std::min(known_left, known_rect[0]); by0 = std::min(known_top, known_rect[1]); bx1 = std::max(known_right, known_rect[2]); by1 = std::max(known_bot, known_rect[3]); }
        else { bx0 = known_rect[0]; by0 = known_rect[1]; bx1 = known_rect[2]; by1 = known_rect[3]; known_top = known_bot = by0; known_left = bx0; known_right = bx1; }
    } else if (known_rect) { bx0 = known_left; by0 = known_top; bx1 = known_right; by1 = known_bot; }
    copy_rect(by0, known_top, bx0, bx1, S.bg_image.p);
    copy_rect(known_bot, by1, bx0, bx1, S.bg_image.p);
    copy_rect(known_top, known_bot, bx0, known_left, S.bg_image.p);
    copy_rect(known_top, known_bot, known_right, bx1, S.bg_image.p);
    CK(cudaEventRecord(ctx->ev_rows, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_rows, 0));
    copy_rect(known_top, known_bot, known_left, known_right, S.image.p);
    if (copied) *copied = moved;
    CK(cudaGetLastError());
}

void finish_stats(nmr_ctx* ctx) {
    if (!ctx->stats_pending) return;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.rays_alive = ctx->h_counters[0];
    ctx->stats.samples = (uint64_t)ctx->h_counters[2] | ((uint64_t)ctx->h_counters[3] << 32);
    ctx->stats.batches = ctx->h_counters[4];
    ctx->stats.batch_passes = ctx->h_counters[5];
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2])); ctx->stats.gpu_ms = ms;
    if (ctx->last_overlapped) {
        // overlapped frame: no event separates the two kernels; the march kernel's own span (first CTA start to last CTA end,
        // globaltimer) includes the time it waited for the set-up kernel to queue rays
        const uint64_t t0 = (uint64_t)ctx->h_counters[kCntMarchStart] | ((uint64_t)ctx->h_counters[kCntMarchStart + 1] << 32);
        const uint64_t t1 = (uint64_t)ctx->h_counters[kCntMarchEnd] | ((uint64_t)ctx->h_counters[kCntMarchEnd + 1] << 32);
        ctx->stats.march_ms = t1 > t0 ? (float)((double)(t1 - t0) * 1e-6) : 0.f;
    } else {
        CK(cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2])); ctx->stats.march_ms = ms;
    }
    ctx->stats_pending = false;
    if (ctx->tl_pending) {
        float a = 0, b = 0, c = 0, d = 0, e0 = 0, e1 = 0;
        cudaEventElapsedTime(&a, ctx->ev_tl[0], ctx->ev_tl[1]); cudaEventElapsedTime(&b, ctx->ev_tl[1], ctx->ev_tl[2]); cudaEventElapsedTime(&c, ctx->ev_tl[2], ctx->ev_tl[3]);
        cudaEventElapsedTime(&d, ctx->ev_tl[3], ctx->ev_tl[4]); cudaEventElapsedTime(&e0, ctx->ev[0], ctx->ev_tl[0]); cudaEventElapsedTime(&e1, ctx->ev_tl[4], ctx->ev[2]);
        std::fprintf(stderr, "libnmr timeline (us): open %.1f | clear %.1f | mesh raster %.1f | set-up %.1f | march %.1f | close (join of the background stream) %.1f | frame %.1f\n",
                     e0 * 1e3f, a * 1e3f, b * 1e3f, c * 1e3f, d * 1e3f, e1 * 1e3f, ctx->stats.gpu_ms * 1e3f);
        ctx->tl_pending = false;
    }
    if (const char* path = std::getenv("NMR_PHASE_LOG")) {
        if (ctx->d_phase_log.p) {        // raw dump of the last timed frame's phase clocks
            std::vector<unsigned long long> h(ctx->d_phase_log.n);
            CK(cudaMemcpy(h.data(), ctx->d_phase_log.p, h.size() * 8, cudaMemcpyDeviceToHost));
            if (FILE* f = std::fopen(path, "wb")) { std::fwrite(h.data(), 8, h.size(), f); std::fclose(f); }
        }
    }
}

// a helper context of nmr_render_views on the parent's device: streams, events and counters of its own, nothing loaded
nmr_ctx* make_lane(nmr_ctx* parent) {
    std::unique_ptr<nmr_ctx> l(new nmr_ctx());
    l->device = parent->device; l->num_sms = parent->num_sms; l->width = parent->width; l->height = parent->height;
    CK(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&l->aux_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&l->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&l->ev_join, cudaEventDisableTiming));
    for (auto& ev : l->ev) CK(cudaEventCreate(&ev));
    for (auto& pair : l->ev_view) for (auto& ev : pair) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    l->d_counters.ensure(kNumCounters);
    CK(cudaMemset(l->d_counters.p, 0, sizeof(uint32_t) * kNumCounters));
    CK(cudaHostAlloc((void**)&l->h_counters, sizeof(uint32_t) * kNumCounters, cudaHostAllocDefault));
    std::memset(l->h_counters, 0, sizeof(uint32_t) * kNumCounters);
    return l.release();
}
void destroy_lane(nmr_ctx* l) {
    if (!l) return;
    if (l->stream) cudaStreamSynchronize(l->stream);
    if (l->h_counters) cudaFreeHost(l->h_counters);
    for (auto& ev : l->ev) if (ev) cudaEventDestroy(ev);
    for (auto& pair : l->ev_view) for (auto& ev : pair) if (ev) cudaEventDestroy(ev);
    if (l->ev_fork) cudaEventDestroy(l->ev_fork);
    if (l->ev_join) cudaEventDestroy(l->ev_join);
    if (l->aux_stream) { cudaStreamSynchronize(l->aux_stream); cudaStreamDestroy(l->aux_stream); }
    if (l->stream) cudaStreamDestroy(l->stream);
    delete l;
}
// what a lane needs to know about the scene for enqueue_pass: the parent's mesh buffers (borrowed) and settings
void sync_lane(nmr_ctx* l, const nmr_ctx* parent) {
    l->mesh_dev = parent->mesh_dev; l->mesh_scale = parent->mesh_scale; l->debug_flags = parent->debug_flags;
    l->scene_has_lens = parent->scene_has_lens; l->surface_mode = parent->surface_mode; l->overlap = parent->overlap;
    l->lens_enabled = parent->lens_enabled; l->lens_ior = parent->lens_ior; l->lens_transmission = parent->lens_transmission;
    std::memcpy(l->lens_tint, parent->lens_tint, 12); l->lens_model = parent->lens_model; l->lens_thickness = parent->lens_thickness;
}

// frame() with more than one NeRF loaded (NerfMeshRenderer::render_frame, S/nerf_mesh_renderer.cu:561-597): every NeRF renders the
// frame into the per-sample linear frame / depth buffers, the mesh hand-off goes to the first one only (:554-558), the buffers are
// merged by depth (the first NeRF's copied, the others through combineBuffersKernel's rule), and the displayed image is the
// accumulate + tonemap of the merged frame with the first NeRF's background and curve.  (The reference tonemaps every NeRF into
// the same texture, so on its screen the last NeRF simply wins and nothing reads the merged buffers; here they are the picture.)
void enqueue_multi_nerf_frame(nmr_ctx* ctx) {
    Surfaces& S = ctx->surf;
    const uint32_t n_px = (uint32_t)ctx->width * (uint32_t)ctx->height;
    S.comb_frame.ensure(n_px); S.comb_accum.ensure(n_px); S.comb_depth.ensure(n_px); S.image_alt.ensure(n_px);
    FrameParams P0{};
    for (size_t i = 0; i < ctx->nerfs.size(); ++i) {
        Nerf& n = *ctx->nerfs[i];
        const FrameParams P = make_params(ctx, n, ctx->width, ctx->height, ctx->cam12, S.spp, true, i == 0);
        if (i == 0) P0 = P;
        enqueue_pass(ctx, n, P, i + 1 == ctx->nerfs.size(), S.image_alt.p, true);
        launch_combine_buffers(S.depth.p, S.frame.p, S.comb_depth.p, S.comb_frame.p, n_px, i == 0, ctx->stream);
    }
    launch_present(P0, S.comb_frame.p, S.comb_accum.p, S.image.p, n_px, ctx->stream);
    S.image_format = kPixelF32; S.comb_valid = true;
    CK(cudaGetLastError());
}

void set_camera_from_orbit(nmr_ctx* ctx) {
    ctx->camera.matrix(ctx->width, ctx->height, ctx->cam12);
    ctx->surf.spp = 0;   // updateModelViewProj -> reset_accumulation (S/nerf_mesh_renderer.cu:934-938)
}

}  // namespace

extern "C" {

NMR_API int nmr_create(int width, int height, int device, nmr_ctx** out_ctx) {
    if (!out_ctx) return fail(nullptr, NMR_ERR_INVALID, "out_ctx is null");
    *out_ctx = nullptr;
    if (width <= 0 || height <= 0 || width > 16384 || height > 16384) return fail(nullptr, NMR_ERR_INVALID, "bad resolution");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(nullptr, NMR_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0"));
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= count) return fail(nullptr, NMR_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, NMR_ERR_CUDA, cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, NMR_ERR_CUDA, std::string("libnmr is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
    std::unique_ptr<nmr_ctx> ctx(new nmr_ctx());
    ctx->device = device; ctx->num_sms = prop.multiProcessorCount; ctx->width = width; ctx->height = height;
    try {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        for (auto& ev : ctx->ev) CK(cudaEventCreate(&ev));
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        for (auto& pair : ctx->ev_view) for (auto& ev : pair) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_rows, cudaEventDisableTiming));
        ctx->d_counters.ensure(kNumCounters);
        CK(cudaMemset(ctx->d_counters.p, 0, sizeof(uint32_t) * kNumCounters));
        ctx->d_scratch.ensure(64);
        CK(cudaHostAlloc((void**)&ctx->h_counters, sizeof(uint32_t) * kNumCounters, cudaHostAllocDefault));
        std::memset(ctx->h_counters, 0, sizeof(uint32_t) * kNumCounters);
    } catch (const std::exception& ex) {
        return fail(nullptr, NMR_ERR_CUDA, ex.what());
    }
    set_camera_from_orbit(ctx.get());
    if (const char* v = std::getenv("NMR_MLP")) if (!std::strcmp(v, "scalar")) ctx->debug_flags |= kDebugScalarMlp;
    if (const char* v = std::getenv("NMR_UMMA_SWAP")) if (!std::strcmp(v, "1")) ctx->debug_flags |= kDebugSwapLboSbo;
    if (const char* v = std::getenv("NMR_NO_SHARED_ENCODE")) if (!std::strcmp(v, "1")) ctx->debug_flags |= kDebugNoSharedEncode;
    if (const char* v = std::getenv("NMR_NO_BRICKS")) if (!std::strcmp(v, "1")) ctx->debug_flags |= kDebugNoBricks;
    *out_ctx = ctx.release();
    return NMR_OK;
}

NMR_API void nmr_destroy(nmr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->ipc_mapped) { cudaIpcCloseMemHandle(ctx->ipc_mapped); ctx->ipc_mapped = nullptr; }
    for (nmr_ctx* l : ctx->lanes) destroy_lane(l);
    ctx->lanes.clear();
    ctx->nerfs.clear(); ctx->meshes.clear();
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_tl) if (ev) cudaEventDestroy(ev);
    for (auto& pair : ctx->ev_view) for (auto& ev : pair) if (ev) cudaEventDestroy(ev);
    if (ctx->ev_rows) cudaEventDestroy(ctx->ev_rows);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

NMR_API const char* nmr_last_error(const nmr_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

NMR_API void* nmr_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
NMR_API void nmr_host_free(void* p) { if (p) cudaFreeHost(p); }

namespace {
// Testbed::load_snapshot (S/ngp/testbed.cu:939-1135) for this renderer: the snapshot on the host, its parameters, occupancy bits,
// tensor-core weight layout, brick layout and near-field bits on the device.  Throws like the loaders it calls.
std::unique_ptr<Nerf> build_nerf(nmr_ctx* ctx, const char* path) {
        std::unique_ptr<Nerf> n(new Nerf());
        n->host = load_snapshot(path);
        HostModel& h = n->host;
        n->d_params.ensure(h.params.size());
        CK(cudaMemcpyAsync(n->d_params.p, h.params.data(), h.params.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
        n->d_bitfield.ensure(kBitfieldBytes);
        if (!h.density_grid.empty()) {
            DevBuf<uint16_t> d_grid;
            d_grid.ensure(h.density_grid.size());
            CK(cudaMemcpyAsync(d_grid.p, h.density_grid.data(), h.density_grid.size() * 2, cudaMemcpyHostToDevice, ctx->stream));
            launch_occupancy_build(d_grid.p, h.max_cascade + 1, n->d_bitfield.p, ctx->d_scratch.p, ctx->stream);
            CK(cudaStreamSynchronize(ctx->stream));
        } else {
            CK(cudaMemsetAsync(n->d_bitfield.p, 0, kBitfieldBytes, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        CK(cudaGetLastError());
        DeviceModel& d = n->dev;
        d.mlp = reinterpret_cast<const __half*>(n->d_params.p);
        n->d_mlp_tc.ensure(10240);
        launch_weights_canonical(n->d_params.p, n->d_mlp_tc.p, ctx->stream);
        d.mlp_tc = std::getenv("NMR_NO_BULK_WEIGHTS") ? nullptr : reinterpret_cast<const __half*>(n->d_mlp_tc.p);
        d.grid = reinterpret_cast<const __half2*>(n->d_params.p + h.mlp_params);
        d.bitfield = n->d_bitfield.p;
        d.dense_mask = 0; d.pow2_mask = 0;
        {   // prime_hash<COHERENT> / reversed_prime_hash multipliers (T/.../encodings/grid.h:111-146)
            static const uint32_t primes[3][3] = {{1958374283u, 2654435761u, 805459861u}, {1u, 2654435761u, 805459861u}, {2165219737u, 1434869437u, 2097192037u}};
            for (int k = 0; k < 3; ++k) d.prime[k] = primes[h.hash_type][k];
        }
        for (int l = 0; l < h.n_levels; ++l) {
            d.level_offset[l] = h.offsets[l]; d.level_ptr[l] = d.grid + h.offsets[l]; d.level_size[l] = h.offsets[l + 1] - h.offsets[l]; d.level_scale[l] = h.scales[l];
            d.stride_y[l] = h.stride_y[l]; d.stride_z[l] = h.stride_z[l];
            if (h.dense[l]) d.dense_mask |= 1u << l;
            if ((d.level_size[l] & (d.level_size[l] - 1u)) == 0) d.pow2_mask |= 1u << l;
        }
        {   // Brick layout of the coarse levels: a prefix of the levels, as many as fit the budget (NMR_BRICK_MB, default 28 MiB:
            // levels 0-5 of the stock configuration - the five dense levels and the first hashed one, 27 MiB next to a 23 MiB
            // table in a 126 MB L2; 0 switches the layout off).  Cells per axis = the level's resolution: a position in the unit
            // cube never produces a larger cell coordinate (scale <= res - 1).
            static const long brick_mb = [] { const char* v = std::getenv("NMR_BRICK_MB"); return v ? std::atol(v) : 28L; }();
            size_t cells_total = 0; int n_brick = 0;
            for (int l = 0; l < h.n_levels; ++l) {
                const size_t r3 = (size_t)h.resolutions[l] * h.resolutions[l] * h.resolutions[l];
                if (h.resolutions[l] > 1024u || (cells_total + r3) * 32 > (size_t)std::max(0L, brick_mb) << 20) break;
                cells_total += r3; ++n_brick;
            }
            d.n_brick = 0;
            if (n_brick > 0) {
                n->d_bricks.ensure(cells_total * 2);
                size_t off = 0;
                for (int l = 0; l < n_brick; ++l) {
                    const uint32_t r = h.resolutions[l];
                    d.brick[l] = n->d_bricks.p + off * 2; d.brick_res[l] = r;
                    launch_brick_build(d, l, r, n->d_bricks.p + off * 2, ctx->stream);
                    off += (size_t)r * r * r;
                }
                CK(cudaStreamSynchronize(ctx->stream));
                CK(cudaGetLastError());
                d.n_brick = (uint32_t)n_brick;
            }
        }
        std::memcpy(n->render_aabb_min, h.render_aabb_min, 12); std::memcpy(n->render_aabb_max, h.render_aabb_max, 12);
        update_occupied_box(ctx, *n);
        n->n_params = h.params.size();
        std::vector<uint16_t>().swap(h.params);
        std::vector<uint16_t>().swap(h.density_grid);
        return n;
}
}  // namespace

NMR_API int nmr_load_nerf(nmr_ctx* ctx, const char* path, int* out_id) {
    return guarded(ctx, [&]() -> int {
        if (!path) return fail(ctx, NMR_ERR_INVALID, "path is null");
        std::unique_ptr<Nerf> n = build_nerf(ctx, path);
        ctx->nerfs.push_back(std::move(n));
        ctx->surf.spp = 0;
        if (out_id) *out_id = (int)ctx->nerfs.size() - 1;
        return NMR_OK;
    });
}

NMR_API int nmr_reload_nerf(nmr_ctx* ctx, int id, const char* path) {
    return guarded(ctx, [&]() -> int {
        if (!path) return fail(ctx, NMR_ERR_INVALID, "path is null");
        Nerf* old; try { old = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        std::unique_ptr<Nerf> n = build_nerf(ctx, path);              // a file that does not load leaves the old model in place
        // what load_snapshot does not touch in the reference stays: the members set through the Python properties
        std::memcpy(n->background, old->background, sizeof(n->background));
        n->min_transmittance = old->min_transmittance; n->tonemap_curve = old->tonemap_curve;
        std::memcpy(n->model_translation, old->model_translation, 12); std::memcpy(n->model_rotation_pi, old->model_rotation_pi, 12);
        std::memcpy(n->model_rot, old->model_rot, sizeof(n->model_rot));
        CK(cudaDeviceSynchronize());                                  // nothing in flight (helper lanes included) reads the old buffers any more
        ctx->l2_window_ptr = nullptr;
        ctx->nerfs[(size_t)id] = std::move(n);
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}

NMR_API int nmr_load_mesh(nmr_ctx* ctx, const char* path, const float t[3], const float s[3], const float r_wxyz[4], int* out_id) {
    return guarded(ctx, [&]() -> int {
        if (!path) return fail(ctx, NMR_ERR_INVALID, "path is null");
        std::unique_ptr<Mesh> m(new Mesh());
        m->host = load_gltf(path);
        // loadMesh overwrites node 0's TRS with its arguments, defaults t=0, s=1, r=(0,0,0,1) (S/nerf_mesh_renderer.cuh:66-70)
        const float dt[3] = {0.f, 0.f, 0.f}, ds[3] = {1.f, 1.f, 1.f}, dr[4] = {0.f, 0.f, 0.f, 1.f};
        std::memcpy(m->t, t ? t : dt, 12); std::memcpy(m->s, s ? s : ds, 12); std::memcpy(m->r, r_wxyz ? r_wxyz : dr, 16);
        if (!m->host.warning.empty()) ctx->last_error = m->host.warning;
        ctx->meshes.push_back(std::move(m));
        ctx->mesh_dirty = true;
        ctx->surf.spp = 0;
        if (out_id) *out_id = (int)ctx->meshes.size() - 1;
        return NMR_OK;
    });
}

namespace {
// setters refuse values a kernel cannot make sense of (a NaN in the camera would turn screen rectangles into garbage grid sizes)
bool all_finite(const float* v, int n) { if (!v) return true; for (int i = 0; i < n; ++i) if (!std::isfinite(v[i])) return false; return true; }
}  // namespace

NMR_API int nmr_set_mesh_transform(nmr_ctx* ctx, int mesh_id, const float t[3], const float s[3], const float r_wxyz[4]) {
    return guarded(ctx, [&]() -> int {
        if (mesh_id < 0 || mesh_id >= (int)ctx->meshes.size()) return fail(ctx, NMR_ERR_INVALID, "unknown mesh id");
        Mesh& m = *ctx->meshes[(size_t)mesh_id];
        if (!all_finite(t, 3) || !all_finite(s, 3) || !all_finite(r_wxyz, 4)) return fail(ctx, NMR_ERR_INVALID, "mesh transform: values must be finite");
        if (t) std::memcpy(m.t, t, 12);
        if (s) std::memcpy(m.s, s, 12);
        if (r_wxyz) std::memcpy(m.r, r_wxyz, 16);
        ctx->mesh_dirty = true; ctx->surf.spp = 0;
        return NMR_OK;
    });
}
NMR_API int nmr_get_mesh_transform(nmr_ctx* ctx, int mesh_id, float t[3], float s[3], float r_wxyz[4]) {
    return guarded(ctx, [&]() -> int {
        if (mesh_id < 0 || mesh_id >= (int)ctx->meshes.size()) return fail(ctx, NMR_ERR_INVALID, "unknown mesh id");
        const Mesh& m = *ctx->meshes[(size_t)mesh_id];
        if (t) std::memcpy(t, m.t, 12);
        if (s) std::memcpy(s, m.s, 12);
        if (r_wxyz) std::memcpy(r_wxyz, m.r, 16);
        return NMR_OK;
    });
}

NMR_API int nmr_set_envmap(nmr_ctx* ctx, const char* path) {
    return guarded(ctx, [&]() -> int { ctx->envmap_path = path ? path : ""; return NMR_OK; });
}

NMR_API int nmr_get_render_aabb(nmr_ctx* ctx, int id, float mn[3], float mx[3]) {
    return guarded(ctx, [&]() -> int { try { Nerf* n = get_nerf(ctx, id); std::memcpy(mn, n->render_aabb_min, 12); std::memcpy(mx, n->render_aabb_max, 12); return NMR_OK; } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); } });
}
NMR_API int nmr_set_render_aabb(nmr_ctx* ctx, int id, const float mn[3], const float mx[3]) {
    return guarded(ctx, [&]() -> int { try { Nerf* n = get_nerf(ctx, id); if (!all_finite(mn, 3) || !all_finite(mx, 3)) return fail(ctx, NMR_ERR_INVALID, "render_aabb: values must be finite"); if (mn) std::memcpy(n->render_aabb_min, mn, 12); if (mx) std::memcpy(n->render_aabb_max, mx, 12); ctx->surf.spp = 0; return NMR_OK; } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); } });
}
NMR_API int nmr_get_aabb(nmr_ctx* ctx, int id, float mn[3], float mx[3]) {
    return guarded(ctx, [&]() -> int { try { Nerf* n = get_nerf(ctx, id); std::memcpy(mn, n->host.aabb_min, 12); std::memcpy(mx, n->host.aabb_max, 12); return NMR_OK; } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); } });
}
NMR_API int nmr_get_background(nmr_ctx* ctx, int id, float rgba[4]) {
    return guarded(ctx, [&]() -> int { try { std::memcpy(rgba, get_nerf(ctx, id)->background, 16); return NMR_OK; } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); } });
}
NMR_API int nmr_set_background(nmr_ctx* ctx, int id, const float rgba[4]) {
    return guarded(ctx, [&]() -> int { try { if (!rgba || !all_finite(rgba, 4)) return fail(ctx, NMR_ERR_INVALID, "background: four finite values"); std::memcpy(get_nerf(ctx, id)->background, rgba, 16); return NMR_OK; } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); } });
}
NMR_API int nmr_set_tonemap_curve(nmr_ctx* ctx, int id, int curve) {
    return guarded(ctx, [&]() -> int {
        if (curve < 0 || curve > 3) return fail(ctx, NMR_ERR_INVALID, "tonemap curve must be 0 (Identity), 1 (ACES), 2 (Hable) or 3 (Reinhard)");
        try { get_nerf(ctx, id)->tonemap_curve = curve; ctx->surf.spp = 0; return NMR_OK; } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
    });
}
NMR_API int nmr_get_tonemap_curve(nmr_ctx* ctx, int id, int* out_curve) {
    return guarded(ctx, [&]() -> int {
        if (!out_curve) return fail(ctx, NMR_ERR_INVALID, "out_curve is null");
        try { *out_curve = get_nerf(ctx, id)->tonemap_curve; return NMR_OK; } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
    });
}
NMR_API int nmr_set_min_transmittance(nmr_ctx* ctx, int id, float v) {
    return guarded(ctx, [&]() -> int { try { if (!(v >= 0.f && v <= 1.f)) return fail(ctx, NMR_ERR_INVALID, "min_transmittance must lie in [0, 1]"); get_nerf(ctx, id)->min_transmittance = v; return NMR_OK; } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); } });
}

NMR_API int nmr_set_model_transform(nmr_ctx* ctx, int id, const float translation[3], const float rotation_pi[3]) {
    return guarded(ctx, [&]() -> int {
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        if (!all_finite(translation, 3) || !all_finite(rotation_pi, 3)) return fail(ctx, NMR_ERR_INVALID, "model transform: values must be finite");
        if (translation) std::memcpy(n->model_translation, translation, 12);
        if (rotation_pi) std::memcpy(n->model_rotation_pi, rotation_pi, 12);
        // AngleAxis(rx pi, X) * AngleAxis(ry pi, Y) * AngleAxis(rz pi, Z) (S/ngp/testbed.cu:1538-1541), evaluated in double
        const double kPi = 3.14159265358979323846;
        const double ax = n->model_rotation_pi[0] * kPi, ay = n->model_rotation_pi[1] * kPi, az = n->model_rotation_pi[2] * kPi;
        const double cx = std::cos(ax), sx = std::sin(ax), cy = std::cos(ay), sy = std::sin(ay), cz = std::cos(az), sz = std::sin(az);
        const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
        double T[9], Rm[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T[i * 3 + j] = 0; for (int k = 0; k < 3; ++k) T[i * 3 + j] += Rx[i * 3 + k] * Ry[k * 3 + j]; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Rm[i * 3 + j] = 0; for (int k = 0; k < 3; ++k) Rm[i * 3 + j] += T[i * 3 + k] * Rz[k * 3 + j]; }
        const bool zero = n->model_rotation_pi[0] == 0.f && n->model_rotation_pi[1] == 0.f && n->model_rotation_pi[2] == 0.f;
        for (int k = 0; k < 9; ++k) n->model_rot[k] = zero ? (k % 4 == 0 ? 1.f : 0.f) : (float)Rm[k];
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}
NMR_API int nmr_get_model_transform(nmr_ctx* ctx, int id, float translation[3], float rotation_pi[3], float matrix3x3[9]) {
    return guarded(ctx, [&]() -> int {
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        if (translation) std::memcpy(translation, n->model_translation, 12);
        if (rotation_pi) std::memcpy(rotation_pi, n->model_rotation_pi, 12);
        if (matrix3x3) std::memcpy(matrix3x3, n->model_rot, 36);
        return NMR_OK;
    });
}

NMR_API int nmr_get_nerf_info(nmr_ctx* ctx, int id, nmr_nerf_info* o) {
    return guarded(ctx, [&]() -> int {
        if (!o) return fail(ctx, NMR_ERR_INVALID, "out is null");
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        const HostModel& h = n->host;
        *o = nmr_nerf_info{};
        o->training_step = h.training_step; o->loss = h.loss; o->aabb_scale = h.aabb_scale; o->max_cascade = h.max_cascade;
        o->cone_angle_constant = h.cone_angle_constant; o->rgb_activation = h.rgb_activation; o->density_activation = h.density_activation;
        std::memcpy(o->render_aabb_to_local, h.render_aabb_to_local, sizeof(o->render_aabb_to_local));
        o->n_levels = h.n_levels; o->n_features_per_level = h.n_features_per_level; o->log2_hashmap_size = h.log2_hashmap_size;
        o->base_resolution = h.base_resolution; o->per_level_scale = h.per_level_scale; o->n_params = n->n_params;
        return NMR_OK;
    });
}

NMR_API int nmr_get_nerf_dataset(nmr_ctx* ctx, int id, nmr_nerf_dataset* o) {
    return guarded(ctx, [&]() -> int {
        if (!o) return fail(ctx, NMR_ERR_INVALID, "out is null");
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        const HostModel& h = n->host;
        *o = nmr_nerf_dataset{};
        o->scale = h.dataset_scale; o->from_mitsuba = h.from_mitsuba; o->bounding_radius = h.bounding_radius;
        std::memcpy(o->offset, h.dataset_offset, 12); std::memcpy(o->up, h.dataset_up, 12);
        std::memcpy(o->raw_aabb_min, h.aabb_min, 12); std::memcpy(o->raw_aabb_max, h.aabb_max, 12);
        return NMR_OK;
    });
}

NMR_API int nmr_set_render_aabb_to_local(nmr_ctx* ctx, int id, const float m[9]) {
    return guarded(ctx, [&]() -> int {
        if (!m) return fail(ctx, NMR_ERR_INVALID, "matrix is null");
        for (int k = 0; k < 9; ++k) if (!std::isfinite(m[k])) return fail(ctx, NMR_ERR_INVALID, "render_aabb_to_local: values must be finite");
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        std::memcpy(n->host.render_aabb_to_local, m, 36);
        ctx->surf.spp = 0;       // the picture changes: accumulation starts over, as after nmr_set_render_aabb
        return NMR_OK;
    });
}

NMR_API int nmr_orbit(nmr_ctx* ctx, float daz, float dpol, float dzoom) {
    return guarded(ctx, [&]() -> int {
        if (!std::isfinite(daz) || !std::isfinite(dpol) || !std::isfinite(dzoom)) return fail(ctx, NMR_ERR_INVALID, "orbit: values must be finite");
        ctx->camera.orbit(daz, dpol, dzoom); set_camera_from_orbit(ctx); return NMR_OK;
    });
}
NMR_API int nmr_trajectory_pose(nmr_ctx* ctx, float angle, float distance, float height, const float lookat[3]) {
    return guarded(ctx, [&]() -> int {
        const float zero[3] = {0.f, 0.f, 0.f};
        if (!std::isfinite(angle) || !std::isfinite(distance) || !std::isfinite(height) || !all_finite(lookat, 3)) return fail(ctx, NMR_ERR_INVALID, "trajectory pose: values must be finite");
        ctx->camera.trajectory_pose(angle, distance, height, lookat ? lookat : zero);
        set_camera_from_orbit(ctx);
        return NMR_OK;
    });
}
NMR_API int nmr_get_camera(nmr_ctx* ctx, float out12[12]) {
    return guarded(ctx, [&]() -> int { std::memcpy(out12, ctx->cam12, sizeof(ctx->cam12)); return NMR_OK; });
}
NMR_API int nmr_set_camera(nmr_ctx* ctx, const float in12[12]) {
    return guarded(ctx, [&]() -> int {
        if (!in12 || !all_finite(in12, 12)) return fail(ctx, NMR_ERR_INVALID, "camera: twelve finite values");
        std::memcpy(ctx->cam12, in12, sizeof(ctx->cam12));
        for (int k = 0; k < 3; ++k) ctx->camera.eye[k] = in12[9 + k];   // keeps orbit() and the mesh stage coherent with the new pose
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}

NMR_API int nmr_set_shard(nmr_ctx* ctx, int rank, int world, int band) {
    return guarded(ctx, [&]() -> int {
        if (world < 1 || rank < 0 || rank >= world || band < 1) return fail(ctx, NMR_ERR_INVALID, "bad shard specification");
        ctx->shard_rank = rank; ctx->shard_world = world; ctx->shard_band = band; ctx->surf.spp = 0;
        return NMR_OK;
    });
}

NMR_API int nmr_set_lens(nmr_ctx* ctx, int enabled, float ior, float transmission, const float tint[3]) {
    return guarded(ctx, [&]() -> int {
        ctx->lens_enabled = enabled ? 1 : 0;
        if (ior > 1.f || transmission >= 0.f || tint) {
            if (!ctx->lens_params_set) ctx->lens_params_set = true;
            if (ior > 1.f) ctx->lens_ior = ior;
            if (transmission >= 0.f) ctx->lens_transmission = transmission > 1.f ? 1.f : transmission;
            if (tint) std::memcpy(ctx->lens_tint, tint, 12);
        }
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}

NMR_API int nmr_set_lens_model(nmr_ctx* ctx, int model, float thickness) {
    return guarded(ctx, [&]() -> int {
        if (model < 0 || model > 1 || !(thickness >= 0.f)) return fail(ctx, NMR_ERR_INVALID, "lens model must be 0 (thin sheet) or 1 (plate) with a thickness >= 0");
        ctx->lens_model = model; ctx->lens_thickness = thickness; ctx->surf.spp = 0;
        return NMR_OK;
    });
}

NMR_API int nmr_set_overlap(nmr_ctx* ctx, int enabled) {
    return guarded(ctx, [&]() -> int { ctx->overlap = enabled ? 1 : 0; return NMR_OK; });
}

NMR_API int nmr_set_surface_insertion(nmr_ctx* ctx, int mode) {
    return guarded(ctx, [&]() -> int {
        if (mode < NMR_SURFACE_AUTO || mode > NMR_SURFACE_BATCH8) return fail(ctx, NMR_ERR_INVALID, "bad surface insertion mode");
        ctx->surface_mode = mode; ctx->surf.spp = 0;
        return NMR_OK;
    });
}

namespace {
// A frame whose pixels go to a shared image (nmr_gather_*).  Device-side protocol, no host or NCCL step:
//   other ranks wait until the destination is done with the previous frame, render (their stores cross NVLink), signal;
//   the destination marks the previous frame consumed, renders its own rows, signals, then waits for every rank's signal -
//   so once the destination's stream has drained, the shared image holds the whole frame.
void enqueue_gather_frame(nmr_ctx* ctx, Nerf& n, const FrameParams& P0) {
    const uint32_t seq = ++ctx->gather_seq;
    uint32_t* f = ctx->gather_flags;
    // The destination's NVLink ingress is what a shared frame costs (a 4K float4 frame is 133 MB).  Pixels outside both screen
    // rectangles are constant background and the rectangles are the same on every rank, so the destination fills those itself,
    // for all rows, and nobody sends them.
    FrameParams P = P0;
    P.bg_filled_elsewhere = 1;
    if (ctx->gather_is_dst) {
        launch_gather_signal(f + kGatherConsumed, seq - 1u, ctx->stream);
        launch_fill_background(P, ctx->frame_target, ctx->stream);
    } else {
        launch_gather_wait(f, kGatherConsumed, 1, seq - 1u, f + kGatherError, ctx->stream);
    }
    enqueue_pass(ctx, n, P, true, ctx->frame_target);
    launch_gather_signal(f + ctx->shard_rank, seq, ctx->stream);
    if (ctx->gather_is_dst) launch_gather_wait(f, 0, ctx->shard_world, seq, f + kGatherError, ctx->stream);
    CK(cudaGetLastError());
}

}  // namespace

NMR_API int nmr_frame(nmr_ctx* ctx, int* keep_running) {
    return guarded(ctx, [&]() -> int {
        if (keep_running) *keep_running = 1;
        if (ctx->nerfs.empty()) return NMR_OK;   // the reference draws an empty window
        upload_mesh_if_dirty(ctx);
        Nerf& n = *ctx->nerfs[0];
        ctx->surf.resize(ctx->width, ctx->height, ctx->mesh_scale);
        const FrameParams P = make_params(ctx, n, ctx->width, ctx->height, ctx->cam12, ctx->surf.spp, true, true);
        ctx->surf.comb_valid = false;
        if (ctx->nerfs.size() > 1 && !ctx->frame_target && ctx->shard_world == 1) enqueue_multi_nerf_frame(ctx);
        else if (ctx->frame_target) enqueue_gather_frame(ctx, n, P); else enqueue_pass(ctx, n, P, true);
        ctx->surf.image_is_frame = true;
        ++ctx->surf.spp;
        CK(cudaStreamSynchronize(ctx->stream));   // frame() returns a finished frame (S/nerf_mesh_renderer.cu:578)
        if (ctx->frame_target && ctx->gather_is_dst) {
            // a wait of the shared-frame protocol that ran out of time leaves an incomplete image: fail this frame, not the detach
            uint32_t err = 0;
            CK(cudaMemcpy(&err, ctx->gather_flags + kGatherError, sizeof(err), cudaMemcpyDeviceToHost));
            if (err) return fail(ctx, NMR_ERR_STATE, "shared frame target: a rank did not deliver its rows in time; the image is incomplete");
        }
        return NMR_OK;
    });
}

NMR_API int nmr_read_frame(nmr_ctx* ctx, float* out_rgba) {
    return guarded(ctx, [&]() -> int {
        if (!out_rgba) return fail(ctx, NMR_ERR_INVALID, "out_rgba is null");
        if (!ctx->surf.image.p || ctx->surf.w == 0) return fail(ctx, NMR_ERR_STATE, "nothing rendered yet");
        // the caller's buffer holds width * height float4 of the CONSTRUCTOR resolution: only the image of a frame() may go there
        // (render() / render_views() resize the surfaces to their own resolution and format)
        if (!ctx->surf.image_is_frame || ctx->surf.w != ctx->width || ctx->surf.h != ctx->height || ctx->surf.image_format != kPixelF32)
            return fail(ctx, NMR_ERR_STATE, "the last image on this context is not a frame(): call frame() before read_frame()");
        const float4* src = (ctx->gather_image.p && ctx->frame_target == ctx->gather_image.p) ? ctx->gather_image.p : ctx->surf.image.p;
        CK(cudaMemcpyAsync(out_rgba, src, (size_t)ctx->width * ctx->height * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return NMR_OK;
    });
}

NMR_API int nmr_read_combined(nmr_ctx* ctx, float* out_frame_rgba, float* out_depth) {
    return guarded(ctx, [&]() -> int {
        Surfaces& S = ctx->surf;
        if (!S.image_is_frame || S.w != ctx->width || S.h != ctx->height) return fail(ctx, NMR_ERR_STATE, "the last image on this context is not a frame()");
        const size_t n = (size_t)ctx->width * ctx->height;
        const float4* f = nullptr; const float* d = nullptr;
        if (S.comb_valid) { f = S.comb_frame.p; d = S.comb_depth.p; }
        else if ((ctx->debug_flags & kDebugKeepProbes) && S.frame.p && S.depth.p) { f = S.frame.p; d = S.depth.p; }     // one NeRF: its own buffers are the merged ones
        else return fail(ctx, NMR_ERR_STATE, "frame / depth buffers are only kept for frames of several NeRFs, or with debug flag 4 set before the frame");
        if (out_frame_rgba) CK(cudaMemcpyAsync(out_frame_rgba, f, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
        if (out_depth) CK(cudaMemcpyAsync(out_depth, d, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return NMR_OK;
    });
}

NMR_API int nmr_render_format(nmr_ctx* ctx, int nerf_id, int width, int height, int spp, int linear, int format, void* out_rgba) {
    return guarded(ctx, [&]() -> int {
        if (width <= 0 || height <= 0 || spp < 1 || !out_rgba) return fail(ctx, NMR_ERR_INVALID, "bad render arguments");
        if (format < NMR_PIXEL_F32 || format > NMR_PIXEL_U8) return fail(ctx, NMR_ERR_INVALID, "bad pixel format");
        const size_t bpp = pixel_bytes(format);
        Nerf* n;
        try { n = get_nerf(ctx, nerf_id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        ctx->host_images.erase(std::remove_if(ctx->host_images.begin(), ctx->host_images.end(), [&](const nmr_ctx::HostImage& h) { return h.p == out_rgba; }), ctx->host_images.end());
        upload_mesh_if_dirty(ctx);
        ctx->surf.resize(width, height, ctx->mesh_scale);
        ctx->surf.spp = 0;   // reset_accumulation (S/python_api.cu:85)
        // Testbed::render uses Testbed::m_camera, which frame()/orbit keep equal to viewProjectionMat; its aspect comes from the
        // renderer's constructor resolution, not from (width, height), exactly like the reference.
        static const bool no_bands = std::getenv("NMR_NO_BANDS") != nullptr;     // measurement aid: plain render + one copy
        if (spp == 1 && ctx->shard_world == 1 && height >= 256 && !no_bands) {
            // one sample per pixel: the rows known to be background leave while the frame is still being rendered
            FrameParams P = make_params(ctx, *n, width, height, ctx->cam12, 0, !linear, true);
            P.out_format = format;
            enqueue_pass_with_copy(ctx, *n, P, out_rgba);
            CK(cudaStreamSynchronize(ctx->copy_stream));
            CK(cudaStreamSynchronize(ctx->stream));
            ctx->surf.spp = 0;
            return NMR_OK;
        }
        for (int i = 0; i < spp; ++i) {
            FrameParams P = make_params(ctx, *n, width, height, ctx->cam12, ctx->surf.spp, !linear, true);
            P.out_format = format;
            enqueue_pass(ctx, *n, P, i == spp - 1);
            ++ctx->surf.spp;
        }
        CK(cudaMemcpyAsync(out_rgba, ctx->surf.image.p, (size_t)width * height * bpp, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->surf.spp = 0;   // the next frame() starts a fresh accumulation at its own resolution
        return NMR_OK;
    });
}
NMR_API int nmr_render_update(nmr_ctx* ctx, int nerf_id, int width, int height, int linear, int format, void* out_rgba, int assume_valid, size_t* out_copied_bytes) {
    return guarded(ctx, [&]() -> int {
        if (width <= 0 || height <= 0 || !out_rgba) return fail(ctx, NMR_ERR_INVALID, "bad render arguments");
        if (format < NMR_PIXEL_F32 || format > NMR_PIXEL_U8) return fail(ctx, NMR_ERR_INVALID, "bad pixel format");
        if (ctx->shard_world != 1) return fail(ctx, NMR_ERR_STATE, "nmr_render_update: not on a sharded context");
        Nerf* n;
        try { n = get_nerf(ctx, nerf_id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        upload_mesh_if_dirty(ctx);
        ctx->surf.resize(width, height, ctx->mesh_scale);
        ctx->surf.spp = 0;
        FrameParams P = make_params(ctx, *n, width, height, ctx->cam12, 0, !linear, true);
        P.out_format = format;
        auto it = std::find_if(ctx->host_images.begin(), ctx->host_images.end(), [&](const nmr_ctx::HostImage& h) { return h.p == out_rgba; });
        const bool known = assume_valid != 0 && it != ctx->host_images.end() && it->w == width && it->h == height && it->fmt == format && std::memcmp(it->bg, P.background_out, 16) == 0;
        nmr_ctx::HostImage rec{out_rgba, width, height, format, {P.background_out[0], P.background_out[1], P.background_out[2], P.background_out[3]}, {0, 0, 0, 0}};
        const int prev[4] = {known ? it->rect[0] : 0, known ? it->rect[1] : 0, known ? it->rect[2] : 0, known ? it->rect[3] : 0};
        if (it != ctx->host_images.end()) ctx->host_images.erase(it);        // (re-entered below once the copies are in the stream)
        size_t moved = 0;
        enqueue_pass_with_copy(ctx, *n, P, out_rgba, known ? prev : nullptr, rec.rect, &moved);
        CK(cudaStreamSynchronize(ctx->copy_stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->host_images.size() >= 8) ctx->host_images.erase(ctx->host_images.begin());
        ctx->host_images.push_back(rec);
        if (out_copied_bytes) *out_copied_bytes = moved;
        return NMR_OK;
    });
}
NMR_API int nmr_render(nmr_ctx* ctx, int nerf_id, int width, int height, int spp, int linear, float* out_rgba) {
    return nmr_render_format(ctx, nerf_id, width, height, spp, linear, NMR_PIXEL_F32, out_rgba);
}

NMR_API int nmr_render_views_format(nmr_ctx* ctx, int nerf_id, int n_views, const float* cams12, int width, int height, int linear, int format, void* out_rgba_v) {
    return guarded(ctx, [&]() -> int {
        if (width <= 0 || height <= 0 || n_views < 1 || !cams12) return fail(ctx, NMR_ERR_INVALID, "bad render_views arguments");
        for (int v = 0; v < n_views; ++v) if (!all_finite(cams12 + (size_t)v * 12, 12)) return fail(ctx, NMR_ERR_INVALID, "render_views: cameras must be finite");
        if (format < NMR_PIXEL_F32 || format > NMR_PIXEL_U8) return fail(ctx, NMR_ERR_INVALID, "bad pixel format");
        const size_t bpp = pixel_bytes(format);
        char* out_rgba = static_cast<char*>(out_rgba_v);
        Nerf* n;
        try { n = get_nerf(ctx, nerf_id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        upload_mesh_if_dirty(ctx);
        ctx->surf.resize(width, height, ctx->mesh_scale);
        const size_t px = (size_t)width * height;
        // lanes: 8 up to 1 Mpixel per view, 4 up to 1080p, 2 above (a lane holds ~130 bytes per pixel); NMR_VIEW_LANES overrides
        static const int n_lanes_env = [] { const char* v = std::getenv("NMR_VIEW_LANES"); return v ? std::atoi(v) : 0; }();
        const int n_lanes_auto = px <= ((size_t)1 << 20) ? 8 : (px <= (size_t)1920 * 1080 ? 4 : 2);
        const int K = std::max(1, std::min(std::min(n_lanes_env > 0 ? n_lanes_env : n_lanes_auto, 8), n_views));
        if (K > 1 && ctx->shard_world == 1 && !(ctx->debug_flags & kDebugKeepProbes)) {      // (parity probes read the context's own surfaces)
            // K views in flight: lane v % K renders view v on its own stream with a march grid of one CTA per SM (K marches fill
            // the SMs together, and one view's tail overlaps the others' set-up); a view leaves over PCIe on the copy stream as
            // soon as it is complete, and a lane renders its next view once its image has been copied out.
            while ((int)ctx->lanes.size() < K) ctx->lanes.push_back(make_lane(ctx));
            CK(cudaEventRecord(ctx->ev_view[0][0], ctx->stream));          // everything enqueued on the context's stream so far comes first
            for (int k = 0; k < K; ++k) {
                nmr_ctx* l = ctx->lanes[(size_t)k];
                sync_lane(l, ctx);
                l->march_ctas = K >= 3 ? 1 : 2;
                l->surf.resize(width, height, ctx->mesh_scale);
                CK(cudaStreamWaitEvent(l->stream, ctx->ev_view[0][0], 0));
            }
            for (int v = 0; v < n_views; ++v) {
                nmr_ctx* l = ctx->lanes[(size_t)(v % K)];
                if (v >= K && out_rgba) CK(cudaStreamWaitEvent(l->stream, l->ev_view[0][1], 0));      // the lane's previous image has been copied out
                FrameParams P = make_params(ctx, *n, width, height, cams12 + (size_t)v * 12, 0, !linear, true);
                P.out_format = format;
                enqueue_pass(l, *n, P, v == n_views - 1);
                if (!out_rgba) continue;                                                   // images stay on the device
                CK(cudaEventRecord(l->ev_view[0][0], l->stream));
                CK(cudaStreamWaitEvent(ctx->copy_stream, l->ev_view[0][0], 0));
                CK(cudaMemcpyAsync(out_rgba + (size_t)v * px * bpp, l->surf.image.p, px * bpp, cudaMemcpyDefault, ctx->copy_stream));     // (host or device destination)
                CK(cudaEventRecord(l->ev_view[0][1], ctx->copy_stream));
            }
            nmr_ctx* last = ctx->lanes[(size_t)((n_views - 1) % K)];
            // nmr_get_device_image / nmr_get_stats after the call refer to the last view
            CK(cudaMemcpyAsync(ctx->surf.image.p, last->surf.image.p, px * bpp, cudaMemcpyDeviceToDevice, last->stream));
            ctx->surf.image_format = format; ctx->surf.image_is_frame = false;
            CK(cudaStreamSynchronize(ctx->copy_stream));
            for (int k = 0; k < K; ++k) CK(cudaStreamSynchronize(ctx->lanes[(size_t)k]->stream));
            finish_stats(last);
            ctx->stats = last->stats; ctx->stats_pending = false;
            ctx->surf.spp = 0;
            return NMR_OK;
        }
        // one view at a time (sharded contexts, NMR_VIEW_LANES=1): two image buffers, two streams - view v renders into buffer
        // v % 2 while view v - 1 leaves the other one over PCIe.  (The target is handed to enqueue_pass explicitly: the two
        // DevBufs keep owning their own allocation whatever happens in between.)
        Surfaces& S = ctx->surf;
        S.image_alt.ensure(px);
        float4* bufs[2] = {S.image.p, S.image_alt.p};
        for (int v = 0; v < n_views; ++v) {
            const int b = v & 1;
            if (v >= 2) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_view[b][1], 0));      // buffer b has been copied out
            FrameParams P = make_params(ctx, *n, width, height, cams12 + (size_t)v * 12, 0, !linear, true);
            P.out_format = format;
            enqueue_pass(ctx, *n, P, v == n_views - 1, bufs[b]);
            CK(cudaEventRecord(ctx->ev_view[b][0], ctx->stream));
            if (!out_rgba) { CK(cudaEventRecord(ctx->ev_view[b][1], ctx->stream)); continue; }
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_view[b][0], 0));
            CK(cudaMemcpyAsync(out_rgba + (size_t)v * px * bpp, bufs[b], px * bpp, cudaMemcpyDefault, ctx->copy_stream));
            CK(cudaEventRecord(ctx->ev_view[b][1], ctx->copy_stream));
        }
        if (((n_views - 1) & 1) == 1) { std::swap(S.image.p, S.image_alt.p); std::swap(S.image.n, S.image_alt.n); }   // nmr_get_device_image: the last view's buffer
        S.image_format = format; S.image_is_frame = false;
        CK(cudaStreamSynchronize(ctx->copy_stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}
NMR_API int nmr_render_views(nmr_ctx* ctx, int nerf_id, int n_views, const float* cams12, int width, int height, int linear, float* out_rgba) {
    return nmr_render_views_format(ctx, nerf_id, n_views, cams12, width, height, linear, NMR_PIXEL_F32, out_rgba);
}

NMR_API int nmr_get_device_image(nmr_ctx* ctx, void** out_dev_ptr, int* out_w, int* out_h) {
    return guarded(ctx, [&]() -> int {
        if (!ctx->surf.image.p) return fail(ctx, NMR_ERR_STATE, "nothing rendered yet");
        if (out_dev_ptr) *out_dev_ptr = ctx->surf.image.p;
        if (out_w) *out_w = ctx->surf.w;
        if (out_h) *out_h = ctx->surf.h;
        return NMR_OK;
    });
}

NMR_API int nmr_copy_device_image(nmr_ctx* ctx, void* dst, size_t dst_bytes) {
    return guarded(ctx, [&]() -> int {
        if (!ctx->surf.image.p || !dst) return fail(ctx, NMR_ERR_STATE, "nothing rendered yet");
        const size_t bytes = (size_t)ctx->surf.w * ctx->surf.h * pixel_bytes(ctx->surf.image_format);
        if (dst_bytes < bytes) return fail(ctx, NMR_ERR_INVALID, "destination holds " + std::to_string(dst_bytes) + " bytes, the last image has " + std::to_string(bytes));
        CK(cudaMemcpyAsync(dst, ctx->surf.image.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return NMR_OK;
    });
}

NMR_API int nmr_flush_l2(nmr_ctx* ctx) {
    return guarded(ctx, [&]() -> int {
        const size_t bytes = (size_t)256 << 20;
        ctx->d_flush.ensure(bytes);
        CK(cudaMemsetAsync(ctx->d_flush.p, 0x5A, bytes, ctx->stream));
        return NMR_OK;
    });
}

NMR_API int nmr_measure_l2(nmr_ctx* ctx, size_t bytes, int mode, float* out_gbs) {
    return guarded(ctx, [&]() -> int {
        if (!out_gbs || (mode != 0 && mode != 1)) return fail(ctx, NMR_ERR_INVALID, "bad arguments");
        size_t pow2 = (size_t)1 << 24;
        while (pow2 * 2 <= bytes && pow2 < ((size_t)1 << 30)) pow2 *= 2;       // power-of-two size (the gather mode masks its indices)
        DevBuf<uint8_t> buf; buf.ensure(pow2);
        CK(cudaMemsetAsync(buf.p, 0x11, pow2, ctx->stream));
        const uint32_t n_vec = (uint32_t)(pow2 / 16), loads = mode == 0 ? 512u : 2048u;
        uint32_t* sink = reinterpret_cast<uint32_t*>(ctx->d_scratch.p);          // never written in practice (the kernel's condition does not hold)
        launch_l2_probe(buf.p, n_vec, loads, mode, sink, ctx->num_sms, ctx->stream);      // warm: the buffer is L2-resident afterwards
        float best = 0.f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(ctx->ev[0], ctx->stream));
            launch_l2_probe(buf.p, n_vec, loads, mode, sink, ctx->num_sms, ctx->stream);
            CK(cudaEventRecord(ctx->ev[1], ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
            const double total = (double)ctx->num_sms * 8 * 256 * loads * (mode == 0 ? 16.0 : 4.0);
            if (ms > 0.f) best = std::max(best, (float)(total / (ms * 1e-3) / 1e9));
        }
        CK(cudaGetLastError());
        *out_gbs = best;
        return NMR_OK;
    });
}

NMR_API int nmr_get_stats(nmr_ctx* ctx, nmr_stats* out) {
    return guarded(ctx, [&]() -> int { if (!out) return fail(ctx, NMR_ERR_INVALID, "out is null"); finish_stats(ctx); *out = ctx->stats; return NMR_OK; });
}
NMR_API int nmr_synchronize(nmr_ctx* ctx) {
    return guarded(ctx, [&]() -> int { CK(cudaStreamSynchronize(ctx->stream)); return NMR_OK; });
}

NMR_API int nmr_get_stream(nmr_ctx* ctx, void** out_stream) {
    return guarded(ctx, [&]() -> int {
        if (!out_stream) return fail(ctx, NMR_ERR_INVALID, "out_stream is null");
        *out_stream = static_cast<void*>(ctx->stream);
        return NMR_OK;
    });
}

NMR_API int nmr_gather_create(nmr_ctx* ctx, uint8_t handle64[64], void** out_dev_ptr) {
    return guarded(ctx, [&]() -> int {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
        if (!handle64) return fail(ctx, NMR_ERR_INVALID, "handle is null");
        if (ctx->shard_world > kGatherMaxRanks) return fail(ctx, NMR_ERR_INVALID, "too many ranks for a shared frame target");
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->ipc_mapped) { CK(cudaIpcCloseMemHandle(ctx->ipc_mapped)); ctx->ipc_mapped = nullptr; }
        const size_t px = (size_t)ctx->width * ctx->height;
        ctx->gather_image.ensure(px + kGatherFlagWords * sizeof(uint32_t) / sizeof(float4));     // image, then the sequence flags
        ctx->gather_flags = reinterpret_cast<uint32_t*>(ctx->gather_image.p + px);
        CK(cudaMemset(ctx->gather_flags, 0, kGatherFlagWords * sizeof(uint32_t)));
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ctx->gather_image.p));
        std::memcpy(handle64, &h, 64);
        ctx->frame_target = ctx->gather_image.p; ctx->gather_is_dst = true; ctx->gather_seq = 0;
        if (out_dev_ptr) *out_dev_ptr = ctx->gather_image.p;
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}
NMR_API int nmr_gather_attach(nmr_ctx* ctx, const uint8_t handle64[64]) {
    return guarded(ctx, [&]() -> int {
        if (!handle64) return fail(ctx, NMR_ERR_INVALID, "handle is null");
        if (ctx->shard_rank >= kGatherMaxRanks) return fail(ctx, NMR_ERR_INVALID, "too many ranks for a shared frame target");
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->ipc_mapped) { CK(cudaIpcCloseMemHandle(ctx->ipc_mapped)); ctx->ipc_mapped = nullptr; ctx->frame_target = nullptr; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->ipc_mapped = p; ctx->frame_target = static_cast<float4*>(p);
        ctx->gather_flags = reinterpret_cast<uint32_t*>(ctx->frame_target + (size_t)ctx->width * ctx->height);
        ctx->gather_is_dst = false; ctx->gather_seq = 0;
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}
NMR_API int nmr_gather_detach(nmr_ctx* ctx) {
    return guarded(ctx, [&]() -> int {
        CK(cudaStreamSynchronize(ctx->stream));
        uint32_t err = 0;
        if (ctx->gather_flags) CK(cudaMemcpy(&err, ctx->gather_flags + kGatherError, sizeof(err), cudaMemcpyDeviceToHost));
        if (ctx->ipc_mapped) { CK(cudaIpcCloseMemHandle(ctx->ipc_mapped)); ctx->ipc_mapped = nullptr; }
        ctx->frame_target = nullptr; ctx->gather_flags = nullptr; ctx->gather_is_dst = false; ctx->gather_seq = 0;
        ctx->surf.spp = 0;
        if (err) return fail(ctx, NMR_ERR_STATE, "a rank of the shared frame target did not render a frame the others waited for");
        return NMR_OK;
    });
}

// Asynchronous frame for throughput measurements and pipelined callers: same work as nmr_frame without the trailing
// synchronisation; call nmr_synchronize() / nmr_get_stats() / nmr_read_frame() to wait.
NMR_API int nmr_frame_async(nmr_ctx* ctx) {
    return guarded(ctx, [&]() -> int {
        if (ctx->nerfs.empty()) return fail(ctx, NMR_ERR_STATE, "no NeRF loaded");
        upload_mesh_if_dirty(ctx);
        Nerf& n = *ctx->nerfs[0];
        ctx->surf.resize(ctx->width, ctx->height, ctx->mesh_scale);
        const FrameParams P = make_params(ctx, n, ctx->width, ctx->height, ctx->cam12, ctx->surf.spp, true, true);
        ctx->surf.comb_valid = false;
        if (ctx->nerfs.size() > 1 && !ctx->frame_target && ctx->shard_world == 1) enqueue_multi_nerf_frame(ctx);
        else if (ctx->frame_target) enqueue_gather_frame(ctx, n, P); else enqueue_pass(ctx, n, P, true);
        ctx->surf.image_is_frame = true;
        ++ctx->surf.spp;
        return NMR_OK;
    });
}

NMR_API int nmr_get_density_bitfield(nmr_ctx* ctx, int id, uint8_t* out) {
    return guarded(ctx, [&]() -> int {
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        CK(cudaMemcpyAsync(out, n->d_bitfield.p, kBitfieldBytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return NMR_OK;
    });
}
NMR_API int nmr_set_density_bitfield(nmr_ctx* ctx, int id, const uint8_t* in) {
    return guarded(ctx, [&]() -> int {
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        CK(cudaMemcpyAsync(n->d_bitfield.p, in, kBitfieldBytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        update_occupied_box(ctx, *n);
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}

// NerfMeshRenderer::dumpDensityGrid / loadDensityGrid (S/nerf_mesh_renderer.cu:239-358): the occupancy bitfield as 8 x 128^3 bytes
// (0 / 1), cell (x, y, z) of cascade `mip` at x + 128 y + 128^2 z + 128^3 mip; the bit of that cell is the Morton code of
// (x, y, z) inside the cascade (pos_to_cascaded_grid_idx of the cell's centre-aligned position, which maps back to (x, y, z)).
namespace {
uint32_t host_expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
}  // namespace
NMR_API int nmr_dump_density_grid(nmr_ctx* ctx, int id, const char* path, uint8_t* out_cells) {
    return guarded(ctx, [&]() -> int {
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        if (!path && !out_cells) return fail(ctx, NMR_ERR_INVALID, "neither a path nor a buffer given");
        std::vector<uint8_t> bits(kBitfieldBytes), cells((size_t)kCascades * kGridCells);
        CK(cudaMemcpyAsync(bits.data(), n->d_bitfield.p, kBitfieldBytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (uint32_t mip = 0; mip < kCascades; ++mip)
            for (uint32_t z = 0; z < kGridSize; ++z) for (uint32_t y = 0; y < kGridSize; ++y) for (uint32_t x = 0; x < kGridSize; ++x) {
                const uint32_t idx = host_expand_bits(x) | (host_expand_bits(y) << 1) | (host_expand_bits(z) << 2);
                cells[x + kGridSize * (y + kGridSize * (z + (size_t)kGridSize * mip))] = (bits[idx / 8 + (size_t)mip * kGridCells / 8] >> (idx % 8)) & 1u;
            }
        if (out_cells) std::memcpy(out_cells, cells.data(), cells.size());
        if (path) {
            FILE* f = std::fopen(path, "wb");
            if (!f) return fail(ctx, NMR_ERR_IO, std::string("cannot open ") + path);
            const size_t w = std::fwrite(cells.data(), 1, cells.size(), f);
            std::fclose(f);
            if (w != cells.size()) return fail(ctx, NMR_ERR_IO, std::string("short write to ") + path);
        }
        return NMR_OK;
    });
}
NMR_API int nmr_load_density_grid(nmr_ctx* ctx, int id, const char* path, const uint8_t* in_cells) {
    return guarded(ctx, [&]() -> int {
        Nerf* n; try { n = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        if (!path && !in_cells) return fail(ctx, NMR_ERR_INVALID, "neither a path nor a buffer given");
        std::vector<uint8_t> cells;
        if (path) {
            FILE* f = std::fopen(path, "rb");
            if (!f) return fail(ctx, NMR_ERR_IO, std::string("cannot open ") + path);
            cells.resize((size_t)kCascades * kGridCells);
            const size_t r = std::fread(cells.data(), 1, cells.size(), f);
            std::fclose(f);
            if (r != cells.size()) return fail(ctx, NMR_ERR_FORMAT, "density grid file must hold 8 x 128^3 bytes");
            in_cells = cells.data();
        }
        std::vector<uint8_t> bits(kBitfieldBytes, 0);
        for (uint32_t mip = 0; mip < kCascades; ++mip)
            for (uint32_t z = 0; z < kGridSize; ++z) for (uint32_t y = 0; y < kGridSize; ++y) for (uint32_t x = 0; x < kGridSize; ++x) {
                if (!in_cells[x + kGridSize * (y + kGridSize * (z + (size_t)kGridSize * mip))]) continue;
                const uint32_t idx = host_expand_bits(x) | (host_expand_bits(y) << 1) | (host_expand_bits(z) << 2);
                bits[idx / 8 + (size_t)mip * kGridCells / 8] |= (uint8_t)(1u << (idx % 8));
            }
        CK(cudaMemcpyAsync(n->d_bitfield.p, bits.data(), kBitfieldBytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        update_occupied_box(ctx, *n);
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}

NMR_API int nmr_remove_floaties(nmr_ctx* ctx, int* out_clusters, int64_t* out_kept) {
    return guarded(ctx, [&]() -> int {
        if (ctx->nerfs.empty()) return fail(ctx, NMR_ERR_STATE, "no NeRF loaded");
        Nerf& n = *ctx->nerfs.back();   // _nerfs.back(), S/nerf_mesh_renderer.cu:248
        DevBuf<uint32_t> labels; labels.ensure(floaties_label_bytes() / 4);
        DevBuf<unsigned long long> scratch; scratch.ensure(floaties_scratch_bytes() / 8);
        launch_remove_floaties(n.d_bitfield.p, n.host.max_cascade, labels.p, scratch.p, ctx->stream);
        unsigned long long res[4] = {0, 0, 0, 0};
        CK(cudaMemcpyAsync(res, scratch.p, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        update_occupied_box(ctx, n);
        if (out_clusters) *out_clusters = (int)res[0];
        if (out_kept) *out_kept = (int64_t)res[1];
        ctx->surf.spp = 0;
        return NMR_OK;
    });
}

// ---- parity probes ------------------------------------------------------------------------------------------------
namespace {
int run_probe(nmr_ctx* ctx, int id, int mode, int64_t n, const float* points_world, const float* direction, float* out) {
    return guarded(ctx, [&]() -> int {
        Nerf* nf; try { nf = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        if (n < 0 || (n > 0 && (!points_world || !out)) || !direction) return fail(ctx, NMR_ERR_INVALID, "null argument");
        if (n == 0) return NMR_OK;
        DevBuf<float> d_pts, d_out;
        d_pts.ensure((size_t)n * 3); d_out.ensure((size_t)n);
        CK(cudaMemcpyAsync(d_pts.p, points_world, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
        const FrameParams P = make_params(ctx, *nf, ctx->width, ctx->height, ctx->cam12, 0, true, false);
        launch_probe(P, model_for(ctx, *nf), d_pts.p, direction, n, mode, d_out.p, ctx->debug_flags, ctx->num_sms, ctx->stream);
        CK(cudaMemcpyAsync(out, d_out.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        return NMR_OK;
    });
}
}  // namespace

NMR_API int nmr_probe_points(nmr_ctx* ctx, int nerf_id, int64_t n, const float* points_world, const float direction[3], float* out_alpha) {
    return run_probe(ctx, nerf_id, 0, n, points_world, direction, out_alpha);
}
NMR_API int nmr_probe_rays(nmr_ctx* ctx, int nerf_id, int64_t n, const float* origins_world, const float direction[3], float* out_distance) {
    return run_probe(ctx, nerf_id, 1, n, origins_world, direction, out_distance);
}

NMR_API int nmr_debug_encode(nmr_ctx* ctx, int id, const float* pos, int64_t n, uint16_t* out) {
    return guarded(ctx, [&]() -> int {
        Nerf* nf; try { nf = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        if (n <= 0) return NMR_OK;
        DevBuf<float> d_pos; DevBuf<uint16_t> d_out;
        d_pos.ensure((size_t)n * 3); d_out.ensure((size_t)n * ENC_WIDTH);
        CK(cudaMemcpyAsync(d_pos.p, pos, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
        launch_debug_encode(model_for(ctx, *nf), d_pos.p, n, d_out.p, ctx->stream);
        CK(cudaMemcpyAsync(out, d_out.p, (size_t)n * ENC_WIDTH * 2, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        return NMR_OK;
    });
}

NMR_API int nmr_debug_network(nmr_ctx* ctx, int id, const float* pos, const float* dir, int64_t n, uint16_t* out4) {
    return guarded(ctx, [&]() -> int {
        Nerf* nf; try { nf = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        if (n <= 0) return NMR_OK;
        DevBuf<float> d_pos, d_dir; DevBuf<uint16_t> d_out;
        d_pos.ensure((size_t)n * 3); d_dir.ensure((size_t)n * 3); d_out.ensure((size_t)n * 4);
        CK(cudaMemcpyAsync(d_pos.p, pos, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_dir.p, dir, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
        launch_debug_network(model_for(ctx, *nf), d_pos.p, d_dir.p, n, d_out.p, ctx->debug_flags, ctx->stream);
        CK(cudaMemcpyAsync(out4, d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        return NMR_OK;
    });
}

NMR_API int nmr_debug_trace(nmr_ctx* ctx, int id, int width, int height, const uint32_t* pixels, int64_t n_pix, uint32_t max_samples,
                            float* o_t, uint32_t* o_cell, uint32_t* o_mip, float* o_pos, uint32_t* o_count, float* o_ray) {
    return guarded(ctx, [&]() -> int {
        Nerf* nf; try { nf = get_nerf(ctx, id); } catch (const std::invalid_argument& e) { return fail(ctx, NMR_ERR_INVALID, e.what()); }
        if (n_pix <= 0 || max_samples == 0) return NMR_OK;
        const size_t ns = (size_t)n_pix * max_samples;
        DevBuf<uint32_t> d_pix, d_cell, d_mip, d_cnt; DevBuf<float> d_t, d_pos, d_ray;
        d_pix.ensure((size_t)n_pix); d_cell.ensure(ns); d_mip.ensure(ns); d_cnt.ensure((size_t)n_pix); d_t.ensure(ns); d_pos.ensure(ns * 3); d_ray.ensure((size_t)n_pix * 8);
        CK(cudaMemcpyAsync(d_pix.p, pixels, (size_t)n_pix * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(d_t.p, 0, ns * 4, ctx->stream)); CK(cudaMemsetAsync(d_cell.p, 0, ns * 4, ctx->stream));
        CK(cudaMemsetAsync(d_mip.p, 0, ns * 4, ctx->stream)); CK(cudaMemsetAsync(d_pos.p, 0, ns * 12, ctx->stream));
        const FrameParams P = make_params(ctx, *nf, width, height, ctx->cam12, 0, true, false);
        launch_debug_trace(P, nf->dev, d_pix.p, n_pix, max_samples, d_t.p, d_cell.p, d_mip.p, d_pos.p, d_cnt.p, d_ray.p, ctx->stream);
        CK(cudaMemcpyAsync(o_t, d_t.p, ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(o_cell, d_cell.p, ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(o_mip, d_mip.p, ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(o_pos, d_pos.p, ns * 12, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(o_count, d_cnt.p, (size_t)n_pix * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(o_ray, d_ray.p, (size_t)n_pix * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        return NMR_OK;
    });
}

// shared by the mesh probes: frame parameters of the mesh stage alone + the visibility buffer of the current camera
static void debug_mesh_stage(nmr_ctx* ctx, int width, int height, FrameParams& P, MeshDevice& mesh, DevBuf<unsigned long long>& zbuf) {
    const int ms = ctx->mesh_scale;
    P = FrameParams{};
    P.width = width; P.height = height; std::memcpy(P.cam, ctx->cam12, sizeof(P.cam));
    P.shard_world = 1; P.shard_band = 8; P.mesh_scale = ms; std::memcpy(P.light, ctx->light, 12);
    P.model_rot[0] = P.model_rot[4] = P.model_rot[8] = 1.f;
    invert3(ctx->cam12, P.cam_inv);
    mesh_screen_box(ctx, P);
    P.lens_on = (ctx->scene_has_lens && ctx->lens_enabled) ? 1 : 0;
    mesh = ctx->mesh_dev;
    if (!P.lens_on) mesh.tri_lens = nullptr;
    zbuf.ensure((size_t)width * ms * height * ms * 2);
    launch_mesh_raster(mesh, P, height, zbuf.p, ctx->stream);
}

NMR_API int nmr_debug_mesh(nmr_ctx* ctx, int width, int height, float* o_rgba2, float* o_depth2, int32_t* o_tri2, float* o_surf, float* o_tsurf) {
    return guarded(ctx, [&]() -> int {
        upload_mesh_if_dirty(ctx);
        if (ctx->mesh_dev.n_tris == 0) return fail(ctx, NMR_ERR_STATE, "no mesh loaded");
        const int ms = ctx->mesh_scale;
        const size_t n2 = (size_t)width * ms * height * ms, n1 = (size_t)width * height;
        DevBuf<unsigned long long> zbuf;
        DevBuf<float> d_rgba2, d_depth2, d_surf, d_ts; DevBuf<int32_t> d_tri2;
        d_rgba2.ensure(n2 * 4); d_depth2.ensure(n2); d_tri2.ensure(n2); d_surf.ensure(n1 * 4); d_ts.ensure(n1);
        FrameParams P; MeshDevice mesh;
        debug_mesh_stage(ctx, width, height, P, mesh, zbuf);
        launch_debug_mesh(mesh, P, zbuf.p, d_rgba2.p, d_depth2.p, d_tri2.p, d_surf.p, d_ts.p, nullptr, ctx->stream);
        if (o_rgba2) CK(cudaMemcpyAsync(o_rgba2, d_rgba2.p, n2 * 16, cudaMemcpyDeviceToHost, ctx->stream));
        if (o_depth2) CK(cudaMemcpyAsync(o_depth2, d_depth2.p, n2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (o_tri2) CK(cudaMemcpyAsync(o_tri2, d_tri2.p, n2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (o_surf) CK(cudaMemcpyAsync(o_surf, d_surf.p, n1 * 16, cudaMemcpyDeviceToHost, ctx->stream));
        if (o_tsurf) CK(cudaMemcpyAsync(o_tsurf, d_ts.p, n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        return NMR_OK;
    });
}

NMR_API int nmr_debug_lens(nmr_ctx* ctx, int width, int height, float* o_w, float* o_t, float* o_normal) {
    return guarded(ctx, [&]() -> int {
        upload_mesh_if_dirty(ctx);
        if (ctx->mesh_dev.n_tris == 0) return fail(ctx, NMR_ERR_STATE, "no mesh loaded");
        const size_t n1 = (size_t)width * height;
        DevBuf<unsigned long long> zbuf;
        DevBuf<float> d_lens; d_lens.ensure(n1 * 5);
        FrameParams P; MeshDevice mesh;
        debug_mesh_stage(ctx, width, height, P, mesh, zbuf);
        launch_debug_mesh(mesh, P, zbuf.p, nullptr, nullptr, nullptr, nullptr, nullptr, d_lens.p, ctx->stream);
        std::vector<float> h(n1 * 5);
        CK(cudaMemcpyAsync(h.data(), d_lens.p, n1 * 20, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        for (size_t i = 0; i < n1; ++i) {
            if (o_w) o_w[i] = h[i * 5]; if (o_t) o_t[i] = h[i * 5 + 1];
            if (o_normal) { o_normal[i * 3] = h[i * 5 + 2]; o_normal[i * 3 + 1] = h[i * 5 + 3]; o_normal[i * 3 + 2] = h[i * 5 + 4]; }
        }
        return NMR_OK;
    });
}

NMR_API int nmr_debug_last_frame(nmr_ctx* ctx, float* o_frame, float* o_depth, uint32_t* o_ns) {
    return guarded(ctx, [&]() -> int {
        if (!ctx->surf.frame.p || ctx->surf.w == 0) return fail(ctx, NMR_ERR_STATE, "nothing rendered yet");
        const size_t n = (size_t)ctx->surf.w * ctx->surf.h;
        if (o_frame) CK(cudaMemcpyAsync(o_frame, ctx->surf.frame.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
        if (o_depth) CK(cudaMemcpyAsync(o_depth, ctx->surf.depth.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (o_ns) CK(cudaMemcpyAsync(o_ns, ctx->surf.n_samples.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        return NMR_OK;
    });
}

NMR_API int nmr_debug_parse_gltf(const char* path, int64_t out_counts[5], char* err, size_t err_len, float* out_tangents, size_t tangent_capacity) {
    auto put = [&](const std::string& m) { if (err && err_len) { std::snprintf(err, err_len, "%s", m.c_str()); } };
    put("");
    if (!path) { put("path is null"); return NMR_ERR_INVALID; }
    try {
        const HostMesh m = load_gltf(path);
        if (out_counts) {
            int64_t lens = 0; for (uint8_t f : m.tri_lens) lens += f ? 1 : 0;
            out_counts[0] = (int64_t)(m.positions.size() / 3); out_counts[1] = (int64_t)(m.indices.size() / 3); out_counts[2] = lens;
            out_counts[3] = m.tex_w; out_counts[4] = m.tex_h;
        }
        if (out_tangents) std::memcpy(out_tangents, m.tangents.data(), std::min(tangent_capacity, m.tangents.size()) * sizeof(float));
        // what nmr_load_mesh does next on the host: the arrays transform_mesh walks must cover every vertex
        std::vector<float> wp, wn;
        const float t[3] = {0.f, 0.f, 0.f}, sc[3] = {1.f, 1.f, 1.f}, q[4] = {0.f, 0.f, 0.f, 1.f};
        transform_mesh(m, t, sc, q, wp, wn);
        if (!m.warning.empty()) put(m.warning);
        return NMR_OK;
    } catch (const std::bad_alloc&) {
        put("out of host memory"); return NMR_ERR_INVALID;
    } catch (const std::exception& e) {
        const std::string msg = e.what();
        put(msg);
        return msg.rfind("cannot open", 0) == 0 ? NMR_ERR_IO : NMR_ERR_FORMAT;
    }
}

NMR_API int nmr_mikk_tangents(const float* positions, const float* normals, const float* texcoords, int64_t n_vertices,
                              const uint32_t* indices, int64_t n_indices, float* out_tangents) {
    if (!positions || !normals || !texcoords || !indices || !out_tangents || n_vertices <= 0 || n_indices < 0) return NMR_ERR_INVALID;
    for (int64_t i = 0; i < n_indices; ++i) if ((int64_t)indices[i] >= n_vertices) return NMR_ERR_INVALID;
    try {
        for (int64_t v = 0; v < n_vertices; ++v) { float* o = out_tangents + v * 4; o[0] = 1.f; o[1] = 0.f; o[2] = 0.f; o[3] = -1.f; }
        mikk_tangents(positions, normals, texcoords, (size_t)n_vertices, indices, (size_t)n_indices, out_tangents);
        return NMR_OK;
    } catch (const std::exception&) { return NMR_ERR_INVALID; }
}

// debug knob: bit 0 CUDA-core MLP, bit 1 swap UMMA descriptor offsets
NMR_API int nmr_debug_set_flags(nmr_ctx* ctx, uint32_t flags) {
    return guarded(ctx, [&]() -> int { ctx->debug_flags = flags; return NMR_OK; });
}

}  // extern "C"
