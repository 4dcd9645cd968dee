// host.cpp - snapshot / glTF / PNG loaders and the orbit camera of libnmr (host side, plain C++17).
#include "host.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "value.h"

namespace nmr {

// ------------------------------------------------------------------------------------------------------------
// fp16 <-> fp32 on the host (parameters of type "float" are converted once at load)
// ------------------------------------------------------------------------------------------------------------
static uint16_t float_to_half_rn(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7FFFFFFFu;
    if (ax >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (ax > 0x7F800000u ? 0x200u : 0u));
    if (ax >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);
    if (ax < 0x38800000u) {
        if (ax < 0x33000000u) return (uint16_t)sign;
        const uint32_t e = ax >> 23, man = (ax & 0x7FFFFFu) | 0x800000u, shift = 126u - e;
        uint32_t hm = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1u);
        if (rem > half || (rem == half && (hm & 1u))) ++hm;
        return (uint16_t)(sign | hm);
    }
    uint32_t hm = (((ax >> 23) - 112u) << 10) | ((ax & 0x7FFFFFu) >> 13);
    const uint32_t rem = ax & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (hm & 1u))) ++hm;
    return (uint16_t)(sign | hm);
}

// ------------------------------------------------------------------------------------------------------------
// hash-grid level table (T/include/tiny-cuda-nn/encodings/grid.h:164-186, 196-205, 985-1016)
// ------------------------------------------------------------------------------------------------------------
void build_level_table(HostModel& m) {
    const float log2_pls = std::log2(m.per_level_scale);
    uint32_t offset = 0;
    for (int lvl = 0; lvl < m.n_levels; ++lvl) {
        const float scale = exp2f((float)lvl * log2_pls) * (float)m.base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1;
        const uint32_t max_params = 0xFFFFFFFFu / 2;
        uint32_t n = powf((float)res, 3) > (float)max_params ? max_params : res * res * res;
        n = (n + 7u) / 8u * 8u;
        n = std::min(n, 1u << m.log2_hashmap_size);
        m.offsets[lvl] = offset;
        offset += n;
        m.scales[lvl] = scale;
        m.resolutions[lvl] = res;
        // grid_index(): strides accumulate in uint32 while stride <= level size; hashed iff level size < final stride
        uint32_t stride = 1, st[3] = {0, 0, 0};
        for (int d = 0; d < 3 && stride <= n; ++d) { st[d] = stride; stride *= res; }
        m.dense[lvl] = !(n < stride);
        m.stride_y[lvl] = st[1];
        m.stride_z[lvl] = st[2];
    }
    m.offsets[m.n_levels] = offset;
}

static std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }

static void read_vec3(const Value& v, float out[3]) { for (int k = 0; k < 3; ++k) out[k] = v.at((size_t)k).as_float(); }

static void read_box(const Value& v, float mn[3], float mx[3]) { read_vec3(v.at("min"), mn); read_vec3(v.at("max"), mx); }

static void read_mat3(const Value& v, float out[9]) {
    for (size_t r = 0; r < 3 && r < v.size(); ++r) {
        const Value& row = v.at(r);
        for (size_t c = 0; c < 3 && c < row.size(); ++c) out[r * 3 + c] = row.at(c).as_float();
    }
}

static void check_mlp(const Value& net, const char* what, int& hidden) {
    const std::string otype = lower(net.value_str("otype", "FullyFusedMLP"));
    if (otype != "fullyfusedmlp" && otype != "megakernelmlp" && otype != "cutlassmlp")
        throw std::runtime_error(std::string(what) + ": unsupported network otype '" + otype + "'");
    if (net.value_int("n_neurons", 128) != 64) throw std::runtime_error(std::string(what) + ": only 64-neuron networks are supported");
    if (lower(net.value_str("activation", "ReLU")) != "relu") throw std::runtime_error(std::string(what) + ": only ReLU hidden activations are supported");
    if (lower(net.value_str("output_activation", "None")) != "none") throw std::runtime_error(std::string(what) + ": only output_activation None is supported");
    hidden = (int)net.value_int("n_hidden_layers", 5);
}

HostModel load_snapshot(const std::string& path) {
    const size_t dot = path.find_last_of('.');
    if (dot == std::string::npos || lower(path.substr(dot + 1)) != "msgpack")
        throw std::runtime_error("only .msgpack snapshots are supported");   // S/ngp/testbed.cu:1016-1018
    std::vector<uint8_t> file = read_file(path);
    Value root = parse_msgpack(file.data(), file.size());
    if (!root.is_map() || !root.contains("snapshot")) throw std::runtime_error("file does not contain a snapshot");
    const Value& snap = root.at("snapshot");
    if (snap.value_int("version", 0) < 1) throw std::runtime_error("snapshot uses an old format");
    if (snap.at("density_grid_size").as_int() != (int64_t)kGridSize) throw std::runtime_error("incompatible grid size");

    HostModel m;
    if (snap.contains("bounding_radius")) m.bounding_radius = snap.at("bounding_radius").as_float();
    const Value& nerf = snap.at("nerf");
    bool have_dataset_box = false;
    float ds_min[3], ds_max[3];
    bool is_hdr = false;
    if (nerf.contains("dataset")) {
        const Value& ds = nerf.at("dataset");
        m.aabb_scale = (int)ds.at("aabb_scale").as_int();
        if (ds.contains("render_aabb")) { read_box(ds.at("render_aabb"), ds_min, ds_max); have_dataset_box = true; }
        if (ds.contains("render_aabb_to_local")) read_mat3(ds.at("render_aabb_to_local"), m.render_aabb_to_local);
        is_hdr = ds.value_bool("is_hdr", false);
        if (ds.contains("scale")) m.dataset_scale = ds.at("scale").as_float();
        if (ds.contains("offset")) read_vec3(ds.at("offset"), m.dataset_offset);
        if (ds.contains("up")) read_vec3(ds.at("up"), m.dataset_up);
        m.from_mitsuba = ds.value_bool("from_mitsuba", false) ? 1 : 0;
    } else if (nerf.contains("aabb_scale")) {
        m.aabb_scale = (int)nerf.at("aabb_scale").as_int();
    }
    // load_nerf_post (S/ngp/testbed.cu:1085-1115)
    if (m.aabb_scale < 1 || (m.aabb_scale & (m.aabb_scale - 1)) != 0) throw std::runtime_error("aabb_scale must be a power of two");
    if (m.aabb_scale > (1 << (kCascades - 1))) throw std::runtime_error("aabb_scale exceeds 128");
    m.rgb_activation = is_hdr ? 3 : 2;
    const float half = 0.5f * (float)std::min(1 << (kCascades - 1), m.aabb_scale);
    for (int k = 0; k < 3; ++k) { m.aabb_min[k] = 0.5f - half; m.aabb_max[k] = 0.5f + half; m.render_aabb_min[k] = m.aabb_min[k]; m.render_aabb_max[k] = m.aabb_max[k]; }
    if (have_dataset_box) {
        bool empty = false;
        for (int k = 0; k < 3; ++k) empty |= ds_max[k] < ds_min[k];
        if (!empty) for (int k = 0; k < 3; ++k) { m.render_aabb_min[k] = std::max(ds_min[k], m.aabb_min[k]); m.render_aabb_max[k] = std::min(ds_max[k], m.aabb_max[k]); }
    }
    m.max_cascade = 0;
    while ((1 << m.max_cascade) < m.aabb_scale) ++m.max_cascade;
    m.cone_angle_constant = m.aabb_scale <= 1 ? 0.0f : (1.0f / 256.0f);

    // density grid (fp16) - S/ngp/testbed.cu:975-987
    const Value& dg = snap.at("density_grid_binary");
    if (dg.type != Value::Binary) throw std::runtime_error("density_grid_binary is not binary");
    const size_t n_grid = dg.bin_size / 2;
    if (n_grid != 0 && n_grid != (size_t)kGridCells * (size_t)(m.max_cascade + 1)) throw std::runtime_error("incompatible number of grid cascades");
    m.density_grid.resize(n_grid);
    if (n_grid) std::memcpy(m.density_grid.data(), dg.bin, n_grid * 2);

    if (snap.contains("render_aabb_to_local")) read_mat3(snap.at("render_aabb_to_local"), m.render_aabb_to_local);
    if (snap.contains("render_aabb")) read_box(snap.at("render_aabb"), m.render_aabb_min, m.render_aabb_max);

    // network configuration - reset_network (S/ngp/testbed.cu:1158-1236)
    const Value& enc = root.at("encoding");
    if (lower(enc.value_str("otype", "OneBlob")).find("grid") == std::string::npos) throw std::runtime_error("only grid encodings are supported");
    m.n_features_per_level = (int)enc.value_int("n_features_per_level", 2);
    if (m.n_features_per_level != 2) throw std::runtime_error("only n_features_per_level = 2 is supported");
    if (enc.contains("n_features") && enc.at("n_features").as_int() > 0) m.n_levels = (int)enc.at("n_features").as_int() / m.n_features_per_level;
    else m.n_levels = (int)enc.value_int("n_levels", 16);
    if (m.n_levels != 16) throw std::runtime_error("only 16-level grids (32 encoded features) are supported");
    m.log2_hashmap_size = (int)enc.value_int("log2_hashmap_size", 19);
    if (m.log2_hashmap_size < 1 || m.log2_hashmap_size > 28) throw std::runtime_error("log2_hashmap_size out of range");
    m.base_resolution = (int)enc.value_int("base_resolution", 0);
    if (!m.base_resolution) m.base_resolution = 1 << ((int)enc.value_int("log2_hashmap_size", 15) / 3);
    m.per_level_scale = (float)enc.value("per_level_scale", 0.0);
    if (m.per_level_scale <= 0.0f && m.n_levels > 1)
        m.per_level_scale = std::exp(std::log(2048.0f * (float)m.aabb_scale / (float)m.base_resolution) / (float)(m.n_levels - 1));
    if (lower(enc.value_str("interpolation", "Linear")) != "linear") throw std::runtime_error("only linear grid interpolation is supported");
    {
        const std::string otype = lower(enc.value_str("otype", "Grid"));
        const std::string dflt = otype == "tiledgrid" ? "Tiled" : (otype == "densegrid" ? "Dense" : "Hash");
        if (lower(enc.value_str("type", dflt)) != "hash") throw std::runtime_error("only hash grids are supported");
        const std::string h = lower(enc.value_str("hash", "CoherentPrime"));
        if (h == "prime") m.hash_type = 0; else if (h == "coherentprime") m.hash_type = 1; else if (h == "reversedprime") m.hash_type = 2;
        else throw std::runtime_error("unsupported hash '" + h + "'");
    }
    build_level_table(m);

    check_mlp(root.at("network"), "network", m.density_hidden);
    check_mlp(root.at("rgb_network"), "rgb_network", m.rgb_hidden);
    if (root.at("network").value_int("n_output_dims", 16) != 16) throw std::runtime_error("density network must have 16 outputs");
    if (m.density_hidden != 1 || m.rgb_hidden != 2)
        throw std::runtime_error("only the stock topology (density: 1 hidden layer, rgb: 2 hidden layers) is supported by the fused kernel");
    {
        // dir_encoding: Composite{SphericalHarmonics degree 4 on 3 dims, Identity on the (zero) extra dims} or plain SH
        const Value& de = root.at("dir_encoding");
        const Value* sh = &de;
        if (lower(de.value_str("otype", "")) == "composite") {
            if (!de.contains("nested") || de.at("nested").size() < 1) throw std::runtime_error("dir_encoding: empty composite");
            sh = &de.at("nested").at(0);
        }
        if (lower(sh->value_str("otype", "")) != "sphericalharmonics" || sh->value_int("degree", 4) != 4)
            throw std::runtime_error("dir_encoding: only SphericalHarmonics degree 4 is supported");
    }

    // parameters - Trainer::deserialize (T/include/tiny-cuda-nn/trainer.h:285-310)
    m.training_step = snap.value_int("training_step", 0);
    m.loss = (float)snap.value("loss", 0.0);
    const std::string ptype = snap.value_str("params_type", "__half");
    const Value& pb = snap.at("params_binary");
    if (pb.type != Value::Binary) throw std::runtime_error("params_binary is not binary");
    m.mlp_params = (size_t)64 * 32 + (size_t)16 * 64 + (size_t)64 * 32 + (size_t)64 * 64 + (size_t)16 * 64;
    const size_t expect = m.mlp_params + (size_t)m.offsets[m.n_levels] * 2;
    if (ptype == "__half") {
        if (pb.bin_size / 2 != expect) throw std::runtime_error("params_binary holds " + std::to_string(pb.bin_size / 2) + " parameters, network needs " + std::to_string(expect));
        m.params.resize(expect);
        std::memcpy(m.params.data(), pb.bin, expect * 2);
    } else if (ptype == "float") {
        if (pb.bin_size / 4 != expect) throw std::runtime_error("params_binary holds " + std::to_string(pb.bin_size / 4) + " parameters, network needs " + std::to_string(expect));
        m.params.resize(expect);
        for (size_t i = 0; i < expect; ++i) { float f; std::memcpy(&f, pb.bin + i * 4, 4); m.params[i] = float_to_half_rn(f); }
    } else {
        throw std::runtime_error("snapshot parameters must be of type float or __half");
    }
    return m;
}

// ------------------------------------------------------------------------------------------------------------
// PNG
// ------------------------------------------------------------------------------------------------------------
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

void decode_png(const uint8_t* data, size_t size, int& w, int& h, std::vector<uint8_t>& rgba) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || std::memcmp(data, sig, 8) != 0) throw std::runtime_error("png: bad signature");
    size_t p = 8;
    int bit_depth = 0, color_type = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool have_ihdr = false;
    while (p + 12 <= size) {
        const uint32_t len = be32(data + p);
        const uint8_t* tag = data + p + 4;
        if (p + 12 + (size_t)len > size) throw std::runtime_error("png: truncated chunk");
        const uint8_t* body = data + p + 8;
        if (!std::memcmp(tag, "IHDR", 4)) {
            if (len < 13) throw std::runtime_error("png: short IHDR");
            w = (int)be32(body); h = (int)be32(body + 4); bit_depth = body[8]; color_type = body[9]; interlace = body[12];
            have_ihdr = true;
        } else if (!std::memcmp(tag, "PLTE", 4)) palette.assign(body, body + len);
        else if (!std::memcmp(tag, "tRNS", 4)) trns.assign(body, body + len);
        else if (!std::memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(tag, "IEND", 4)) break;
        p += 12 + (size_t)len;
    }
    if (!have_ihdr || w <= 0 || h <= 0 || w > 16384 || h > 16384) throw std::runtime_error("png: bad header");
    if (bit_depth != 8 || interlace != 0) throw std::runtime_error("png: only 8-bit non-interlaced images are supported");
    int channels;
    switch (color_type) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break; case 4: channels = 2; break; case 6: channels = 4; break;
        default: throw std::runtime_error("png: bad colour type"); }
    const size_t stride = (size_t)w * channels;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size()) throw std::runtime_error("png: inflate failed");
    std::vector<uint8_t> img(stride * (size_t)h);
    for (int y = 0; y < h; ++y) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* src = raw.data() + (stride + 1) * y + 1;
        uint8_t* dst = img.data() + stride * y;
        const uint8_t* up = y ? img.data() + stride * (y - 1) : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)channels ? dst[x - channels] : 0;
            const int b = up ? up[x] : 0;
            const int c = (up && x >= (size_t)channels) ? up[x - channels] : 0;
            int v = src[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: { const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                          v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
                default: throw std::runtime_error("png: bad filter");
            }
            dst[x] = (uint8_t)v;
        }
    }
    rgba.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        uint8_t r, g, b, a = 255;
        const uint8_t* s = img.data() + i * channels;
        switch (color_type) {
            case 0: r = g = b = s[0]; break;
            case 2: r = s[0]; g = s[1]; b = s[2]; break;
            case 3: { const size_t k = s[0]; if (k * 3 + 2 >= palette.size()) throw std::runtime_error("png: palette index out of range");
                      r = palette[k * 3]; g = palette[k * 3 + 1]; b = palette[k * 3 + 2]; if (k < trns.size()) a = trns[k]; break; }
            case 4: r = g = b = s[0]; a = s[1]; break;
            default: r = s[0]; g = s[1]; b = s[2]; a = s[3]; break;
        }
        rgba[i * 4] = r; rgba[i * 4 + 1] = g; rgba[i * 4 + 2] = b; rgba[i * 4 + 3] = a;
    }
}

// ------------------------------------------------------------------------------------------------------------
// glTF
// ------------------------------------------------------------------------------------------------------------
static std::vector<uint8_t> base64_decode(const std::string& s, size_t start) {
    std::vector<uint8_t> out;
    uint32_t acc = 0; int bits = 0;
    for (size_t i = start; i < s.size(); ++i) {
        const char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A'; else if (c >= 'a' && c <= 'z') v = c - 'a' + 26; else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+') v = 62; else if (c == '/') v = 63; else continue;
        acc = (acc << 6) | (uint32_t)v; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)((acc >> bits) & 0xFF)); }
    }
    return out;
}

static std::vector<uint8_t> load_uri(const std::string& uri, const std::string& base_dir) {
    if (uri.rfind("data:", 0) == 0) {
        const size_t comma = uri.find(',');
        if (comma == std::string::npos) throw std::runtime_error("gltf: malformed data URI");
        return base64_decode(uri, comma + 1);
    }
    return read_file(base_dir + uri);
}

HostMesh load_gltf(const std::string& path) {
    std::vector<uint8_t> file = read_file(path);
    const size_t slash = path.find_last_of("/\\");
    const std::string base_dir = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
    Value doc;
    std::vector<uint8_t> glb_bin;
    const bool is_glb = path.size() >= 4 && path.compare(path.size() - 4, 4, ".glb") == 0;
    if (is_glb) {
        if (file.size() < 20 || std::memcmp(file.data(), "glTF", 4) != 0) throw std::runtime_error("glb: bad header");
        size_t p = 12;
        bool have_json = false;
        while (p + 8 <= file.size()) {
            uint32_t len, type; std::memcpy(&len, file.data() + p, 4); std::memcpy(&type, file.data() + p + 4, 4);
            if (p + 8 + (size_t)len > file.size()) throw std::runtime_error("glb: truncated chunk");
            if (type == 0x4E4F534Au) { doc = parse_json((const char*)file.data() + p + 8, len); have_json = true; }
            else if (type == 0x004E4942u) glb_bin.assign(file.begin() + (long)p + 8, file.begin() + (long)p + 8 + len);
            p += 8 + (size_t)len;
        }
        if (!have_json) throw std::runtime_error("glb: no JSON chunk");
    } else {
        doc = parse_json((const char*)file.data(), file.size());
    }
    if (!doc.is_map()) throw std::runtime_error("gltf: document is not an object");

    std::vector<std::vector<uint8_t>> buffers;
    if (doc.contains("buffers")) for (const Value& b : doc.at("buffers").arr) {
        if (b.contains("uri")) buffers.push_back(load_uri(b.at("uri").as_string(), base_dir));
        else buffers.push_back(glb_bin);
    }
    // every offset / length / count / index of the document is untrusted: negative values are rejected before the cast to size_t,
    // and range checks are written so that they cannot wrap (len > size || off > size - len)
    auto nonneg = [](int64_t v, const char* what) -> size_t {
        if (v < 0) throw std::runtime_error(std::string("gltf: negative ") + what);
        return (size_t)v;
    };
    auto elem = [&](const char* array, int64_t idx) -> const Value& {
        if (!doc.contains(array)) throw std::runtime_error(std::string("gltf: no ") + array);
        const Value& arr = doc.at(array);
        if (idx < 0 || (size_t)idx >= arr.size()) throw std::runtime_error(std::string("gltf: index into ") + array + " out of range");
        return arr.at((size_t)idx);
    };
    auto view_bytes = [&](int64_t view_idx, size_t& stride) -> std::pair<const uint8_t*, size_t> {
        const Value& bv = elem("bufferViews", view_idx);
        const size_t bi = nonneg(bv.at("buffer").as_int(), "buffer index");
        if (bi >= buffers.size()) throw std::runtime_error("gltf: buffer index out of range");
        const size_t off = nonneg(bv.value_int("byteOffset", 0), "byteOffset"), len = nonneg(bv.at("byteLength").as_int(), "byteLength");
        const size_t size = buffers[bi].size();
        if (len > size || off > size - len) throw std::runtime_error("gltf: bufferView exceeds buffer");
        stride = nonneg(bv.value_int("byteStride", 0), "byteStride");
        if (stride > 4096) throw std::runtime_error("gltf: byteStride out of range");
        return {buffers[bi].data() + off, len};
    };
    // accessor of `count` elements of `esz` bytes, `stride` apart, starting `off` bytes into a view of `avail` bytes
    auto check_span = [](size_t count, size_t off, size_t stride, size_t esz, size_t avail, const char* what) {
        if (!count) return;
        if (esz > avail || off > avail - esz || (count - 1) > (avail - esz - off) / (stride ? stride : 1)) throw std::runtime_error(std::string("gltf: ") + what + " exceeds bufferView");
    };
    auto read_accessor_f32 = [&](int64_t acc_idx, int ncomp, std::vector<float>& out) {
        const Value& a = elem("accessors", acc_idx);
        if (a.at("componentType").as_int() != 5126) throw std::runtime_error("gltf: vertex attributes must be float");
        const std::string type = a.at("type").as_string();
        const int have = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : 0;
        if (have < ncomp) throw std::runtime_error("gltf: accessor type too narrow");
        size_t stride; auto vb = view_bytes(a.at("bufferView").as_int(), stride);
        const size_t count = nonneg(a.at("count").as_int(), "accessor count"), off = nonneg(a.value_int("byteOffset", 0), "accessor byteOffset");
        if (!stride) stride = (size_t)have * 4;
        check_span(count, off, stride, (size_t)ncomp * 4, vb.second, "accessor");
        for (size_t i = 0; i < count; ++i) for (int c = 0; c < ncomp; ++c) { float f; std::memcpy(&f, vb.first + off + i * stride + (size_t)c * 4, 4); out.push_back(f); }
        return count;
    };
    auto read_indices = [&](int64_t acc_idx, uint32_t base, std::vector<uint32_t>& out) {
        const Value& a = elem("accessors", acc_idx);
        const int64_t ct = a.at("componentType").as_int();
        const size_t esz = ct == 5121 ? 1 : ct == 5123 ? 2 : ct == 5125 ? 4 : 0;
        if (!esz) throw std::runtime_error("gltf: unsupported index type");
        size_t stride; auto vb = view_bytes(a.at("bufferView").as_int(), stride);
        const size_t count = nonneg(a.at("count").as_int(), "accessor count"), off = nonneg(a.value_int("byteOffset", 0), "accessor byteOffset");
        if (!stride) stride = esz;
        check_span(count, off, stride, esz, vb.second, "index accessor");
        for (size_t i = 0; i < count; ++i) {
            const uint8_t* q = vb.first + off + i * stride;
            uint32_t v = esz == 1 ? q[0] : esz == 2 ? (uint32_t)(q[0] | (q[1] << 8)) : (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
            out.push_back(base + v);
        }
    };

    HostMesh m;
    // texture reference {"index": i} -> decoded RGBA8 image; an unusable image is reported and the factor alone is used
    auto read_texture = [&](const Value& ref, const char* what) -> HostMesh::Texture {
        HostMesh::Texture t;
        try {
            const Value& tex = elem("textures", ref.at("index").as_int());
            const Value& img = elem("images", tex.at("source").as_int());
            std::vector<uint8_t> bytes;
            if (img.contains("uri")) bytes = load_uri(img.at("uri").as_string(), base_dir);
            else { size_t st; auto vb = view_bytes(img.at("bufferView").as_int(), st); bytes.assign(vb.first, vb.first + vb.second); }
            decode_png(bytes.data(), bytes.size(), t.w, t.h, t.rgba8);
        } catch (const std::exception& e) {
            t = HostMesh::Texture{};
            m.warning += (m.warning.empty() ? "" : "; ") + std::string(what) + " texture unusable (" + e.what() + "), using the factor only";
        }
        return t;
    };
    const size_t scene_idx = nonneg(doc.value_int("scene", 0), "scene index");
    if (!doc.contains("scenes") || scene_idx >= doc.at("scenes").size()) throw std::runtime_error("gltf: no default scene");
    const Value& scene_nodes = doc.at("scenes").at(scene_idx).at("nodes");
    if (scene_nodes.size() == 0) throw std::runtime_error("gltf: scene has no nodes");
    bool have_material = false;
    // depth-first over the scene like GltfScene::getMeshPrimitives; child transforms are ignored as in the reference
    std::vector<int64_t> stack;
    for (size_t i = scene_nodes.size(); i-- > 0;) stack.push_back(scene_nodes.at(i).as_int());
    bool missing_normals = false;
    struct PrimRange { uint32_t vbase; size_t nv, ibegin, iend; };
    std::vector<PrimRange> untangented;     // primitives without TANGENT: generated below, one primitive at a time like the reference
    // a node is visited once: a child list that leads back to an ancestor (or a node shared by two parents) is not a tree
    std::vector<uint8_t> visited(doc.contains("nodes") ? doc.at("nodes").size() : 0, 0);
    while (!stack.empty()) {
        const int64_t ni = stack.back(); stack.pop_back();
        const Value& node = elem("nodes", ni);
        if (visited[(size_t)ni]) throw std::runtime_error("gltf: node graph is not a tree (cycle or shared node)");
        visited[(size_t)ni] = 1;
        if (node.contains("children")) for (const Value& c : node.at("children").arr) stack.push_back(c.as_int());
        if (!node.contains("mesh")) continue;
        const Value& mesh = elem("meshes", node.at("mesh").as_int());
        for (const Value& prim : mesh.at("primitives").arr) {
            if (prim.value_int("mode", 4) != 4) continue;   // triangles only
            const Value& attrs = prim.at("attributes");
            if (!attrs.contains("POSITION")) continue;
            const uint32_t base = (uint32_t)(m.positions.size() / 3);
            const size_t nv = read_accessor_f32(attrs.at("POSITION").as_int(), 3, m.positions);
            // every attribute of a primitive has one entry per vertex (glTF 2.0, 3.7.2.1): a shorter NORMAL / TEXCOORD_0 accessor
            // would leave transform_mesh and the shader reading past the end of the arrays
            if (attrs.contains("NORMAL")) { if (read_accessor_f32(attrs.at("NORMAL").as_int(), 3, m.normals) != nv) throw std::runtime_error("gltf: NORMAL count differs from POSITION count"); }
            else { m.normals.resize(m.normals.size() + nv * 3, 0.f); missing_normals = true; }
            if (attrs.contains("TEXCOORD_0")) { if (read_accessor_f32(attrs.at("TEXCOORD_0").as_int(), 2, m.texcoords) != nv) throw std::runtime_error("gltf: TEXCOORD_0 count differs from POSITION count"); }
            else m.texcoords.resize(m.texcoords.size() + nv * 2, 0.f);
            if (attrs.contains("TANGENT")) { if (read_accessor_f32(attrs.at("TANGENT").as_int(), 4, m.tangents) != nv) throw std::runtime_error("gltf: TANGENT count differs from POSITION count"); }
            else { m.tangents.resize(m.tangents.size() + nv * 4, 0.f); untangented.push_back({base, nv, m.indices.size(), 0}); }
            if (prim.contains("indices")) read_indices(prim.at("indices").as_int(), base, m.indices);
            else for (uint32_t i = 0; i < (uint32_t)nv; ++i) m.indices.push_back(base + i);
            if (!attrs.contains("TANGENT")) untangented.back().iend = m.indices.size() / 3 * 3;
            {   // lens material?  (flags cover every triangle appended so far)
                bool lens = false; float ior = 1.5f, transmission = 1.f, tint[3] = {1.f, 1.f, 1.f};
                if (prim.contains("material") && doc.contains("materials")) {
                    const Value& mat = doc.at("materials").at((size_t)prim.at("material").as_int());
                    float alpha = 1.f;
                    if (mat.contains("pbrMetallicRoughness") && mat.at("pbrMetallicRoughness").contains("baseColorFactor")) {
                        const Value& bc = mat.at("pbrMetallicRoughness").at("baseColorFactor");
                        for (int k = 0; k < 3; ++k) tint[k] = bc.at((size_t)k).as_float();
                        alpha = bc.at(3).as_float();
                    }
                    if (mat.contains("alphaMode") && mat.at("alphaMode").as_string() == "BLEND" && alpha < 1.f) { lens = true; transmission = 1.f - alpha; }
                    if (mat.contains("extensions")) {
                        const Value& ext = mat.at("extensions");
                        if (ext.contains("KHR_materials_transmission")) {
                            const float tf = (float)ext.at("KHR_materials_transmission").value("transmissionFactor", 0.0);
                            if (tf > 0.f) { lens = true; transmission = tf; }
                        }
                        if (ext.contains("KHR_materials_ior")) ior = (float)ext.at("KHR_materials_ior").value("ior", 1.5);
                    }
                }
                m.tri_lens.resize(m.indices.size() / 3, lens ? 1 : 0);
                if (lens && !m.has_lens) { m.has_lens = true; m.lens_ior = ior; m.lens_transmission = transmission; std::memcpy(m.lens_tint, tint, 12); }
            }
            if (!have_material && prim.contains("material") && doc.contains("materials")) {
                have_material = true;
                const Value& mat = doc.at("materials").at((size_t)prim.at("material").as_int());
                if (mat.contains("emissiveFactor")) read_vec3(mat.at("emissiveFactor"), m.emissive);
                if (mat.contains("pbrMetallicRoughness")) {
                    const Value& pbr = mat.at("pbrMetallicRoughness");
                    if (pbr.contains("baseColorFactor")) for (int k = 0; k < 4; ++k) m.base_color[k] = pbr.at("baseColorFactor").at((size_t)k).as_float();
                    m.metallic = (float)pbr.value("metallicFactor", 1.0);
                    m.roughness = (float)pbr.value("roughnessFactor", 1.0);
                    if (pbr.contains("baseColorTexture")) {
                        HostMesh::Texture t = read_texture(pbr.at("baseColorTexture"), "base colour");
                        m.tex_w = t.w; m.tex_h = t.h; m.tex_rgba8 = std::move(t.rgba8);
                    }
                    if (pbr.contains("metallicRoughnessTexture")) m.tex_metallic_roughness = read_texture(pbr.at("metallicRoughnessTexture"), "metallic-roughness");
                }
                if (mat.contains("emissiveTexture")) m.tex_emissive = read_texture(mat.at("emissiveTexture"), "emissive");
                if (mat.contains("normalTexture")) {
                    m.tex_normal = read_texture(mat.at("normalTexture"), "normal");
                    m.normal_scale = (float)mat.at("normalTexture").value("scale", 1.0);
                }
                if (mat.contains("occlusionTexture")) {
                    m.tex_occlusion = read_texture(mat.at("occlusionTexture"), "occlusion");
                    m.occlusion_strength = (float)mat.at("occlusionTexture").value("strength", 1.0);
                }
            }
        }
    }
    if (m.indices.empty()) throw std::runtime_error("gltf: no triangles found");
    m.indices.resize(m.indices.size() / 3 * 3);
    m.tri_lens.resize(m.indices.size() / 3, 0);
    const uint32_t nverts = (uint32_t)(m.positions.size() / 3);
    for (uint32_t idx : m.indices) if (idx >= nverts) throw std::runtime_error("gltf: index out of range");
    if (missing_normals) {
        // area-weighted vertex normals where the file has none (the reference would read out of bounds)
        std::vector<float> acc(m.positions.size(), 0.f);
        for (size_t t = 0; t + 2 < m.indices.size(); t += 3) {
            const float* a = &m.positions[m.indices[t] * 3]; const float* b = &m.positions[m.indices[t + 1] * 3]; const float* c = &m.positions[m.indices[t + 2] * 3];
            const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
            const float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) acc[m.indices[t + k] * 3 + j] += n[j];
        }
        for (uint32_t v = 0; v < nverts; ++v) {
            float* n = &m.normals[v * 3];
            if (n[0] == 0.f && n[1] == 0.f && n[2] == 0.f) { n[0] = acc[v * 3]; n[1] = acc[v * 3 + 1]; n[2] = acc[v * 3 + 2]; }
        }
    }
    for (const PrimRange& pr : untangented) {
        // Mikkelsen's tangent space per primitive (S/gltf_scene.cpp:150-155); vertices outside every triangle get the method's
        // default frame (1, 0, 0, -1) instead of the zeros the reference leaves there
        if (pr.iend <= pr.ibegin) continue;
        std::vector<uint32_t> local(m.indices.begin() + (ptrdiff_t)pr.ibegin, m.indices.begin() + (ptrdiff_t)pr.iend);
        for (uint32_t& i : local) { i -= pr.vbase; if (i >= pr.nv) throw std::runtime_error("gltf: index outside its primitive's vertices"); }
        float* tg = &m.tangents[(size_t)pr.vbase * 4];
        for (size_t v = 0; v < pr.nv; ++v) { tg[v * 4] = 1.f; tg[v * 4 + 1] = 0.f; tg[v * 4 + 2] = 0.f; tg[v * 4 + 3] = -1.f; }
        try {
            mikk_tangents(&m.positions[(size_t)pr.vbase * 3], &m.normals[(size_t)pr.vbase * 3], &m.texcoords[(size_t)pr.vbase * 2], pr.nv,
                          local.data(), local.size(), tg);
        } catch (const std::length_error& e) {
            // a mesh the generator refuses still loads: only a normal map reads the tangents, and it gets the default frame
            for (size_t v = 0; v < pr.nv; ++v) { tg[v * 4] = 1.f; tg[v * 4 + 1] = 0.f; tg[v * 4 + 2] = 0.f; tg[v * 4 + 3] = -1.f; }
            m.warning += (m.warning.empty() ? "" : "; ") + std::string("tangents not generated (") + e.what() + "), default frames used";
        }
    }
    // node 0 TRS (GltfLoader::traverse, S/gltf_scene.cpp:63-118); matrix-form nodes are not decomposed (load_mesh overwrites TRS anyway)
    const Value& n0 = doc.at("nodes").at((size_t)scene_nodes.at(0).as_int());
    if (n0.contains("translation")) read_vec3(n0.at("translation"), m.t);
    if (n0.contains("scale")) read_vec3(n0.at("scale"), m.s);
    if (n0.contains("rotation")) { const Value& r = n0.at("rotation"); m.r_wxyz[0] = r.at(3).as_float(); m.r_wxyz[1] = r.at(0).as_float(); m.r_wxyz[2] = r.at(1).as_float(); m.r_wxyz[3] = r.at(2).as_float(); }
    return m;
}

// T * R * S applied to positions; (R S)^-T = R S^-1 applied to normals.  R = glm::mat3_cast(quat(w,x,y,z)).
void transform_mesh(const HostMesh& m, const float t[3], const float s[3], const float q[4], std::vector<float>& wp, std::vector<float>& wn) {
    const float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    const float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    const float R[9] = {
        1.f - 2.f * (qyy + qzz), 2.f * (qxy - qwz), 2.f * (qxz + qwy),
        2.f * (qxy + qwz), 1.f - 2.f * (qxx + qzz), 2.f * (qyz - qwx),
        2.f * (qxz - qwy), 2.f * (qyz + qwx), 1.f - 2.f * (qxx + qyy)};
    const size_t nv = m.positions.size() / 3;
    wp.resize(nv * 3); wn.resize(nv * 3);
    for (size_t i = 0; i < nv; ++i) {
        const float px = m.positions[i * 3] * s[0], py = m.positions[i * 3 + 1] * s[1], pz = m.positions[i * 3 + 2] * s[2];
        wp[i * 3 + 0] = ((R[0] * px + R[1] * py) + R[2] * pz) + t[0];
        wp[i * 3 + 1] = ((R[3] * px + R[4] * py) + R[5] * pz) + t[1];
        wp[i * 3 + 2] = ((R[6] * px + R[7] * py) + R[8] * pz) + t[2];
        const float nx = m.normals[i * 3] / s[0], ny = m.normals[i * 3 + 1] / s[1], nz = m.normals[i * 3 + 2] / s[2];
        wn[i * 3 + 0] = (R[0] * nx + R[1] * ny) + R[2] * nz;
        wn[i * 3 + 1] = (R[3] * nx + R[4] * ny) + R[5] * nz;
        wn[i * 3 + 2] = (R[6] * nx + R[7] * ny) + R[8] * nz;
    }
}

void transform_tangent_frames(const HostMesh& m, const float s[3], const float q[4], std::vector<float>& wtbn, float nmat[9]) {
    const float qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    const float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
    const float R[9] = {
        1.f - 2.f * (qyy + qzz), 2.f * (qxy - qwz), 2.f * (qxz + qwy),
        2.f * (qxy + qwz), 1.f - 2.f * (qxx + qzz), 2.f * (qyz - qwx),
        2.f * (qxz - qwy), 2.f * (qyz + qwx), 1.f - 2.f * (qxx + qyy)};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) nmat[i * 3 + j] = R[i * 3 + j] / s[j];
    const size_t nv = m.positions.size() / 3;
    wtbn.assign(nv * 8, 0.f);
    if (m.tangents.size() < nv * 4) return;
    for (size_t i = 0; i < nv; ++i) {
        const float nx = m.normals[i * 3] * s[0], ny = m.normals[i * 3 + 1] * s[1], nz = m.normals[i * 3 + 2] * s[2];
        const float tx = m.tangents[i * 4] * s[0], ty = m.tangents[i * 4 + 1] * s[1], tz = m.tangents[i * 4 + 2] * s[2];
        float* o = &wtbn[i * 8];
        for (int k = 0; k < 3; ++k) {
            o[k] = (R[k * 3] * nx + R[k * 3 + 1] * ny) + R[k * 3 + 2] * nz;
            o[3 + k] = (R[k * 3] * tx + R[k * 3 + 1] * ty) + R[k * 3 + 2] * tz;
        }
        o[6] = m.tangents[i * 4 + 3];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Orbit camera
// ------------------------------------------------------------------------------------------------------------
namespace {
struct F3 { float x, y, z; };
inline F3 f3(const float* p) { return {p[0], p[1], p[2]}; }
inline float len(F3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
inline F3 unit(F3 a) { const float l = len(a); return {a.x / l, a.y / l, a.z / l}; }
inline F3 cross(F3 a, F3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// flythrough_camera_look_to, right-handed (R/dependencies/flythrough_camera.h:256-333)
void build_view(const float eye[3], const float look[3], const float up[3], float view[16]) {
    const F3 upn = unit(f3(up));
    F3 f = unit(f3(look));
    const F3 s = unit(cross(f, upn));
    const F3 u = unit(cross(s, f));
    f = {-f.x, -f.y, -f.z};
    const float tx = s.x * -eye[0] + s.y * -eye[1] + s.z * -eye[2];
    const float ty = u.x * -eye[0] + u.y * -eye[1] + u.z * -eye[2];
    const float tz = f.x * -eye[0] + f.y * -eye[1] + f.z * -eye[2];
    const float v[16] = {s.x, u.x, f.x, 0.f, s.y, u.y, f.y, 0.f, s.z, u.z, f.z, 0.f, tx, ty, tz, 1.f};
    std::memcpy(view, v, sizeof(v));
}
}  // namespace

OrbitCamera::OrbitCamera() { build_view(eye, look, up, view); }

// orbitcam (S/orbit_camera.h:7-77) as called by NerfMeshRenderer::orbit (S/nerf_mesh_renderer.cu:896-899)
void OrbitCamera::orbit(float delta_azimuth, float delta_polar, float delta_scroll) {
    const double kPi = 3.14159265359;   // the reference redefines M_PI with this literal
    const F3 rel = {eye[0] - pivot[0], eye[1] - pivot[1], eye[2] - pivot[2]};
    float radius = len(rel);
    const F3 d = unit(rel);
    float azimuth = atan2f(d.z, d.x);
    float polar = atan2f(d.y, sqrtf(d.x * d.x + d.z * d.z));
    azimuth += delta_azimuth;
    azimuth = fmodf(azimuth, (float)(2 * kPi));
    if (azimuth < 0.f) azimuth = (float)((double)azimuth + 2 * kPi);
    polar += delta_polar;
    const float cap = (float)(kPi / 2.f - 0.001f);
    polar = fminf(cap, fmaxf(-cap, polar));
    radius -= delta_scroll * radius * 0.1f;
    if (radius < 1.f) radius = 1.f;
    const float sa = sinf(azimuth), ca = cosf(azimuth), sp = sinf(polar), cp = cosf(polar);
    eye[0] = pivot[0] + radius * cp * ca;
    eye[1] = pivot[1] + radius * sp;
    eye[2] = pivot[2] + radius * cp * sa;
    for (int k = 0; k < 3; ++k) look[k] = pivot[k] - eye[k];
    build_view(eye, look, up, view);
}

void OrbitCamera::look_from(const float eye3[3], const float look3[3]) {
    for (int k = 0; k < 3; ++k) { eye[k] = eye3[k]; look[k] = look3[k]; }
    build_view(eye, look, up, view);
}
// One pose of the trajectory tool's arc (S/nerf_mesh_renderer.cu:649-658): the eye on a circle of `distance` at `height`, looking at
// `lookat` (glm::normalize = v * inversesqrt(dot(v, v)), dot summed left to right)
void OrbitCamera::trajectory_pose(float angle, float distance, float height, const float lookat3[3]) {
    const float e[3] = {cosf(angle) * distance, height, sinf(angle) * distance};
    const float d[3] = {lookat3[0] - e[0], lookat3[1] - e[1], lookat3[2] - e[2]};
    const float inv = 1.0f / sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    const float l[3] = {d[0] * inv, d[1] * inv, d[2] * inv};
    look_from(e, l);
}

// updateModelViewProj (S/nerf_mesh_renderer.cu:919-932): vLength = tanf(0.5f * 45) is evaluated in radians
void OrbitCamera::matrix(int screen_w, int screen_h, float out[12]) const {
    const float aspect = (float)(uint32_t)screen_w / (float)(uint32_t)screen_h;
    const float v_len = tanf(0.5f * 45);
    const float u_len = v_len * aspect;
    for (int k = 0; k < 3; ++k) {
        out[k] = view[4 * k + 0] * u_len;
        out[3 + k] = view[4 * k + 1] * v_len;
        out[6 + k] = -view[4 * k + 2];
        out[9 + k] = eye[k];
    }
}

}  // namespace nmr
