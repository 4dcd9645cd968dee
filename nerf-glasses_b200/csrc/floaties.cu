// floaties.cu - GPU floatie pruning: connected components on the cascaded occupancy grid, keep the most important one.
//
// Replaces NerfMeshRenderer::removeFloaties (S/nerf_mesh_renderer.cu:901-917): dumpDensityGrid (D2H + de-Mortonisation,
// :239-287) -> NgpGrid hash-set BFS clustering (S/floatyremover.h:35-234) -> max_element by point_set_importance
// (:253-266) -> to_ngp_grid (:236-251) -> loadDensityGrid (:289-336), all single-threaded CPU code in the reference.
// Here the 2 MiB bitfield never leaves the device:
//   1. label every "point" (set cell outside the inner 64^3 of cascades >= 1) with its own linear index
//   2. lock-free union-find over the reference's neighbour relation (6-neighbourhood inside a cascade plus the
//      child<->parent links across cascade faces), smaller index wins, so a component's root is its smallest index
//   3. per-root size and importance sum(16 - 2^level) with integer atomics
//   4. best root = max importance among components with >= 2 points (isolated points never form a cluster in the
//      reference), ties broken towards the smaller root (the CPU oracle's discovery order)
//   5. rebuild the bitfield from the kept component and its parents in every coarser cascade.
#include "kernels.cuh"

namespace nmr {

namespace {

constexpr uint32_t kCells = NERF_CASCADES * GRID_CELLS;   // 16 777 216
constexpr uint32_t kNone = 0xFFFFFFFFu;

__device__ __forceinline__ void decode(uint32_t c, int& x, int& y, int& z, int& l) { x = c & 127; y = (c >> 7) & 127; z = (c >> 14) & 127; l = c >> 21; }
__device__ __forceinline__ uint32_t encode(int x, int y, int z, int l) { return (uint32_t)x | ((uint32_t)y << 7) | ((uint32_t)z << 14) | ((uint32_t)l << 21); }

__device__ __forceinline__ bool bit_at(const uint8_t* __restrict__ bf, int x, int y, int z, int l) {
    const uint32_t idx = morton3D((uint32_t)x, (uint32_t)y, (uint32_t)z);
    return (bf[idx / 8 + (GRID_CELLS / 8) * (uint32_t)l] >> (idx % 8)) & 1u;
}
// NgpGrid ctor: cascades >= 1 contribute only cells outside their inner 64^3 block
__device__ __forceinline__ bool is_point(const uint8_t* __restrict__ bf, int x, int y, int z, int l) {
    if ((unsigned)x > 127u || (unsigned)y > 127u || (unsigned)z > 127u || (unsigned)l > 7u) return false;
    if (l > 0 && x >= 32 && x < 96 && y >= 32 && y < 96 && z >= 32 && z < 96) return false;
    return bit_at(bf, x, y, z, l);
}

__global__ void k_init(const uint8_t* __restrict__ bf, uint32_t* __restrict__ label) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kCells) return;
    int x, y, z, l; decode(c, x, y, z, l);
    label[c] = is_point(bf, x, y, z, l) ? c : kNone;
}

__device__ __forceinline__ uint32_t find_root(uint32_t* label, uint32_t i) {
    uint32_t p = atomicAdd(&label[i], 0u);   // atomic read: other threads relink concurrently
    while (p != i) { i = p; p = atomicAdd(&label[i], 0u); }
    return i;
}
__device__ void unite(uint32_t* label, uint32_t a, uint32_t b) {
    while (true) {
        a = find_root(label, a); b = find_root(label, b);
        if (a == b) return;
        if (a > b) { const uint32_t t = a; a = b; b = t; }
        const uint32_t old = atomicMin(&label[b], a);
        if (old == b) return;
        b = old;
    }
}

// Every undirected edge of get_neighbors (S/floatyremover.h:60-193) is visited from at least one endpoint: +x/+y/+z inside a
// cascade, and the child -> parent links at the six outer faces (their reverse, parent -> 4 children, is the same edge set).
__global__ void k_union(const uint8_t* __restrict__ bf, uint32_t* __restrict__ label) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kCells || label[c] == kNone) return;
    int x, y, z, l; decode(c, x, y, z, l);
    if (x + 1 < 128 && is_point(bf, x + 1, y, z, l)) unite(label, c, encode(x + 1, y, z, l));
    if (y + 1 < 128 && is_point(bf, x, y + 1, z, l)) unite(label, c, encode(x, y + 1, z, l));
    if (z + 1 < 128 && is_point(bf, x, y, z + 1, l)) unite(label, c, encode(x, y, z + 1, l));
    if (l < 7) {
        const int mx = 32 + x / 2, my = 32 + y / 2, mz = 32 + z / 2;
        if (x == 0 && is_point(bf, 31, my, mz, l + 1)) unite(label, c, encode(31, my, mz, l + 1));
        if (x == 127 && is_point(bf, 96, my, mz, l + 1)) unite(label, c, encode(96, my, mz, l + 1));
        if (y == 0 && is_point(bf, mx, 31, mz, l + 1)) unite(label, c, encode(mx, 31, mz, l + 1));
        if (y == 127 && is_point(bf, mx, 96, mz, l + 1)) unite(label, c, encode(mx, 96, mz, l + 1));
        if (z == 0 && is_point(bf, mx, my, 31, l + 1)) unite(label, c, encode(mx, my, 31, l + 1));
        if (z == 127 && is_point(bf, mx, my, 96, l + 1)) unite(label, c, encode(mx, my, 96, l + 1));
    }
}

__global__ void k_flatten_score(uint32_t* __restrict__ label, int* __restrict__ score, int* __restrict__ size) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kCells || label[c] == kNone) return;
    const uint32_t root = find_root(label, c);
    label[c] = root;   // roots keep pointing at themselves, so concurrent finds stay correct
    atomicAdd(&score[root], 16 - (1 << (c >> 21)));
    atomicAdd(&size[root], 1);
}

// results[0] = clusters, results[1] = size of the kept cluster, results[2] = packed best key
__global__ void k_best(const uint32_t* __restrict__ label, const int* __restrict__ score, const int* __restrict__ size, unsigned long long* __restrict__ results) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kCells || label[c] != c || size[c] < 2) return;
    atomicAdd(&results[0], 1ull);
    const unsigned long long key = ((unsigned long long)(uint32_t)(score[c] + 0x40000000) << 32) | (unsigned long long)(kNone - c);
    atomicMax(&results[2], key);
}

__global__ void k_clear(const unsigned long long* __restrict__ results, uint32_t* __restrict__ bf_words) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < kCells / 32 && results[2] != 0ull) bf_words[w] = 0u;   // without any cluster the grid is left as it was
}

__global__ void k_rebuild(const uint32_t* __restrict__ label, const int* __restrict__ size, unsigned long long* __restrict__ results, uint32_t* __restrict__ bf_words) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= kCells) return;
    const unsigned long long key = results[2];
    if (key == 0ull) return;                              // no cluster at all: the reference dereferences end(); we keep nothing
    const uint32_t best = kNone - (uint32_t)(key & 0xFFFFFFFFull);
    if (c == best) results[1] = (unsigned long long)size[best];
    if (label[c] != best) return;
    int x, y, z, l; decode(c, x, y, z, l);
    for (int lvl = l; lvl < 8; ++lvl) {                   // to_ngp_grid: the cell and its parents in all coarser cascades
        const uint32_t idx = morton3D((uint32_t)x, (uint32_t)y, (uint32_t)z) + GRID_CELLS * (uint32_t)lvl;
        atomicOr(&bf_words[idx >> 5], 1u << (idx & 31));
        x = 32 + x / 2; y = 32 + y / 2; z = 32 + z / 2;
    }
}

}  // namespace

size_t floaties_label_bytes() { return (size_t)kCells * sizeof(uint32_t); }
size_t floaties_scratch_bytes() { return 64 + 2 * (size_t)kCells * sizeof(int); }

void launch_remove_floaties(uint8_t* d_bitfield, int max_cascade, uint32_t* d_labels, unsigned long long* d_scratch, cudaStream_t s) {
    (void)max_cascade;
    int* score = reinterpret_cast<int*>(reinterpret_cast<char*>(d_scratch) + 64);
    int* size = score + kCells;
    cudaMemsetAsync(d_scratch, 0, floaties_scratch_bytes(), s);
    const unsigned blocks = kCells / 256;
    k_init<<<blocks, 256, 0, s>>>(d_bitfield, d_labels);
    k_union<<<blocks, 256, 0, s>>>(d_bitfield, d_labels);
    k_flatten_score<<<blocks, 256, 0, s>>>(d_labels, score, size);
    k_best<<<blocks, 256, 0, s>>>(d_labels, score, size, d_scratch);
    k_clear<<<kCells / 32 / 256, 256, 0, s>>>(d_scratch, reinterpret_cast<uint32_t*>(d_bitfield));
    k_rebuild<<<blocks, 256, 0, s>>>(d_labels, size, d_scratch, reinterpret_cast<uint32_t*>(d_bitfield));
}

}  // namespace nmr
