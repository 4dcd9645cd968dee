// Drives the host-side file readers of libnmr (csrc/host.cpp, csrc/value.cpp: msgpack snapshot, glTF + embedded / external buffers,
// PNG) over a list of files under AddressSanitizer + UBSan (tests/test_host_sanitized.py builds and runs it).  A reader may reject a
// file by throwing; it must not read or write out of bounds, overflow, or hang.
//   host_fuzz gltf|snapshot file...
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

#include "host.h"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const bool gltf = std::strcmp(argv[1], "gltf") == 0;
    int accepted = 0, rejected = 0;
    for (int i = 2; i < argc; ++i) {
        try {
            if (gltf) {
                nmr::HostMesh m = nmr::load_gltf(argv[i]);
                // touch what a caller would touch
                volatile float sink = 0.f;
                for (float v : m.positions) sink = sink + v;
                for (float v : m.normals) sink = sink + v;
                for (uint32_t v : m.indices) sink = sink + (float)v;
                (void)sink;
            } else {
                nmr::HostModel h = nmr::load_snapshot(argv[i]);
                volatile size_t sink = h.params.size() + h.density_grid.size();
                (void)sink;
            }
            ++accepted;
        } catch (const std::exception&) {
            ++rejected;
        }
    }
    std::printf("accepted %d rejected %d\n", accepted, rejected);
    return 0;
}
