// Reference pin for the on-disk map format.  TEST INFRASTRUCTURE ONLY (built into oracle/_ref/ply_ref).
//
// Drives the reference's OWN PLY writer -- tinyply, compiled from where it lies (/root/reference/src/tinyply.h) -- with the
// same sequence of add_properties_to_element calls and the same tensor preparation (transpose(1,2).flatten(1)) that
// GaussianModel::saveMap performs (/root/reference/src/gaussian.cpp:305-397), on arrays read from a raw float32 blob:
//     uint32 P, uint32 M, xyz[P*3], f_dc[P*1*3], f_rest[P*M*3], opacity[P], scale[P*3], rotation[P*4]
// usage: ply_ref <in.blob> <out.ply>
#define TINYPLY_IMPLEMENTATION
#include "tinyply.h"

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

static std::vector<float> take(std::ifstream& in, size_t n) {
    std::vector<float> v(n);
    in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(float)));
    return v;
}

static std::vector<std::string> names(const std::string& prefix, size_t n) {
    std::vector<std::string> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = prefix + std::to_string(i);
    return v;
}

int main(int argc, char** argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: ply_ref in.blob out.ply\n"); return 2; }
    std::ifstream in(argv[1], std::ios::binary);
    uint32_t P = 0, M = 0;
    in.read(reinterpret_cast<char*>(&P), 4);
    in.read(reinterpret_cast<char*>(&M), 4);
    std::vector<float> xyz = take(in, (size_t)P * 3), dc = take(in, (size_t)P * 3), rest = take(in, (size_t)P * M * 3);
    std::vector<float> opacity = take(in, P), scale = take(in, (size_t)P * 3), rotation = take(in, (size_t)P * 4);
    if (!in) { std::fprintf(stderr, "short blob\n"); return 3; }
    // features_*.transpose(1, 2).flatten(1): [P, K, 3] -> [P, 3, K]
    std::vector<float> f_dc((size_t)P * 3), f_rest((size_t)P * 3 * M);
    for (uint32_t i = 0; i < P; ++i)
        for (int c = 0; c < 3; ++c) {
            f_dc[(size_t)i * 3 + c] = dc[(size_t)i * 3 + c];
            for (uint32_t k = 0; k < M; ++k) f_rest[((size_t)i * 3 + c) * M + k] = rest[((size_t)i * M + k) * 3 + c];
        }
    std::filebuf fb;
    fb.open(argv[2], std::ios::out | std::ios::binary);
    std::ostream os(&fb);
    tinyply::PlyFile file;
    auto add = [&](const std::vector<std::string>& n, std::vector<float>& data) {
        file.add_properties_to_element("vertex", n, tinyply::Type::FLOAT32, P, reinterpret_cast<uint8_t*>(data.data()),
                                       tinyply::Type::INVALID, 0);
    };
    add({"x", "y", "z"}, xyz);
    add(names("f_dc_", 3), f_dc);
    add(names("f_rest_", (size_t)3 * M), f_rest);
    add({"opacity"}, opacity);
    add(names("scale_", 3), scale);
    add(names("rot_", 4), rotation);
    file.write(os, true);
    fb.close();
    return 0;
}
