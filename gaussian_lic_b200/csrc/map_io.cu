// map_io.cu -- on-disk map format of Gaussian-LIC: binary little-endian PLY, one "vertex" element,
// float properties x y z | f_dc_0..2 | f_rest_0..3M-1 | opacity | scale_0..2 | rot_0..3, exactly the file
// GaussianModel::saveMap produces through tinyply (/root/reference/src/gaussian.cpp:305-397; header text
// tinyply.h:664-703, row-interleaved body tinyply.h:588-623) and that 3DGS viewers read.  Values are the RAW
// parameters (opacity logit, log-scale, un-normalised quaternion); f_rest is stored channel-major
// (features_rest.transpose(1,2).flatten(1): f_rest_{c*M+k} = sh[k][c]).  Host-only code: pointers are HOST pointers.
#include "common.cuh"

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace glic {
namespace {

std::string ply_header(uint32_t P, uint32_t M) {
    std::string h = "ply\nformat binary_little_endian 1.0\nelement vertex " + std::to_string(P) + "\n";
    auto prop = [&h](const std::string& name) { h += "property float " + name + "\n"; };
    prop("x"); prop("y"); prop("z");
    for (int i = 0; i < 3; ++i) prop("f_dc_" + std::to_string(i));
    for (uint32_t i = 0; i < 3 * M; ++i) prop("f_rest_" + std::to_string(i));
    prop("opacity");
    for (int i = 0; i < 3; ++i) prop("scale_" + std::to_string(i));
    for (int i = 0; i < 4; ++i) prop("rot_" + std::to_string(i));
    h += "end_header\n";
    return h;
}

inline size_t row_floats(uint32_t M) { return 3 + 3 + 3 * (size_t)M + 1 + 3 + 4; }

struct File {
    FILE* f;
    explicit File(const char* path, const char* mode) : f(path ? std::fopen(path, mode) : nullptr) {}
    ~File() { if (f) std::fclose(f); }
};

// one row of the file <-> the six parameter arrays (write = false: file row -> arrays)
template <bool WRITE>
inline void row_xfer(float* row, uint32_t i, uint32_t M, float* xyz, float* f_dc, float* f_rest, float* opacity, float* scale,
                     float* rotation) {
    auto mv = [](float& file_v, float& mem_v) { if (WRITE) file_v = mem_v; else mem_v = file_v; };
    float* r = row;
    for (int c = 0; c < 3; ++c) mv(*r++, xyz[3 * (size_t)i + c]);
    for (int c = 0; c < 3; ++c) mv(*r++, f_dc[3 * (size_t)i + c]);                       // [P,1,3] -> transpose -> same order
    for (int c = 0; c < 3; ++c)
        for (uint32_t k = 0; k < M; ++k) mv(*r++, f_rest[((size_t)i * M + k) * 3 + c]);  // channel-major
    mv(*r++, opacity[i]);
    for (int c = 0; c < 3; ++c) mv(*r++, scale[3 * (size_t)i + c]);
    for (int c = 0; c < 4; ++c) mv(*r++, rotation[4 * (size_t)i + c]);
}

}  // namespace
}  // namespace glic

using namespace glic;

extern "C" {

size_t glic_ply_bytes(uint32_t P, uint32_t M) { return ply_header(P, M).size() + (size_t)P * row_floats(M) * sizeof(float); }

int glic_ply_write(const char* path, uint32_t P, uint32_t M, const float* xyz, const float* f_dc, const float* f_rest,
                   const float* opacity, const float* scale, const float* rotation) {
    if (!path) { set_error("ply_write: null path"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (P > 0 && (!xyz || !f_dc || (M > 0 && !f_rest) || !opacity || !scale || !rotation)) { set_error("ply_write: null array"); return GLIC_ERR_INVALID_ARGUMENT; }
    File out(path, "wb");
    if (!out.f) { set_error(std::string("ply_write: cannot open ") + path); return GLIC_ERR_INVALID_ARGUMENT; }
    const std::string h = ply_header(P, M);
    if (std::fwrite(h.data(), 1, h.size(), out.f) != h.size()) { set_error("ply_write: short write"); return GLIC_ERR_INVALID_ARGUMENT; }
    const size_t rf = row_floats(M);
    const uint32_t CHUNK = 4096;
    std::vector<float> buf((size_t)CHUNK * rf);
    for (uint32_t i0 = 0; i0 < P; i0 += CHUNK) {
        const uint32_t n = std::min(CHUNK, P - i0);
        for (uint32_t i = 0; i < n; ++i)
            row_xfer<true>(buf.data() + (size_t)i * rf, i0 + i, M, const_cast<float*>(xyz), const_cast<float*>(f_dc),
                           const_cast<float*>(f_rest), const_cast<float*>(opacity), const_cast<float*>(scale),
                           const_cast<float*>(rotation));
        if (std::fwrite(buf.data(), sizeof(float), (size_t)n * rf, out.f) != (size_t)n * rf) { set_error("ply_write: short write"); return GLIC_ERR_INVALID_ARGUMENT; }
    }
    return GLIC_OK;
}

// Parses the header of a map file written by glic_ply_write / GaussianModel::saveMap; *data_offset = first body byte.
int glic_ply_read_header(const char* path, uint32_t* P, uint32_t* M, size_t* data_offset) {
    if (!path || !P || !M) { set_error("ply_read_header: null argument"); return GLIC_ERR_INVALID_ARGUMENT; }
    File in(path, "rb");
    if (!in.f) { set_error(std::string("ply_read_header: cannot open ") + path); return GLIC_ERR_INVALID_ARGUMENT; }
    char line[256];
    bool ply = false, fmt = false, end = false;
    long long count = -1;
    std::vector<std::string> props;
    while (std::fgets(line, sizeof(line), in.f)) {
        std::string s(line);
        while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
        if (s == "ply") ply = true;
        else if (s == "format binary_little_endian 1.0") fmt = true;
        else if (s.rfind("element vertex ", 0) == 0) count = std::atoll(s.c_str() + 15);
        else if (s.rfind("property float ", 0) == 0) props.push_back(s.substr(15));
        else if (s.rfind("comment", 0) == 0) continue;
        else if (s == "end_header") { end = true; break; }
        else { set_error("ply_read_header: unsupported header line: " + s); return GLIC_ERR_INVALID_ARGUMENT; }
    }
    if (!ply || !fmt || !end || count < 0) { set_error("ply_read_header: not a binary little-endian Gaussian map"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (props.size() < 14 || (props.size() - 14) % 3 != 0) { set_error("ply_read_header: unexpected property count"); return GLIC_ERR_INVALID_ARGUMENT; }
    const uint32_t m = (uint32_t)((props.size() - 14) / 3);
    // the property list must be exactly the one saveMap writes, in its order
    const std::string want = ply_header((uint32_t)count, m);
    std::string got = "ply\nformat binary_little_endian 1.0\nelement vertex " + std::to_string(count) + "\n";
    for (const auto& p : props) got += "property float " + p + "\n";
    got += "end_header\n";
    if (got != want) { set_error("ply_read_header: property names / order differ from the Gaussian-LIC map layout"); return GLIC_ERR_INVALID_ARGUMENT; }
    *P = (uint32_t)count; *M = m;
    if (data_offset) *data_offset = (size_t)std::ftell(in.f);
    return GLIC_OK;
}

int glic_ply_read(const char* path, uint32_t P, uint32_t M, float* xyz, float* f_dc, float* f_rest, float* opacity, float* scale,
                  float* rotation) {
    uint32_t p = 0, m = 0;
    size_t off = 0;
    if (int e = glic_ply_read_header(path, &p, &m, &off)) return e;
    if (p != P || m != M) { set_error("ply_read: P / M do not match the file"); return GLIC_ERR_INVALID_ARGUMENT; }
    if (P > 0 && (!xyz || !f_dc || (M > 0 && !f_rest) || !opacity || !scale || !rotation)) { set_error("ply_read: null array"); return GLIC_ERR_INVALID_ARGUMENT; }
    File in(path, "rb");
    if (!in.f || std::fseek(in.f, (long)off, SEEK_SET) != 0) { set_error("ply_read: cannot seek"); return GLIC_ERR_INVALID_ARGUMENT; }
    const size_t rf = row_floats(M);
    const uint32_t CHUNK = 4096;
    std::vector<float> buf((size_t)CHUNK * rf);
    for (uint32_t i0 = 0; i0 < P; i0 += CHUNK) {
        const uint32_t n = std::min(CHUNK, P - i0);
        if (std::fread(buf.data(), sizeof(float), (size_t)n * rf, in.f) != (size_t)n * rf) { set_error("ply_read: truncated file"); return GLIC_ERR_INVALID_ARGUMENT; }
        for (uint32_t i = 0; i < n; ++i) row_xfer<false>(buf.data() + (size_t)i * rf, i0 + i, M, xyz, f_dc, f_rest, opacity, scale, rotation);
    }
    return GLIC_OK;
}

// Convenience for the packed model (glic_packed_offsets layout), HOST copy of the planar parameter buffer.
int glic_ply_write_packed(const char* path, uint32_t P, uint32_t M, const float* params_host) {
    if (P > 0 && !params_host) { set_error("ply_write_packed: null buffer"); return GLIC_ERR_INVALID_ARGUMENT; }
    const float* rot = params_host;
    const float* xyz = rot + 4 * (size_t)P;
    const float* scl = xyz + 3 * (size_t)P;
    const float* opa = scl + 3 * (size_t)P;
    const float* dc = opa + (size_t)P;
    const float* rest = dc + 3 * (size_t)P;
    return glic_ply_write(path, P, M, xyz, dc, rest, opa, scl, rot);
}

}  // extern "C"
