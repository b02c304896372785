// pcd_io.hpp - PCD v0.7 reader (ascii / binary) and binary writers, dependency-free.
//
// Stands in for pcl::io::loadPCDFile<pcl::PointXYZI> / savePCDFileBinary as used by the reference harness
// (DCReg/src/icp_test_runner.cpp:156-176, 362-381).  Shipped clouds are
// `FIELDS x y z intensity`, `SIZE 4 4 4 4`, `TYPE F F F F`, `DATA binary`.
#pragma once
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pcd {

struct Cloud {
    std::vector<float> xyzi;          // 4 floats per point (x, y, z, intensity)
    size_t size() const { return xyzi.size() / 4; }
    bool empty() const { return xyzi.empty(); }
};

inline bool load(const std::string& path, Cloud& out, std::string* err = nullptr) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) { if (err) *err = "cannot open " + path; return false; }
    std::vector<std::string> fields, types;
    std::vector<int> sizes, counts;
    size_t npoints = 0;
    std::string data_kind, line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string tag;
        ss >> tag;
        if (tag == "FIELDS") { std::string s; while (ss >> s) fields.push_back(s); }
        else if (tag == "SIZE") { int v; while (ss >> v) sizes.push_back(v); }
        else if (tag == "TYPE") { std::string s; while (ss >> s) types.push_back(s); }
        else if (tag == "COUNT") { int v; while (ss >> v) counts.push_back(v); }
        else if (tag == "POINTS") { ss >> npoints; }
        else if (tag == "DATA") { ss >> data_kind; break; }
    }
    if (fields.empty() || fields.size() != sizes.size() || fields.size() != types.size()) { if (err) *err = "bad PCD header: " + path; return false; }
    if (counts.empty()) counts.assign(fields.size(), 1);
    int ix = -1, iy = -1, iz = -1, ii = -1;
    std::vector<size_t> offs(fields.size());
    size_t stride = 0;
    for (size_t k = 0; k < fields.size(); ++k) {
        offs[k] = stride;
        stride += (size_t)sizes[k] * counts[k];
        if (fields[k] == "x") ix = (int)k;
        if (fields[k] == "y") iy = (int)k;
        if (fields[k] == "z") iz = (int)k;
        if (fields[k] == "intensity") ii = (int)k;
    }
    if (ix < 0 || iy < 0 || iz < 0) { if (err) *err = "PCD without x y z fields: " + path; return false; }
    for (int k : {ix, iy, iz})
        if (types[k] != "F" || sizes[k] != 4) { if (err) *err = "x y z must be float32: " + path; return false; }
    out.xyzi.assign(npoints * 4, 0.f);
    if (data_kind == "binary") {
        std::vector<char> buf(stride * npoints);
        f.read(buf.data(), (std::streamsize)buf.size());
        if ((size_t)f.gcount() != buf.size()) { if (err) *err = "truncated PCD: " + path; return false; }
        for (size_t p = 0; p < npoints; ++p) {
            const char* rec = buf.data() + p * stride;
            std::memcpy(&out.xyzi[4 * p + 0], rec + offs[ix], 4);
            std::memcpy(&out.xyzi[4 * p + 1], rec + offs[iy], 4);
            std::memcpy(&out.xyzi[4 * p + 2], rec + offs[iz], 4);
            if (ii >= 0 && types[ii] == "F" && sizes[ii] == 4) std::memcpy(&out.xyzi[4 * p + 3], rec + offs[ii], 4);
        }
    } else if (data_kind == "ascii") {
        for (size_t p = 0; p < npoints; ++p) {
            if (!std::getline(f, line)) { if (err) *err = "truncated PCD: " + path; return false; }
            std::istringstream ss(line);
            size_t col = 0;
            for (size_t k = 0; k < fields.size(); ++k)
                for (int c = 0; c < counts[k]; ++c, ++col) {
                    double v; ss >> v;
                    if ((int)k == ix) out.xyzi[4 * p + 0] = (float)v;
                    if ((int)k == iy) out.xyzi[4 * p + 1] = (float)v;
                    if ((int)k == iz) out.xyzi[4 * p + 2] = (float)v;
                    if ((int)k == ii) out.xyzi[4 * p + 3] = (float)v;
                }
        }
    } else { if (err) *err = "unsupported PCD DATA '" + data_kind + "': " + path; return false; }
    return true;
}

// x y z intensity, float32, DATA binary (what savePCDFileBinary writes for pcl::PointXYZI)
inline bool save_xyzi_binary(const std::string& path, const float* xyzi, size_t n) {
    std::ofstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n"
      << "COUNT 1 1 1 1\nWIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
    f.write(reinterpret_cast<const char*>(xyzi), (std::streamsize)(n * 16));
    return (bool)f;
}

// x y z rgb (packed float), DATA binary (pcl::PointXYZRGB clouds of saveAlignedClouds / saveErrorPointCloud)
inline bool save_xyzrgb_binary(const std::string& path, const std::vector<float>& xyz, const std::vector<uint32_t>& rgb) {
    const size_t n = rgb.size();
    std::ofstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\n"
      << "COUNT 1 1 1 1\nWIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
    for (size_t i = 0; i < n; ++i) {
        f.write(reinterpret_cast<const char*>(&xyz[3 * i]), 12);
        f.write(reinterpret_cast<const char*>(&rgb[i]), 4);
    }
    return (bool)f;
}

}  // namespace pcd
