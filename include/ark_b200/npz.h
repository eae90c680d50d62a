// npz.h -- minimal reader for numpy .npz archives (zip, stored or deflate) holding little-endian
// f4/f8/i4/u4/i8/u8 arrays in C order.  Stands in for the reference's cnpy + UtilCnpy (cnpy.cpp:246-300,
// Util.cpp:249-309) in the C++ facade; header-only, needs zlib.
#pragma once
#include <zlib.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ark_b200 {

struct NpyArray {
    std::vector<size_t> shape;
    std::string descr;            // e.g. "<f4"
    std::vector<unsigned char> raw;
    size_t count() const { size_t n = 1; for (size_t s : shape) n *= s; return n; }
    template <class T> T as(size_t i) const {
        const unsigned char* p = raw.data();
        const char k = descr[1];
        const int w = descr[2] - '0';
        if (k == 'f' && w == 4) { float v; std::memcpy(&v, p + 4 * i, 4); return (T)v; }
        if (k == 'f' && w == 8) { double v; std::memcpy(&v, p + 8 * i, 8); return (T)v; }
        if (k == 'i' && w == 4) { int32_t v; std::memcpy(&v, p + 4 * i, 4); return (T)v; }
        if (k == 'u' && w == 4) { uint32_t v; std::memcpy(&v, p + 4 * i, 4); return (T)v; }
        if (k == 'i' && w == 8) { int64_t v; std::memcpy(&v, p + 8 * i, 8); return (T)v; }
        if (k == 'u' && w == 8) { uint64_t v; std::memcpy(&v, p + 8 * i, 8); return (T)v; }
        throw std::runtime_error("npz: unsupported dtype " + descr);
    }
    std::vector<double> to_double() const { std::vector<double> o(count()); for (size_t i = 0; i < o.size(); ++i) o[i] = as<double>(i); return o; }
};

inline NpyArray parse_npy(const std::vector<unsigned char>& buf) {
    if (buf.size() < 10 || std::memcmp(buf.data(), "\x93NUMPY", 6) != 0) throw std::runtime_error("npz: bad npy magic");
    const int major = buf[6];
    size_t hlen, hoff;
    if (major == 1) { hlen = buf[8] | (buf[9] << 8); hoff = 10; }
    else { hlen = buf[8] | (buf[9] << 8) | (buf[10] << 16) | ((size_t)buf[11] << 24); hoff = 12; }
    const std::string h(reinterpret_cast<const char*>(buf.data() + hoff), hlen);
    NpyArray a;
    size_t p = h.find("'descr'");
    p = h.find('\'', h.find(':', p));
    a.descr = h.substr(p + 1, h.find('\'', p + 1) - p - 1);
    if (a.descr.size() != 3 || (a.descr[0] != '<' && a.descr[0] != '|')) throw std::runtime_error("npz: unsupported descr " + a.descr);
    if (h.find("'fortran_order': True") != std::string::npos) throw std::runtime_error("npz: fortran order unsupported");
    p = h.find('(', h.find("'shape'"));
    const size_t e = h.find(')', p);
    std::string s = h.substr(p + 1, e - p - 1);
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && (s[i] < '0' || s[i] > '9')) ++i;
        if (i >= s.size()) break;
        size_t v = 0;
        while (i < s.size() && s[i] >= '0' && s[i] <= '9') v = v * 10 + (s[i++] - '0');
        a.shape.push_back(v);
    }
    a.raw.assign(buf.begin() + hoff + hlen, buf.end());
    const size_t need = a.count() * (size_t)(a.descr[2] - '0');
    if (a.raw.size() < need) throw std::runtime_error("npz: truncated array");
    return a;
}

inline std::map<std::string, NpyArray> load_npz(const std::string& path) {
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) throw std::runtime_error("npz: cannot open " + path);
    std::fseek(fp, 0, SEEK_END);
    const long fsz = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::vector<unsigned char> f((size_t)fsz);
    if (std::fread(f.data(), 1, f.size(), fp) != f.size()) { std::fclose(fp); throw std::runtime_error("npz: short read"); }
    std::fclose(fp);
    auto u16 = [&](size_t o) { return (uint32_t)f[o] | ((uint32_t)f[o + 1] << 8); };
    auto u32 = [&](size_t o) { return (uint64_t)u16(o) | ((uint64_t)u16(o + 2) << 16); };
    auto u64 = [&](size_t o) { return u32(o) | (u32(o + 4) << 32); };
    long eocd = -1;
    for (long o = fsz - 22; o >= 0 && o >= fsz - 22 - 65536; --o)
        if (u32(o) == 0x06054b50) { eocd = o; break; }
    if (eocd < 0) throw std::runtime_error("npz: no end-of-central-directory record");
    uint64_t n = u16(eocd + 10), cdoff = u32(eocd + 16);
    if (cdoff == 0xFFFFFFFFull || n == 0xFFFF) {  // zip64
        const long loc = eocd - 20;
        if (loc < 0 || u32(loc) != 0x07064b50) throw std::runtime_error("npz: zip64 locator missing");
        const uint64_t e64 = u64(loc + 8);
        n = u64(e64 + 32);
        cdoff = u64(e64 + 48);
    }
    std::map<std::string, NpyArray> out;
    size_t o = cdoff;
    for (uint64_t k = 0; k < n; ++k) {
        if (u32(o) != 0x02014b50) throw std::runtime_error("npz: bad central directory");
        const uint32_t method = u16(o + 10), nl = u16(o + 28), xl = u16(o + 30), cl = u16(o + 32);
        uint64_t csz = u32(o + 20), usz = u32(o + 24), lho = u32(o + 42);
        std::string name(reinterpret_cast<const char*>(&f[o + 46]), nl);
        size_t x = o + 46 + nl;
        const size_t xe = x + xl;
        while (x + 4 <= xe) {
            const uint32_t id = u16(x), sz = u16(x + 2);
            if (id == 0x0001) {
                size_t q = x + 4;
                if (usz == 0xFFFFFFFFull) { usz = u64(q); q += 8; }
                if (csz == 0xFFFFFFFFull) { csz = u64(q); q += 8; }
                if (lho == 0xFFFFFFFFull) { lho = u64(q); q += 8; }
            }
            x += 4 + sz;
        }
        o = xe + cl;
        if (u32(lho) != 0x04034b50) throw std::runtime_error("npz: bad local header");
        const size_t dstart = lho + 30 + u16(lho + 26) + u16(lho + 28);
        std::vector<unsigned char> buf((size_t)usz);
        if (method == 0) {
            std::memcpy(buf.data(), &f[dstart], (size_t)usz);
        } else if (method == 8) {
            z_stream zs;
            std::memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -MAX_WBITS) != Z_OK) throw std::runtime_error("npz: inflateInit2");
            zs.next_in = &f[dstart]; zs.avail_in = (uInt)csz;
            zs.next_out = buf.data(); zs.avail_out = (uInt)usz;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END) throw std::runtime_error("npz: inflate failed for " + name);
        } else {
            throw std::runtime_error("npz: unsupported compression");
        }
        if (name.size() > 4 && name.substr(name.size() - 4) == ".npy") name = name.substr(0, name.size() - 4);
        out[name] = parse_npy(buf);
    }
    return out;
}
}  // namespace ark_b200
