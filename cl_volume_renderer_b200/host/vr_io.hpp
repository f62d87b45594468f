// vr_io.hpp — data ingest either side of the hot path: NRRD volumes and environment maps.
//
// nrrd_loader  — behaviour of app/nrrd_loader.cpp:18-151,164-198 (header ends at the first empty line; `type: short`,
//                `dimension: 3`, `endian: little`, `encoding: raw|gzip` only; `space directions` give the voxel sizes,
//                normalised by the x size), but 64-bit clean: the reference keeps byte offsets in `int`
//                (nrrd_loader.hpp:18-19) and cannot load 1024^3.  gzip payloads are inflated in a streaming
//                loop (zlib, gzip/zlib auto-detect like inflateInit2(.., 15+32) at nrrd_loader.cpp:180).
// hdre_loader  — app/hdre_loader.cpp:7-24 loads any stb_image format as RGBA8 with HDR->LDR gamma 2.2 / scale 1.0.  stb is
//                not vendored here; this loader covers Radiance .hdr (flat and new-style RLE scanlines) with stb's
//                conversion  z = pow(v * scale, 1/gamma) * 255 + 0.5, clamped and truncated, alpha 255
//                (stb_image.h:1783-1808), binary PPM (P6), and raw RGBA8 ("WxH.rgba").
// Errors print a message and exit(1) like the reference loaders (nrrd_loader.cpp:23-26,56-93, hdre_loader.cpp:15-18).
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "vr_host.hpp"

namespace vr_io_detail {
[[noreturn]] inline void die(const std::string& msg) {
  std::cerr << "Error: " << msg << '\n';
  exit(1);
}
inline std::vector<std::string> split(const std::string& s, char sep) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, sep)) out.push_back(item);
  return out;
}
inline std::string ltrim(const std::string& s) {
  size_t i = s.find_first_not_of(' ');
  return i == std::string::npos ? std::string() : s.substr(i);
}
inline std::vector<unsigned char> read_all(const std::string& path) {
  std::ifstream f(path, std::ios::in | std::ios::binary);
  if (f.fail()) die("failed to open file: " + path);
  f.seekg(0, std::ios::end);
  const std::streamoff n = f.tellg();
  f.seekg(0, std::ios::beg);
  std::vector<unsigned char> buf((size_t)n);
  f.read(reinterpret_cast<char*>(buf.data()), n);
  return buf;
}
}  // namespace vr_io_detail

struct nrrd_header {
  unsigned int x = 0, y = 0, z = 0;
  float x_voxel_size = 1, y_voxel_size = 1, z_voxel_size = 1;
  size_t data_start = 0, data_end = 0;  // 64-bit (the reference uses int)
  bool raw = true;
};

class nrrd_loader {
 public:
  volume_block load_file(const std::string path) {
    using namespace vr_io_detail;
    const std::vector<unsigned char> file = read_all(path);
    const nrrd_header h = parse_header(file);
    const size_t voxels = (size_t)h.x * h.y * h.z;
    if (voxels == 0) die("NRRD file does not declare sizes correctly.");
    std::vector<short> data(voxels);
    const size_t want = voxels * sizeof(short);
    const unsigned char* src = file.data() + h.data_start;
    const size_t avail = h.data_end - h.data_start;
    if (h.raw) {
      memcpy(data.data(), src, std::min(avail, want));  // a short file leaves zeros, like the reference's read()
    } else {
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, 15 + 32) != Z_OK) die("zlib init failed");
      unsigned char* dst = reinterpret_cast<unsigned char*>(data.data());
      size_t in_off = 0, out_off = 0;
      int rc = Z_OK;
      while (rc != Z_STREAM_END && out_off < want) {  // stream in <= 1 GiB pieces: avail_in/avail_out are 32-bit
        if (zs.avail_in == 0 && in_off < avail) {
          const size_t n = std::min<size_t>(avail - in_off, (size_t)1 << 30);
          zs.next_in = const_cast<unsigned char*>(src + in_off);
          zs.avail_in = (uInt)n;
          in_off += n;
        }
        const size_t m = std::min<size_t>(want - out_off, (size_t)1 << 30);
        zs.next_out = dst + out_off;
        zs.avail_out = (uInt)m;
        rc = inflate(&zs, Z_NO_FLUSH);
        out_off += m - zs.avail_out;
        if (rc != Z_OK && rc != Z_STREAM_END) break;
        if (rc == Z_OK && zs.avail_in == 0 && in_off >= avail && zs.avail_out != 0) break;  // truncated input
      }
      inflateEnd(&zs);
    }
    return volume_block(std::move(data), h.x, h.y, h.z, h.x_voxel_size, h.y_voxel_size, h.z_voxel_size);
  }

 private:
  static nrrd_header parse_header(const std::vector<unsigned char>& file) {
    using namespace vr_io_detail;
    nrrd_header h;
    size_t pos = 0;
    while (pos < file.size()) {
      size_t eol = pos;
      while (eol < file.size() && file[eol] != '\n') ++eol;
      std::string line(file.begin() + pos, file.begin() + eol);
      pos = eol + 1;
      if (line.empty()) break;  // header ends at the first empty line (nrrd_loader.cpp:31)
      if (line[0] == '#' || line.compare(0, 4, "NRRD") == 0) continue;
      const std::vector<std::string> kv = split(line, ':');
      if (kv.size() < 2) die("NRRD file does not declare tags correctly.");
      const std::string tag = kv[0];
      const std::vector<std::string> val = split(ltrim(kv[1]), ' ');
      if (val.empty()) die("NRRD file does not declare tags correctly.");
      if (tag == "type") {
        if (val[0] != "short") die("NRRD file not using short as type.");
      } else if (tag == "encoding") {
        if (val[0] == "gzip") h.raw = false;
        else if (val[0] == "raw") h.raw = true;
        else die("NRRD file not using gzip compression or raw.");
      } else if (tag == "endian") {
        if (val[0] != "little") die("NRRD file not using little endian format.");
      } else if (tag == "dimension") {
        if (val[0] != "3") die("NRRD file not using dimension of 3.");
      } else if (tag == "sizes") {
        if (val.size() != 3) die("NRRD file does not declare sizes correctly.");
        h.x = (unsigned)std::stoul(val[0]); h.y = (unsigned)std::stoul(val[1]); h.z = (unsigned)std::stoul(val[2]);
      } else if (tag == "space directions") {
        if (val.size() != 3) die("NRRD file does not declare space direction correctly.");
        // "(a,b,c) (d,e,f) (g,h,i)": the diagonal entries are the voxel sizes (nrrd_loader.cpp:96-109)
        const float sx = std::stof(split(val[0], ',')[0].substr(1));
        const float sy = std::stof(split(val[1], ',')[1]);
        std::string last = split(val[2], ',')[2];
        const float sz = std::stof(last.substr(0, last.size() - 1));
        h.z_voxel_size = sz / sx; h.y_voxel_size = sy / sx; h.x_voxel_size = 1.0f;
      }
    }
    h.data_start = std::min(pos, file.size());
    h.data_end = file.size();
    return h;
  }
};

class hdre_loader {
 public:
  image load_file(const std::string path) {
    using namespace vr_io_detail;
    const std::vector<unsigned char> f = read_all(path);
    if (f.size() >= 2 && f[0] == '#' && f[1] == '?') return load_radiance(f, path);
    if (f.size() >= 2 && f[0] == 'P' && f[1] == '6') return load_ppm(f, path);
    // raw RGBA8 named "<anything>.<W>x<H>.rgba"
    const size_t dot = path.rfind(".rgba");
    if (dot != std::string::npos) {
      const size_t d2 = path.rfind('.', dot - 1);
      const std::vector<std::string> wh = split(path.substr(d2 + 1, dot - d2 - 1), 'x');
      if (wh.size() == 2) {
        const unsigned w = (unsigned)std::stoul(wh[0]), h = (unsigned)std::stoul(wh[1]);
        if ((size_t)w * h * 4 == f.size()) return image(std::vector<unsigned char>(f.begin(), f.end()), w, h, 4);
      }
    }
    die("failed to load file: " + path);
  }

 private:
  static std::string next_line(const std::vector<unsigned char>& f, size_t& pos) {
    size_t eol = pos;
    while (eol < f.size() && f[eol] != '\n') ++eol;
    std::string s(f.begin() + pos, f.begin() + eol);
    pos = eol + 1;
    return s;
  }
  static image load_ppm(const std::vector<unsigned char>& f, const std::string& path) {
    size_t pos = 0;
    std::vector<long> nums;
    std::string tok;
    next_line(f, pos);  // P6
    while (nums.size() < 3 && pos < f.size()) {
      std::string l = next_line(f, pos);
      if (!l.empty() && l[0] == '#') continue;
      std::stringstream ss(l);
      long v;
      while (ss >> v) nums.push_back(v);
    }
    if (nums.size() < 3 || nums[2] != 255) vr_io_detail::die("failed to load file: " + path);
    const size_t w = nums[0], h = nums[1];
    if (pos + w * h * 3 > f.size()) vr_io_detail::die("failed to load file: " + path);
    std::vector<unsigned char> px(w * h * 4);
    for (size_t i = 0; i < w * h; ++i) {
      px[4 * i] = f[pos + 3 * i]; px[4 * i + 1] = f[pos + 3 * i + 1]; px[4 * i + 2] = f[pos + 3 * i + 2]; px[4 * i + 3] = 255;
    }
    return image(std::move(px), (unsigned)w, (unsigned)h, 4);
  }
  // stb's hdr->ldr: z = pow(v*scale, 1/gamma)*255 + 0.5, clamp, truncate (gamma 2.2, scale 1.0 — hdre_loader.cpp:11-12)
  static unsigned char to_ldr(float v) {
    float z = (float)pow(v * 1.0f, 1.0f / 2.2f) * 255 + 0.5f;
    if (z < 0) z = 0;
    if (z > 255) z = 255;
    return (unsigned char)(int)z;
  }
  static image load_radiance(const std::vector<unsigned char>& f, const std::string& path) {
    using namespace vr_io_detail;
    size_t pos = 0;
    bool fmt = false;
    for (;;) {  // header lines up to the empty line
      if (pos >= f.size()) die("failed to load file: " + path);
      const std::string l = next_line(f, pos);
      if (l.empty()) break;
      if (l == "FORMAT=32-bit_rle_rgbe") fmt = true;
    }
    if (!fmt) die("failed to load file: " + path);
    const std::string res = next_line(f, pos);  // "-Y <h> +X <w>"
    unsigned w = 0, h = 0;
    if (sscanf(res.c_str(), "-Y %u +X %u", &h, &w) != 2) die("failed to load file: " + path);
    std::vector<unsigned char> rgbe((size_t)w * h * 4);
    for (unsigned y = 0; y < h; ++y) {
      unsigned char* row = rgbe.data() + (size_t)y * w * 4;
      const bool rle = w >= 8 && w < 32768 && pos + 4 <= f.size() && f[pos] == 2 && f[pos + 1] == 2 && !(f[pos + 2] & 0x80) &&
                       ((unsigned)f[pos + 2] << 8 | f[pos + 3]) == w;
      if (!rle) {  // flat scanline
        if (pos + (size_t)w * 4 > f.size()) die("failed to load file: " + path);
        memcpy(row, f.data() + pos, (size_t)w * 4);
        pos += (size_t)w * 4;
        continue;
      }
      pos += 4;
      for (int c = 0; c < 4; ++c) {  // each channel run-length coded separately
        unsigned x = 0;
        while (x < w) {
          if (pos >= f.size()) die("failed to load file: " + path);
          unsigned count = f[pos++];
          if (count > 128) {
            count -= 128;
            if (pos >= f.size() || x + count > w) die("failed to load file: " + path);
            const unsigned char v = f[pos++];
            for (unsigned k = 0; k < count; ++k) row[(x++) * 4 + c] = v;
          } else {
            if (count == 0 || pos + count > f.size() || x + count > w) die("failed to load file: " + path);
            for (unsigned k = 0; k < count; ++k) row[(x++) * 4 + c] = f[pos++];
          }
        }
      }
    }
    std::vector<unsigned char> px((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
      const unsigned char* p = rgbe.data() + 4 * i;
      float r = 0, g = 0, b = 0;
      if (p[3] != 0) {  // stbi__hdr_convert: f1 = ldexp(1, e - (128+8))
        const float f1 = (float)ldexp(1.0f, (int)p[3] - (128 + 8));
        r = p[0] * f1; g = p[1] * f1; b = p[2] * f1;
      }
      px[4 * i] = to_ldr(r); px[4 * i + 1] = to_ldr(g); px[4 * i + 2] = to_ldr(b); px[4 * i + 3] = 255;
    }
    return image(std::move(px), w, h, 4);
  }
};
