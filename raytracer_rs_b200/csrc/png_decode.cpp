// png_decode.cpp — minimal PNG -> RGB8 decoder (zlib inflate + scanline unfiltering).
//
// Stands in for `image::open(path)?.to_rgb8()` in raytracer_lib/src/scene/texture.rs:36: the result is the
// decoded 8-bit RGB raster, rows top to bottom, with gAMA/sRGB/iCCP chunks ignored (the `image` crate does not
// apply them either). Supported: bit depth 8 (grey, grey+alpha, RGB, RGBA, palette) and 16 (high byte kept),
// non-interlaced. Anything else is reported as an error string like the reference's TextureLoadError.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace rtb {

static uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

static inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}

bool decode_png_rgb8(const std::string& path, uint32_t* width, uint32_t* height, std::vector<uint8_t>* rgb, std::string* err) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        *err = "No such file or directory (os error 2): " + path;
        return false;
    }
    std::vector<uint8_t> file;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + n);
    std::fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) {
        *err = "Format error decoding Png: Invalid PNG signature.";
        return false;
    }
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, palette;
    size_t p = 8;
    bool seen_end = false;
    while (p + 12 <= file.size() && !seen_end) {
        uint32_t len = be32(&file[p]);
        const uint8_t* type = &file[p + 4];
        const uint8_t* data = &file[p + 8];
        if (p + 12 + (size_t)len > file.size()) break;
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            w = be32(data);
            h = be32(data + 4);
            depth = data[8];
            ctype = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            seen_end = true;
        }
        p += 12 + (size_t)len;
    }
    if (w == 0 || h == 0 || ctype < 0) {
        *err = "Format error decoding Png: missing IHDR";
        return false;
    }
    if (interlace != 0 || (depth != 8 && depth != 16)) {
        *err = "The decoder for Png does not support the format features: interlaced or sub-byte bit depth";
        return false;
    }
    int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (channels == 0 || (ctype == 3 && depth != 8)) {
        *err = "Format error decoding Png: invalid colour type";
        return false;
    }
    // the IHDR is untrusted input: bound the dimensions so that none of the size products below can wrap or ask for terabytes
    if (w > (1u << 15) || h > (1u << 15)) {
        *err = "The decoder for Png does not support the format features: image larger than 32768 x 32768";
        return false;
    }
    const size_t bpp = (size_t)channels * (depth / 8);
    const size_t stride = (size_t)w * bpp;
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf out_len = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || out_len != raw.size()) {
        *err = "Format error decoding Png: corrupt deflate stream";
        return false;
    }
    std::vector<uint8_t> img(stride * (size_t)h);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* in = &raw[(stride + 1) * y];
        const int filter = in[0];
        ++in;
        uint8_t* cur = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= bpp ? cur[i - bpp] : 0;
            int b = up ? up[i] : 0;
            int c = (up && i >= bpp) ? up[i - bpp] : 0;
            int x = in[i];
            switch (filter) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) >> 1; break;
                case 4: x += paeth(a, b, c); break;
                default:
                    *err = "Format error decoding Png: unknown filter type";
                    return false;
            }
            cur[i] = (uint8_t)x;
        }
    }
    rgb->resize((size_t)w * h * 3);
    const size_t step = depth / 8;  // for 16-bit samples the high byte comes first
    // 16-bit samples -> 8 bit the way `image` 0.25's to_rgb8 does (FromPrimitive<u16> for u8: (c + 128) / 257, i.e.
    // round(c * 255 / 65535)); the crate is not vendored under /root/reference, so this is its published rule, unpinned
    auto sample8 = [&](const uint8_t* sp) -> uint8_t {
        if (step == 1) return sp[0];
        const uint32_t c = ((uint32_t)sp[0] << 8) | sp[1];
        return (uint8_t)((c + 128u) / 257u);
    };
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        const uint8_t* px = &img[i * bpp];
        uint8_t r, g, b;
        if (ctype == 3) {
            size_t k = (size_t)px[0] * 3;
            if (k + 2 >= palette.size()) {
                *err = "Format error decoding Png: palette index out of range";
                return false;
            }
            r = palette[k];
            g = palette[k + 1];
            b = palette[k + 2];
        } else if (ctype == 0 || ctype == 4) {
            r = g = b = sample8(px);
        } else {
            r = sample8(px);
            g = sample8(px + step);
            b = sample8(px + 2 * step);
        }
        (*rgb)[3 * i] = r;
        (*rgb)[3 * i + 1] = g;
        (*rgb)[3 * i + 2] = b;
    }
    *width = w;
    *height = h;
    return true;
}

}  // namespace rtb
