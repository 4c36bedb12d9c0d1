// cr_image.h -- base64 and PNG decoding for embedded glTF payloads.
// The reference goes through tinygltf + stb_image with req_comp = 4 (support/tinygltf), i.e. every
// image reaches MulticamScene::addImage (libEyeRenderer3/MulticamScene.cpp:753-798) as 4-channel
// 8-bit RGBA, row 0 first.  Only the formats the shipped scenes use are implemented here:
// non-interlaced PNG, bit depth 8 (grey, grey+alpha, RGB, RGBA, palette) and 16 (reduced to the
// high byte), and baseline + progressive JPEG (cr_jpeg.h).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <zlib.h>

namespace cr {

inline std::vector<uint8_t> base64Decode(const char* s, size_t n)
{
    static int8_t lut[256];
    static bool init = false;
    if (!init) {
        memset(lut, -1, sizeof lut);
        const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) lut[static_cast<uint8_t>(a[i])] = static_cast<int8_t>(i);
        lut[static_cast<uint8_t>('-')] = 62;   // url-safe alphabet accepted too
        lut[static_cast<uint8_t>('_')] = 63;
        init = true;
    }
    std::vector<uint8_t> out;
    out.reserve(n / 4 * 3 + 3);
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = 0; i < n; i++) {
        int8_t v = lut[static_cast<uint8_t>(s[i])];
        if (v < 0) continue;   // padding, whitespace
        acc = (acc << 6) | static_cast<uint32_t>(v);
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back(static_cast<uint8_t>((acc >> bits) & 0xFF));
        }
    }
    return out;
}

struct ImageRGBA8 {
    int width = 0, height = 0;
    std::vector<uint8_t> pixels;   // width*height*4, row 0 first
};

inline uint32_t be32(const uint8_t* p)
{ return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]); }

inline ImageRGBA8 decodePNG(const uint8_t* data, size_t size)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || memcmp(data, sig, 8) != 0) throw std::runtime_error("not a PNG stream");
    size_t pos = 8;
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool haveHdr = false;
    while (pos + 12 <= size) {
        uint32_t len = be32(data + pos);
        const uint8_t* tag = data + pos + 4;
        const uint8_t* body = data + pos + 8;
        if (pos + 12 + len > size) throw std::runtime_error("truncated PNG chunk");
        if (!memcmp(tag, "IHDR", 4)) {
            W = be32(body); H = be32(body + 4);
            depth = body[8]; ctype = body[9]; interlace = body[12];
            haveHdr = true;
        } else if (!memcmp(tag, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(tag, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(tag, "IEND", 4)) break;
        pos += 12 + len;
    }
    if (!haveHdr) throw std::runtime_error("PNG without IHDR");
    if (interlace) throw std::runtime_error("interlaced PNG not supported");
    if (depth != 8 && depth != 16) throw std::runtime_error("PNG bit depth not supported (8/16 only)");
    int channels = 0;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: throw std::runtime_error("bad PNG colour type");
    }
    if (ctype == 3 && depth != 8) throw std::runtime_error("palette PNG must be 8-bit here");
    const size_t bpp = static_cast<size_t>(channels) * (depth / 8);
    const size_t stride = bpp * W;
    std::vector<uint8_t> raw((stride + 1) * H);
    uLongf rawLen = static_cast<uLongf>(raw.size());
    int zr = uncompress(raw.data(), &rawLen, idat.data(), static_cast<uLong>(idat.size()));
    if (zr != Z_OK || rawLen != raw.size()) throw std::runtime_error("PNG inflate failed");
    // unfilter in place
    std::vector<uint8_t> img(stride * H);
    for (uint32_t y = 0; y < H; y++) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* in = &raw[(stride + 1) * y + 1];
        uint8_t* out = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= bpp ? out[x - bpp] : 0;
            const int b = up ? up[x] : 0;
            const int c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v = in[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c;
                    const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw std::runtime_error("bad PNG filter");
            }
            out[x] = static_cast<uint8_t>(v);
        }
    }
    ImageRGBA8 res;
    res.width = static_cast<int>(W);
    res.height = static_cast<int>(H);
    res.pixels.resize(static_cast<size_t>(W) * H * 4);
    const size_t step = depth / 8;   // 16-bit: big-endian, keep the high byte (stb's 16->8 reduction)
    for (size_t i = 0; i < static_cast<size_t>(W) * H; i++) {
        const uint8_t* px = &img[i * bpp];
        uint8_t* o = &res.pixels[i * 4];
        switch (ctype) {
            case 0: o[0] = o[1] = o[2] = px[0]; o[3] = 255; break;
            case 2: o[0] = px[0]; o[1] = px[step]; o[2] = px[2 * step]; o[3] = 255; break;
            case 3: {
                const size_t k = px[0];
                if (3 * k + 2 < plte.size()) { o[0] = plte[3 * k]; o[1] = plte[3 * k + 1]; o[2] = plte[3 * k + 2]; }
                else { o[0] = o[1] = o[2] = 0; }
                o[3] = k < trns.size() ? trns[k] : 255;
                break;
            }
            case 4: o[0] = o[1] = o[2] = px[0]; o[3] = px[step]; break;
            default: o[0] = px[0]; o[1] = px[step]; o[2] = px[2 * step]; o[3] = px[3 * step]; break;
        }
    }
    return res;
}

ImageRGBA8 decodeJPEG(const uint8_t* data, size_t size);   // cr_jpeg.h

inline ImageRGBA8 decodeImage(const uint8_t* data, size_t size)
{
    if (size >= 2 && data[0] == 0xFF && data[1] == 0xD8) return decodeJPEG(data, size);
    return decodePNG(data, size);
}

}  // namespace cr
