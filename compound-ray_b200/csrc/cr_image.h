// cr_image.h -- base64 and PNG decoding for embedded glTF payloads.
// The reference goes through tinygltf + stb_image with req_comp = 4 (support/tinygltf), i.e. every
// image reaches MulticamScene::addImage (libEyeRenderer3/MulticamScene.cpp:753-798) as 4-channel
// 8-bit RGBA, row 0 first.  PNG: every colour type and bit depth of the standard incl. tRNS and Adam7
// (16-bit reduced to the high byte); JPEG: baseline + progressive Huffman streams (cr_jpeg.h).
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

// Undoes the PNG row filters of one (sub)image of pw x ph pixels whose scanlines start at `src`
// (filter byte + rowBytes each), and appends one byte (depth <= 8, raw unscaled value) or two bytes
// (depth 16, big-endian) per sample to `out`.  Returns the number of source bytes consumed.
inline size_t pngUnfilterPass(const uint8_t* src, uint32_t pw, uint32_t ph, int channels, int depth, std::vector<uint8_t>& out)
{
    const size_t bitsPerPixel = static_cast<size_t>(channels) * static_cast<size_t>(depth);
    const size_t rowBytes = (static_cast<size_t>(pw) * bitsPerPixel + 7) / 8;
    const size_t bpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;          // filter distance in bytes (PNG spec 9.2)
    std::vector<uint8_t> prev(rowBytes, 0), cur(rowBytes, 0);
    for (uint32_t y = 0; y < ph; y++) {
        const uint8_t ft = src[(rowBytes + 1) * y];
        const uint8_t* in = src + (rowBytes + 1) * y + 1;
        for (size_t x = 0; x < rowBytes; x++) {
            const int a = x >= bpp ? cur[x - bpp] : 0;
            const int b = prev[x];
            const int c = x >= bpp ? prev[x - bpp] : 0;
            int v = in[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    const int pp = a + b - c;
                    const int pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw std::runtime_error("bad PNG filter");
            }
            cur[x] = static_cast<uint8_t>(v);
        }
        if (depth >= 8) {
            out.insert(out.end(), cur.begin(), cur.end());
        } else {                                                           // 1/2/4-bit samples, most significant bits first
            const int mask = (1 << depth) - 1;
            const size_t nSamples = static_cast<size_t>(pw) * static_cast<size_t>(channels);
            for (size_t k = 0; k < nSamples; k++) {
                const size_t bit = k * static_cast<size_t>(depth);
                out.push_back(static_cast<uint8_t>((cur[bit >> 3] >> (8 - depth - static_cast<int>(bit & 7))) & mask));
            }
        }
        prev.swap(cur);
    }
    return (rowBytes + 1) * ph;
}

// PNG -> RGBA8 with the conventions of stb_image's 8-bit API (stbi_load, req_comp = 4), which is what the
// reference's loader hands to the renderer: bit depths 1/2/4/8/16 (16 reduced to the high byte, sub-byte
// greys scaled to 0..255), grey / grey+alpha / RGB / RGBA / palette, tRNS for palette, grey and RGB
// (colour-key transparency is matched before the 16 -> 8 reduction), Adam7 interlacing.
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
            if (len < 13) throw std::runtime_error("bad IHDR");
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
    if (W == 0 || H == 0) throw std::runtime_error("empty PNG");
    if (W > (1u << 24) || H > (1u << 24) || static_cast<uint64_t>(W) * H > (uint64_t(1) << 28)) throw std::runtime_error("PNG too large");
    if (interlace > 1) throw std::runtime_error("bad PNG interlace method");
    int channels = 0;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: throw std::runtime_error("bad PNG colour type");
    }
    const bool depthOk = (ctype == 0) ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                       : (ctype == 3) ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                      : (depth == 8 || depth == 16);
    if (!depthOk) throw std::runtime_error("PNG bit depth not allowed for this colour type");
    // pass geometry: one pass, or the seven Adam7 passes (x/y origin and spacing)
    static const int ox[7] = {0, 4, 0, 2, 0, 1, 0}, oy[7] = {0, 0, 4, 0, 2, 0, 1};
    static const int sx[7] = {8, 8, 4, 4, 2, 2, 1}, sy[7] = {8, 8, 8, 4, 4, 2, 2};
    const int nPass = interlace ? 7 : 1;
    const size_t bitsPerPixel = static_cast<size_t>(channels) * static_cast<size_t>(depth);
    uint32_t pw[7], ph[7];
    size_t need = 0;
    for (int p = 0; p < nPass; p++) {
        pw[p] = interlace ? (W + sx[p] - 1 - ox[p]) / sx[p] : W;
        ph[p] = interlace ? (H + sy[p] - 1 - oy[p]) / sy[p] : H;
        if (interlace && (W <= static_cast<uint32_t>(ox[p]) || H <= static_cast<uint32_t>(oy[p]))) pw[p] = ph[p] = 0;
        if (pw[p] && ph[p]) need += ((static_cast<size_t>(pw[p]) * bitsPerPixel + 7) / 8 + 1) * ph[p];
    }
    // deflate cannot expand by more than 1032:1, so a header that asks for more than the IDAT bytes can deliver is
    // corrupt: refuse it before allocating for it
    if (need / 1032 > idat.size() + 16) throw std::runtime_error("PNG data too short for the image size it declares");
    std::vector<uint8_t> raw(need);
    uLongf rawLen = static_cast<uLongf>(raw.size());
    int zr = uncompress(raw.data(), &rawLen, idat.data(), static_cast<uLong>(idat.size()));
    if (zr != Z_OK || rawLen != raw.size()) throw std::runtime_error("PNG inflate failed");
    const size_t bps = depth == 16 ? 2 : 1;                 // bytes per sample after unpacking
    const size_t pixBytes = bps * static_cast<size_t>(channels);
    std::vector<uint8_t> img(static_cast<size_t>(W) * H * pixBytes);
    size_t off = 0;
    for (int p = 0; p < nPass; p++) {
        if (!pw[p] || !ph[p]) continue;
        std::vector<uint8_t> part;
        part.reserve(static_cast<size_t>(pw[p]) * ph[p] * pixBytes);
        off += pngUnfilterPass(raw.data() + off, pw[p], ph[p], channels, depth, part);
        if (!interlace) { img.swap(part); break; }
        for (uint32_t y = 0; y < ph[p]; y++)
            for (uint32_t x = 0; x < pw[p]; x++)
                memcpy(&img[((static_cast<size_t>(y) * sy[p] + oy[p]) * W + static_cast<size_t>(x) * sx[p] + ox[p]) * pixBytes],
                       &part[(static_cast<size_t>(y) * pw[p] + x) * pixBytes], pixBytes);
    }
    ImageRGBA8 res;
    res.width = static_cast<int>(W);
    res.height = static_cast<int>(H);
    res.pixels.resize(static_cast<size_t>(W) * H * 4);
    static const int greyScale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    const bool key = (ctype == 0 && trns.size() >= 2) || (ctype == 2 && trns.size() >= 6);   // colour-key transparency
    uint16_t key16[3] = {0, 0, 0};
    uint8_t key8[3] = {0, 0, 0};
    if (key)
        for (int k = 0; k < (ctype == 0 ? 1 : 3); k++) {
            key16[k] = static_cast<uint16_t>((trns[2 * k] << 8) | trns[2 * k + 1]);
            key8[k] = static_cast<uint8_t>((key16[k] & 255) * (ctype == 0 ? greyScale[depth <= 8 ? depth : 8] : 1));
        }
    auto s8 = [&](const uint8_t* px, int k) -> uint8_t { return px[static_cast<size_t>(k) * bps]; };   // 16-bit: the high byte
    auto s16 = [&](const uint8_t* px, int k) -> uint16_t { return static_cast<uint16_t>((px[2 * k] << 8) | px[2 * k + 1]); };
    for (size_t i = 0; i < static_cast<size_t>(W) * H; i++) {
        const uint8_t* px = &img[i * pixBytes];
        uint8_t* o = &res.pixels[i * 4];
        switch (ctype) {
            case 0: {
                const uint8_t g = depth < 8 ? static_cast<uint8_t>(px[0] * greyScale[depth]) : s8(px, 0);
                o[0] = o[1] = o[2] = g;
                o[3] = 255;
                if (key && (depth == 16 ? s16(px, 0) == key16[0] : g == key8[0])) o[3] = 0;
                break;
            }
            case 2:
                o[0] = s8(px, 0); o[1] = s8(px, 1); o[2] = s8(px, 2); o[3] = 255;
                if (key && (depth == 16 ? (s16(px, 0) == key16[0] && s16(px, 1) == key16[1] && s16(px, 2) == key16[2])
                                        : (o[0] == key8[0] && o[1] == key8[1] && o[2] == key8[2]))) o[3] = 0;
                break;
            case 3: {
                const size_t k = px[0];
                if (3 * k + 2 < plte.size()) { o[0] = plte[3 * k]; o[1] = plte[3 * k + 1]; o[2] = plte[3 * k + 2]; }
                else { o[0] = o[1] = o[2] = 0; }
                o[3] = k < trns.size() ? trns[k] : 255;
                break;
            }
            case 4: o[0] = o[1] = o[2] = s8(px, 0); o[3] = s8(px, 1); break;
            default: o[0] = s8(px, 0); o[1] = s8(px, 1); o[2] = s8(px, 2); o[3] = s8(px, 3); break;
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
