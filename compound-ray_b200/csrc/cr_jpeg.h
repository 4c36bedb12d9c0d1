// cr_jpeg.h -- baseline and progressive JPEG (SOF0/SOF1/SOF2, 8-bit, Huffman) decoder to RGBA8.
//
// The reference decodes textures through tinygltf -> stb_image with req_comp = 4
// (support/tinygltf/stb_image.h; call site libEyeRenderer3/MulticamScene.cpp:584-600), e.g. the
// ofstad arena's ofstad_patterning.jpg.  To hand the renderer the SAME texels, this decoder follows
// the same published algorithms with the same integer conventions:
//   * IJG "islow" 8x8 inverse DCT (jidctint): 12-bit fixed-point constants, column pass keeps 2 extra
//     bits (+512 >> 10), row pass removes 17 bits with the +128 level shift folded in;
//   * libjpeg "fancy" (triangle) chroma upsampling: (3*near + far + 2) >> 2 in 1-D,
//     (3*t0 + t1 + 8) >> 4 in 2-D, pixel-replication for other ratios;
//   * JFIF YCbCr -> RGB in 20-bit fixed point with 12-bit constants (R = Y + 1.402 Cr, ...), the green
//     chroma-blue product truncated to 16 bits as in the SIMD-compatible formulation.
// Progressive streams (ITU-T T.81 Annex G: spectral selection + successive approximation, DC scans
// interleaved or not, AC scans per component with end-of-band runs) accumulate coefficients over all
// scans and are dequantised + inverse-transformed once at the end, over the blocks that cover the
// component's real extent, as stb_image does.  Arithmetic-coded streams are rejected with an error.
// Pinned by tests/golden/stb_jpeg_kat.json (oracle/kat/stb_kat.cpp compiled against the reference's
// stb_image.h): byte-exact on 4:4:4, 4:2:2, 4:2:0 and greyscale streams, baseline and progressive.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "cr_image.h"

namespace cr {
namespace jpeg {

struct Huffman {
    // canonical code tables: for each code length 1..16, first code value and index of its first symbol
    int mincode[17], maxcode[18], valptr[17];
    uint8_t symbols[256];
    uint8_t lookup[512];       // 9-bit fast path: (length << ... ) stored separately
    uint8_t lookupLen[512];
    bool present = false;
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int blocksW = 0, blocksH = 0;      // in 8x8 blocks, padded to whole MCUs
    int stride = 0, rows = 0;          // decoded plane size in samples
    int w = 0, hgt = 0;                // meaningful samples: ceil(img * h / hmax)
    int dcPred = 0;
    std::vector<uint8_t> plane;
    std::vector<short> coef;           // progressive: blocksW*blocksH*64 coefficients in natural order, before dequantisation
};

class Decoder {
public:
    Decoder(const uint8_t* data, size_t size) : begin_(data), p_(data), end_(data + size) {}

    ImageRGBA8 decode()
    {
        if (end_ - p_ < 2 || p_[0] != 0xFF || p_[1] != 0xD8) fail("not a JPEG stream");
        p_ += 2;
        bool done = false;
        while (!done) {
            const int m = nextMarker();
            switch (m) {
                case 0xC0: case 0xC1: frameHeader(); break;
                case 0xC2: progressive_ = true; frameHeader(); break;
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    fail("unsupported JPEG coding process");
                case 0xC4: huffmanTables(); break;
                case 0xDB: quantTables(); break;
                case 0xDD: { const int len = be16(); if (len != 4) fail("bad DRI"); restartInterval_ = be16(); break; }
                case 0xDA: scan(); break;
                case 0xD9: done = true; break;
                case 0xE0: app0(); break;
                case 0xEE: app14(); break;
                default:
                    if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE || (m >= 0xF0 && m <= 0xFD)) skipSegment();
                    else if (m >= 0xD0 && m <= 0xD7) { /* stray restart marker */ }
                    else fail("unexpected JPEG marker");
            }
            if (p_ >= end_ && !done) { if (scanned_) break; fail("truncated JPEG"); }
        }
        if (!scanned_) fail("JPEG without image data");
        if (progressive_) finishProgressive();
        return assemble();
    }

private:
    [[noreturn]] static void fail(const char* what) { throw std::runtime_error(std::string("JPEG: ") + what); }
    int byte() { if (p_ >= end_) fail("truncated JPEG"); return *p_++; }
    int be16() { const int a = byte(); return (a << 8) | byte(); }
    int nextMarker()
    {
        int c = byte();
        while (c != 0xFF) c = byte();                 // tolerate garbage between segments
        while (c == 0xFF) c = byte();                 // fill bytes
        return c;
    }
    void skipSegment() { const int len = be16(); if (len < 2 || end_ - p_ < len - 2) fail("bad segment length"); p_ += len - 2; }
    void app0()
    {
        const int len = be16();
        if (len < 2 || end_ - p_ < len - 2) fail("bad APP0");
        if (len >= 7 && !memcmp(p_, "JFIF\0", 5)) jfif_ = true;
        p_ += len - 2;
    }
    void app14()
    {
        const int len = be16();
        if (len < 2 || end_ - p_ < len - 2) fail("bad APP14");
        if (len >= 14 && !memcmp(p_, "Adobe\0", 6)) adobeTransform_ = p_[11];
        p_ += len - 2;
    }
    void quantTables()
    {
        int len = be16() - 2;
        while (len > 0) {
            const int q = byte();
            const int prec = q >> 4, t = q & 15;
            if (t > 3 || prec > 1) fail("bad DQT");
            for (int i = 0; i < 64; i++) dequant_[t][kZigzag[i]] = static_cast<uint16_t>(prec ? be16() : byte());
            len -= prec ? 129 : 65;
        }
        if (len != 0) fail("bad DQT length");
    }
    void huffmanTables()
    {
        int len = be16() - 2;
        while (len > 0) {
            const int q = byte();
            const int tc = q >> 4, th = q & 15;
            if (tc > 1 || th > 3) fail("bad DHT");
            int counts[17] = {0}, total = 0;
            for (int i = 1; i <= 16; i++) { counts[i] = byte(); total += counts[i]; }
            if (total > 256) fail("bad DHT");
            Huffman& h = tc ? ac_[th] : dc_[th];
            for (int i = 0; i < total; i++) h.symbols[i] = static_cast<uint8_t>(byte());
            buildHuffman(h, counts);
            len -= 17 + total;
        }
        if (len != 0) fail("bad DHT length");
    }
    static void buildHuffman(Huffman& h, const int* counts)
    {
        int code = 0, k = 0;
        memset(h.lookupLen, 0, sizeof h.lookupLen);
        for (int l = 1; l <= 16; l++) {
            h.valptr[l] = k;
            h.mincode[l] = code;
            if (code + counts[l] > (1 << l)) fail("bad Huffman code lengths");   // before any table write: keeps code < 2^l below
            for (int i = 0; i < counts[l]; i++, k++, code++) {
                if (l <= 9) {
                    const int first = code << (9 - l);
                    for (int j = 0; j < (1 << (9 - l)); j++) { h.lookup[first + j] = h.symbols[k]; h.lookupLen[first + j] = static_cast<uint8_t>(l); }
                }
            }
            h.maxcode[l] = counts[l] ? code - 1 : -1;
            if (code > (1 << l)) fail("bad Huffman code lengths");
            code <<= 1;
        }
        h.maxcode[17] = 0x7fffffff;
        h.present = true;
    }
    void frameHeader()
    {
        const int len = be16();
        if (byte() != 8) fail("only 8-bit JPEG is supported");
        H_ = be16();
        W_ = be16();
        const int n = byte();
        if (H_ <= 0 || W_ <= 0) fail("bad image size");
        // A corrupt header must not turn into a multi-gigabyte allocation: bound the frame, and require at least one
        // byte of stream per 512 pixels (an MCU whose blocks are all "DC unchanged, end of block" still costs 2 bits
        // per block; progressive refinement scans only add to that).
        const size_t pixels = static_cast<size_t>(W_) * static_cast<size_t>(H_);
        if (pixels > (size_t(1) << 28)) fail("image too large");
        if (pixels / 512 > static_cast<size_t>(end_ - begin_)) fail("stream too short for the frame size it declares");
        if (n != 1 && n != 3 && n != 4) fail("bad component count");
        if (len != 8 + 3 * n) fail("bad SOF length");
        comps_.assign(static_cast<size_t>(n), Component());
        hmax_ = vmax_ = 1;
        for (auto& c : comps_) {
            c.id = byte();
            const int q = byte();
            c.h = q >> 4; c.v = q & 15;
            c.tq = byte();
            if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) fail("bad component spec");
            hmax_ = c.h > hmax_ ? c.h : hmax_;
            vmax_ = c.v > vmax_ ? c.v : vmax_;
        }
        for (auto& c : comps_)
            if (hmax_ % c.h || vmax_ % c.v) fail("unsupported sampling factors");
        mcuW_ = 8 * hmax_; mcuH_ = 8 * vmax_;
        mcusX_ = (W_ + mcuW_ - 1) / mcuW_;
        mcusY_ = (H_ + mcuH_ - 1) / mcuH_;
        for (auto& c : comps_) {
            c.w = (W_ * c.h + hmax_ - 1) / hmax_;
            c.hgt = (H_ * c.v + vmax_ - 1) / vmax_;
            c.blocksW = mcusX_ * c.h;
            c.blocksH = mcusY_ * c.v;
            c.stride = c.blocksW * 8;
            c.rows = c.blocksH * 8;
            c.plane.assign(static_cast<size_t>(c.stride) * static_cast<size_t>(c.rows), 0);
            if (progressive_) c.coef.assign(static_cast<size_t>(c.blocksW) * static_cast<size_t>(c.blocksH) * 64, 0);
        }
        haveFrame_ = true;
    }

    // ---- entropy-coded segment -------------------------------------------------------------
    void fillBits()
    {
        while (bitCount_ <= 24) {
            int b = 0;
            if (!hitMarker_ && p_ < end_) {
                b = *p_++;
                if (b == 0xFF) {
                    int c = p_ < end_ ? *p_++ : 0xD9;
                    while (c == 0xFF && p_ < end_) c = *p_++;
                    if (c != 0) { marker_ = c; hitMarker_ = true; b = 0; }
                }
            }
            bitBuf_ |= static_cast<uint32_t>(b) << (24 - bitCount_);
            bitCount_ += 8;
        }
    }
    int decodeSymbol(const Huffman& h)
    {
        if (bitCount_ < 16) fillBits();
        const int peek = static_cast<int>(bitBuf_ >> 23);
        int l = h.lookupLen[peek];
        if (l) { bitBuf_ <<= l; bitCount_ -= l; return h.lookup[peek]; }
        const int code16 = static_cast<int>(bitBuf_ >> 16);
        for (l = 10; l <= 16; l++) {
            const int code = code16 >> (16 - l);
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) {
                bitBuf_ <<= l; bitCount_ -= l;
                return h.symbols[h.valptr[l] + code - h.mincode[l]];
            }
        }
        fail("bad Huffman code");
    }
    int receiveExtend(int n)
    {
        if (bitCount_ < n) fillBits();
        const int v = static_cast<int>(bitBuf_ >> (32 - n));
        bitBuf_ <<= n; bitCount_ -= n;
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
    void resetEntropy()
    {
        bitBuf_ = 0; bitCount_ = 0; hitMarker_ = false; marker_ = 0;
        for (auto& c : comps_) c.dcPred = 0;
    }
    void decodeBlock(Component& c, short* blk)
    {
        memset(blk, 0, 64 * sizeof(short));
        const Huffman& hd = dc_[c.td];
        const Huffman& ha = ac_[c.ta];
        if (!hd.present || !ha.present) fail("missing Huffman table");
        const uint16_t* dq = dequant_[c.tq];
        const int t = decodeSymbol(hd);
        if (t > 16) fail("bad DC code");
        const int diff = t ? receiveExtend(t) : 0;
        c.dcPred += diff;
        blk[0] = static_cast<short>(c.dcPred * dq[0]);
        for (int k = 1; k < 64;) {
            const int rs = decodeSymbol(ha);
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) {
                if (rs != 0xF0) break;      // end of block
                k += 16;
            } else {
                k += r;
                if (k > 63) fail("bad AC run");
                const int z = kZigzag[k++];
                blk[z] = static_cast<short>(receiveExtend(s) * dq[z]);
            }
        }
    }
    // ---- progressive scans (T.81 Annex G) -----------------------------------------------------------
    int getBits(int n)
    {
        if (bitCount_ < n) fillBits();
        const int v = static_cast<int>(bitBuf_ >> (32 - n));
        bitBuf_ <<= n; bitCount_ -= n;
        return v;
    }
    void progressiveDC(Component& c, short* d, int Ah, int Al)
    {
        if (Ah == 0) {                                  // first pass: the DC difference, scaled by the point transform
            const Huffman& hd = dc_[c.td];
            if (!hd.present) fail("missing Huffman table");
            const int t = decodeSymbol(hd);
            if (t > 16) fail("bad DC code");
            c.dcPred += t ? receiveExtend(t) : 0;
            d[0] = static_cast<short>(c.dcPred * (1 << Al));
        } else if (getBits(1)) {                        // refinement: one more bit of precision
            d[0] = static_cast<short>(d[0] + (1 << Al));
        }
    }
    void refineNonZero(short* q, int bit)               // correction bit for a coefficient with history
    {
        if (getBits(1) && (*q & bit) == 0) *q = static_cast<short>(*q > 0 ? *q + bit : *q - bit);
    }
    void progressiveAC(Component& c, short* d, int Ss, int Se, int Ah, int Al)
    {
        const Huffman& ha = ac_[c.ta];
        if (!ha.present) fail("missing Huffman table");
        if (Ah == 0) {                                  // first pass over the band Ss..Se
            if (eobRun_) { --eobRun_; return; }
            for (int k = Ss; k <= Se;) {
                const int rs = decodeSymbol(ha);
                const int r = rs >> 4, sz = rs & 15;
                if (sz == 0) {
                    if (r < 15) {                       // EOBn: this and the next 2^r + extra - 1 blocks are finished
                        eobRun_ = 1 << r;
                        if (r) eobRun_ += getBits(r);
                        --eobRun_;
                        break;
                    }
                    k += 16;                            // ZRL
                } else {
                    k += r;
                    if (k > 63) fail("bad AC run");
                    d[kZigzag[k++]] = static_cast<short>(receiveExtend(sz) * (1 << Al));
                }
            }
            return;
        }
        const int bit = 1 << Al;                        // refinement pass
        if (eobRun_) {
            --eobRun_;
            for (int k = Ss; k <= Se; k++) { short* q = &d[kZigzag[k]]; if (*q != 0) refineNonZero(q, bit); }
            return;
        }
        for (int k = Ss; k <= Se;) {
            const int rs = decodeSymbol(ha);
            int r = rs >> 4, sz = rs & 15, value = 0;
            if (sz == 0) {
                if (r < 15) {                           // EOBn: the rest of this band only receives correction bits
                    eobRun_ = (1 << r) - 1;
                    if (r) eobRun_ += getBits(r);
                    r = 64;
                }                                       // else ZRL: skip 16 zero-history coefficients
            } else {
                if (sz != 1) fail("bad refinement code");
                value = getBits(1) ? bit : -bit;        // a new coefficient of magnitude 1 << Al
            }
            while (k <= Se) {                           // r counts zero-history coefficients only
                short* q = &d[kZigzag[k++]];
                if (*q != 0) refineNonZero(q, bit);
                else if (r == 0) { *q = static_cast<short>(value); break; }
                else --r;
            }
        }
    }
    // true: keep going; false: a marker other than RSTn ended the scan
    bool restartBoundary(int& todo)
    {
        if (--todo > 0) return true;
        if (!hitMarker_) {
            bitCount_ = 0; bitBuf_ = 0;
            if (p_ + 1 < end_ && p_[0] == 0xFF && p_[1] >= 0xD0 && p_[1] <= 0xD7) p_ += 2;
        } else if (!(marker_ >= 0xD0 && marker_ <= 0xD7)) {
            return false;
        }
        resetEntropy();
        eobRun_ = 0;
        todo = restartInterval_;
        return true;
    }
    void scanProgressive()
    {
        const int len = be16();
        const int n = byte();
        if (n < 1 || n > static_cast<int>(comps_.size()) || len != 6 + 2 * n) fail("bad SOS");
        int order[4];
        for (int i = 0; i < n; i++) {
            const int id = byte(), q = byte();
            int which = -1;
            for (size_t j = 0; j < comps_.size(); j++) if (comps_[j].id == id) which = static_cast<int>(j);
            if (which < 0) fail("scan names an unknown component");
            order[i] = which;
            comps_[static_cast<size_t>(which)].td = q >> 4;
            comps_[static_cast<size_t>(which)].ta = q & 15;
            if ((q >> 4) > 3 || (q & 15) > 3) fail("bad table selector");
        }
        const int Ss = byte(), Se = byte(), a = byte();
        const int Ah = a >> 4, Al = a & 15;
        if (Ss > 63 || Se > 63 || Ss > Se || Ah > 13 || Al > 13) fail("bad progressive scan parameters");
        if (Ss == 0 && Se != 0) fail("a progressive scan cannot mix DC and AC coefficients");
        if (Ss != 0 && n != 1) fail("progressive AC scans carry one component");
        resetEntropy();
        eobRun_ = 0;
        int todo = restartInterval_ ? restartInterval_ : 0x7fffffff;
        bool go = true;
        if (n == 1) {                                   // non-interleaved: the component's own block raster, real extent only
            Component& c = comps_[static_cast<size_t>(order[0])];
            const int bw = (c.w + 7) >> 3, bh = (c.hgt + 7) >> 3;
            for (int j = 0; j < bh && go; j++)
                for (int i = 0; i < bw && go; i++) {
                    short* d = &c.coef[64 * (static_cast<size_t>(i) + static_cast<size_t>(j) * static_cast<size_t>(c.blocksW))];
                    if (Ss == 0) progressiveDC(c, d, Ah, Al); else progressiveAC(c, d, Ss, Se, Ah, Al);
                    go = restartBoundary(todo);
                }
        } else {                                        // interleaved DC scan: MCU order
            for (int my = 0; my < mcusY_ && go; my++)
                for (int mx = 0; mx < mcusX_ && go; mx++) {
                    for (int k = 0; k < n; k++) {
                        Component& c = comps_[static_cast<size_t>(order[k])];
                        for (int by = 0; by < c.v; by++)
                            for (int bx = 0; bx < c.h; bx++) {
                                const size_t x2 = static_cast<size_t>(mx * c.h + bx), y2 = static_cast<size_t>(my * c.v + by);
                                progressiveDC(c, &c.coef[64 * (x2 + y2 * static_cast<size_t>(c.blocksW))], Ah, Al);
                            }
                    }
                    go = restartBoundary(todo);
                }
        }
        scanned_ = true;
        if (hitMarker_ && marker_ == 0xD9) p_ = end_;
        else if (hitMarker_) p_ -= 2;
    }
    void finishProgressive()
    {
        short blk[64];
        for (auto& c : comps_) {
            const uint16_t* dq = dequant_[c.tq];
            const int bw = (c.w + 7) >> 3, bh = (c.hgt + 7) >> 3;
            for (int j = 0; j < bh; j++)
                for (int i = 0; i < bw; i++) {
                    const short* d = &c.coef[64 * (static_cast<size_t>(i) + static_cast<size_t>(j) * static_cast<size_t>(c.blocksW))];
                    for (int k = 0; k < 64; k++) blk[k] = static_cast<short>(d[k] * dq[k]);
                    idct(blk, &c.plane[static_cast<size_t>(j) * 8 * static_cast<size_t>(c.stride) + static_cast<size_t>(i) * 8], c.stride);
                }
        }
    }

    void scan()
    {
        if (!haveFrame_) fail("SOS before SOF");
        if (progressive_) { scanProgressive(); return; }
        const int len = be16();
        const int n = byte();
        if (n != static_cast<int>(comps_.size()) || len != 6 + 2 * n) fail("only single, fully interleaved scans are supported");
        for (int i = 0; i < n; i++) {
            const int id = byte(), q = byte();
            if (comps_[static_cast<size_t>(i)].id != id) fail("scan component order mismatch");
            comps_[static_cast<size_t>(i)].td = q >> 4;
            comps_[static_cast<size_t>(i)].ta = q & 15;
            if (comps_[static_cast<size_t>(i)].td > 3 || comps_[static_cast<size_t>(i)].ta > 3) fail("bad table selector");
        }
        byte(); byte(); byte();            // Ss, Se, Ah/Al: fixed for baseline
        resetEntropy();
        short blk[64];
        int todo = restartInterval_ ? restartInterval_ : 0x7fffffff;
        for (int my = 0; my < mcusY_; my++)
            for (int mx = 0; mx < mcusX_; mx++) {
                for (auto& c : comps_)
                    for (int by = 0; by < c.v; by++)
                        for (int bx = 0; bx < c.h; bx++) {
                            decodeBlock(c, blk);
                            const int x0 = (mx * c.h + bx) * 8, y0 = (my * c.v + by) * 8;
                            idct(blk, &c.plane[static_cast<size_t>(y0) * static_cast<size_t>(c.stride) + static_cast<size_t>(x0)], c.stride);
                        }
                if (--todo <= 0) {
                    // expect RSTn: discard buffered bits, consume the marker
                    if (!hitMarker_) {
                        bitCount_ = 0; bitBuf_ = 0;
                        if (p_ + 1 < end_ && p_[0] == 0xFF && p_[1] >= 0xD0 && p_[1] <= 0xD7) p_ += 2;
                    } else if (!(marker_ >= 0xD0 && marker_ <= 0xD7)) {
                        my = mcusY_; break;     // some other marker: stop decoding
                    }
                    resetEntropy();
                    todo = restartInterval_;
                }
            }
        scanned_ = true;
        if (hitMarker_ && marker_ == 0xD9) p_ = end_;      // EOI already consumed by the bit reader
        else if (hitMarker_) p_ -= 2;                      // hand the marker back to the segment parser
    }

    // ---- IJG islow inverse DCT ------------------------------------------------------------------
    static inline int fix(double x) { return static_cast<int>(x * 4096 + 0.5); }
    static inline uint8_t clamp255(int x) { return static_cast<uint8_t>(x < 0 ? 0 : (x > 255 ? 255 : x)); }
    // Integer arithmetic of the inverse DCT wraps modulo 2^32: coefficients of a valid stream stay far below the
    // range of int, and a corrupt stream (whose pixels are garbage anyway) must not reach undefined behaviour.
    static inline int wadd(int a, int b) { return static_cast<int>(static_cast<uint32_t>(a) + static_cast<uint32_t>(b)); }
    static inline int wsub(int a, int b) { return static_cast<int>(static_cast<uint32_t>(a) - static_cast<uint32_t>(b)); }
    static inline int wmul(int a, int b) { return static_cast<int>(static_cast<uint32_t>(a) * static_cast<uint32_t>(b)); }
    struct Idct1D { int x0, x1, x2, x3, t0, t1, t2, t3; };
    static inline Idct1D idct1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7)
    {
        Idct1D r;
        int p2 = s2, p3 = s6;
        int p1 = wmul(wadd(p2, p3), fix(0.5411961f));
        int t2 = wadd(p1, wmul(p3, fix(-1.847759065f)));
        int t3 = wadd(p1, wmul(p2, fix(0.765366865f)));
        p2 = s0; p3 = s4;
        int t0 = wmul(wadd(p2, p3), 4096);
        int t1 = wmul(wsub(p2, p3), 4096);
        r.x0 = wadd(t0, t3); r.x3 = wsub(t0, t3); r.x1 = wadd(t1, t2); r.x2 = wsub(t1, t2);
        t0 = s7; t1 = s5; t2 = s3; t3 = s1;
        p3 = wadd(t0, t2);
        int p4 = wadd(t1, t3);
        p1 = wadd(t0, t3);
        p2 = wadd(t1, t2);
        const int p5 = wmul(wadd(p3, p4), fix(1.175875602f));
        t0 = wmul(t0, fix(0.298631336f));
        t1 = wmul(t1, fix(2.053119869f));
        t2 = wmul(t2, fix(3.072711026f));
        t3 = wmul(t3, fix(1.501321110f));
        p1 = wadd(p5, wmul(p1, fix(-0.899976223f)));
        p2 = wadd(p5, wmul(p2, fix(-2.562915447f)));
        p3 = wmul(p3, fix(-1.961570560f));
        p4 = wmul(p4, fix(-0.390180644f));
        r.t3 = wadd(wadd(t3, p1), p4); r.t2 = wadd(wadd(t2, p2), p3); r.t1 = wadd(wadd(t1, p2), p4); r.t0 = wadd(wadd(t0, p1), p3);
        return r;
    }
    static void idct(const short* d, uint8_t* out, int stride)
    {
        int val[64];
        for (int i = 0; i < 8; i++) {
            const short* c = d + i;
            int* v = val + i;
            if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
                const int dc = c[0] * 4;
                v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
            } else {
                Idct1D r = idct1d(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
                r.x0 = wadd(r.x0, 512); r.x1 = wadd(r.x1, 512); r.x2 = wadd(r.x2, 512); r.x3 = wadd(r.x3, 512);
                v[0] = wadd(r.x0, r.t3) >> 10; v[56] = wsub(r.x0, r.t3) >> 10;
                v[8] = wadd(r.x1, r.t2) >> 10; v[48] = wsub(r.x1, r.t2) >> 10;
                v[16] = wadd(r.x2, r.t1) >> 10; v[40] = wsub(r.x2, r.t1) >> 10;
                v[24] = wadd(r.x3, r.t0) >> 10; v[32] = wsub(r.x3, r.t0) >> 10;
            }
        }
        for (int i = 0; i < 8; i++) {
            const int* v = val + 8 * i;
            uint8_t* o = out + static_cast<size_t>(i) * static_cast<size_t>(stride);
            Idct1D r = idct1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            const int bias = 65536 + (128 << 17);
            r.x0 = wadd(r.x0, bias); r.x1 = wadd(r.x1, bias); r.x2 = wadd(r.x2, bias); r.x3 = wadd(r.x3, bias);
            o[0] = clamp255(wadd(r.x0, r.t3) >> 17); o[7] = clamp255(wsub(r.x0, r.t3) >> 17);
            o[1] = clamp255(wadd(r.x1, r.t2) >> 17); o[6] = clamp255(wsub(r.x1, r.t2) >> 17);
            o[2] = clamp255(wadd(r.x2, r.t1) >> 17); o[5] = clamp255(wsub(r.x2, r.t1) >> 17);
            o[3] = clamp255(wadd(r.x3, r.t0) >> 17); o[4] = clamp255(wsub(r.x3, r.t0) >> 17);
        }
    }

    // ---- upsampling + colour conversion ---------------------------------------------------------
    static void upsampleRow(uint8_t* out, const uint8_t* nearRow, const uint8_t* farRow, int w, int hs, int vs)
    {
        if (hs == 1 && vs == 1) { memcpy(out, nearRow, static_cast<size_t>(w)); return; }
        if (hs == 1 && vs == 2) { for (int i = 0; i < w; i++) out[i] = static_cast<uint8_t>((3 * nearRow[i] + farRow[i] + 2) >> 2); return; }
        if (hs == 2 && vs == 1) {
            const uint8_t* in = nearRow;
            if (w == 1) { out[0] = out[1] = in[0]; return; }
            out[0] = in[0];
            out[1] = static_cast<uint8_t>((in[0] * 3 + in[1] + 2) >> 2);
            int i;
            for (i = 1; i < w - 1; i++) {
                const int n = 3 * in[i] + 2;
                out[i * 2] = static_cast<uint8_t>((n + in[i - 1]) >> 2);
                out[i * 2 + 1] = static_cast<uint8_t>((n + in[i + 1]) >> 2);
            }
            out[i * 2] = static_cast<uint8_t>((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
            out[i * 2 + 1] = in[w - 1];
            return;
        }
        if (hs == 2 && vs == 2) {
            if (w == 1) { out[0] = out[1] = static_cast<uint8_t>((3 * nearRow[0] + farRow[0] + 2) >> 2); return; }
            int t1 = 3 * nearRow[0] + farRow[0];
            out[0] = static_cast<uint8_t>((t1 + 2) >> 2);
            for (int i = 1; i < w; i++) {
                const int t0 = t1;
                t1 = 3 * nearRow[i] + farRow[i];
                out[i * 2 - 1] = static_cast<uint8_t>((3 * t0 + t1 + 8) >> 4);
                out[i * 2] = static_cast<uint8_t>((3 * t1 + t0 + 8) >> 4);
            }
            out[w * 2 - 1] = static_cast<uint8_t>((t1 + 2) >> 2);
            return;
        }
        for (int i = 0; i < w; i++)                   // other ratios: replicate horizontally, nearest row
            for (int j = 0; j < hs; j++) out[i * hs + j] = nearRow[i];
    }
    static inline int f2f(float x) { return static_cast<int>(x * 4096.0f + 0.5f) << 8; }

    ImageRGBA8 assemble()
    {
        ImageRGBA8 img;
        img.width = W_; img.height = H_;
        img.pixels.resize(static_cast<size_t>(W_) * static_cast<size_t>(H_) * 4);
        const size_t nc = comps_.size();
        bool isRgb = false;
        if (nc == 3) {
            const bool idsRgb = comps_[0].id == 'R' && comps_[1].id == 'G' && comps_[2].id == 'B';
            isRgb = idsRgb || (adobeTransform_ == 0 && !jfif_);
        }
        struct Up { int hs, vs, ystep, ypos, wLores; const uint8_t* line0; const uint8_t* line1; std::vector<uint8_t> buf; };
        std::vector<Up> up(nc);
        for (size_t k = 0; k < nc; k++) {
            Up& u = up[k];
            u.hs = hmax_ / comps_[k].h; u.vs = vmax_ / comps_[k].v;
            u.ystep = u.vs >> 1; u.ypos = 0;
            u.wLores = (W_ + u.hs - 1) / u.hs;
            u.line0 = u.line1 = comps_[k].plane.data();
            u.buf.assign(static_cast<size_t>(W_) + 8, 0);
        }
        std::vector<const uint8_t*> row(nc);
        for (int j = 0; j < H_; j++) {
            for (size_t k = 0; k < nc; k++) {
                Up& u = up[k];
                const bool bot = u.ystep >= (u.vs >> 1);
                upsampleRow(u.buf.data(), bot ? u.line1 : u.line0, bot ? u.line0 : u.line1, u.wLores, u.hs, u.vs);
                row[k] = u.buf.data();
                if (++u.ystep >= u.vs) {
                    u.ystep = 0;
                    u.line0 = u.line1;
                    if (++u.ypos < comps_[k].hgt) u.line1 += comps_[k].stride;
                }
            }
            uint8_t* out = &img.pixels[static_cast<size_t>(j) * static_cast<size_t>(W_) * 4];
            if (nc == 1) {
                for (int i = 0; i < W_; i++, out += 4) { out[0] = out[1] = out[2] = row[0][i]; out[3] = 255; }
            } else if (nc == 3 && isRgb) {
                for (int i = 0; i < W_; i++, out += 4) { out[0] = row[0][i]; out[1] = row[1][i]; out[2] = row[2][i]; out[3] = 255; }
            } else if (nc == 3 || (nc == 4 && adobeTransform_ != 0 && adobeTransform_ != 2)) {
                for (int i = 0; i < W_; i++, out += 4) {
                    const int yFixed = (row[0][i] << 20) + (1 << 19);
                    const int cr = row[2][i] - 128, cb = row[1][i] - 128;
                    int r = yFixed + cr * f2f(1.40200f);
                    int g = yFixed + (cr * -f2f(0.71414f)) + ((cb * -f2f(0.34414f)) & static_cast<int>(0xffff0000u));
                    int b = yFixed + cb * f2f(1.77200f);
                    r >>= 20; g >>= 20; b >>= 20;
                    out[0] = clamp255(r); out[1] = clamp255(g); out[2] = clamp255(b); out[3] = 255;
                }
            } else {
                fail("CMYK / YCCK JPEG is not supported");
            }
        }
        return img;
    }

    static constexpr uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                                            41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                                            15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

    const uint8_t* begin_;
    const uint8_t* p_;
    const uint8_t* end_;
    int W_ = 0, H_ = 0, hmax_ = 1, vmax_ = 1, mcuW_ = 8, mcuH_ = 8, mcusX_ = 0, mcusY_ = 0;
    int restartInterval_ = 0;
    bool haveFrame_ = false, scanned_ = false, jfif_ = false, progressive_ = false;
    int eobRun_ = 0;                   // progressive: blocks still covered by the last end-of-band run
    int adobeTransform_ = -1;
    std::vector<Component> comps_;
    uint16_t dequant_[4][64] = {};
    Huffman dc_[4], ac_[4];
    uint32_t bitBuf_ = 0;
    int bitCount_ = 0;
    bool hitMarker_ = false;
    int marker_ = 0;
};

}  // namespace jpeg

inline ImageRGBA8 decodeJPEG(const uint8_t* data, size_t size) { return jpeg::Decoder(data, size).decode(); }

}  // namespace cr
