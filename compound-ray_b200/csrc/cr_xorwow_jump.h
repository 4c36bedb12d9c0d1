// cr_xorwow_jump.h -- jump-ahead tables for the XORWOW generator, built on the host at start-up.
//
// curand_init(seed, subsequence, offset) (curand_kernel.h:800-825) places a stream 2^67 * subsequence +
// offset steps into the sequence of its seed.  The 160-bit shift-register part v[0..4] of the state is
// advanced by a map that is linear over GF(2) (curand_kernel.h:863-874: xors and shifts only), so a jump
// of n steps is the vector-matrix product v * T^n with T the one-step matrix; the Weyl counter d moves
// by 362437*n mod 2^32 (0 for a whole subsequence).  cuRAND ships T^(2^67 * 4^b) and T^(4^b) as
// precomputed constants and applies each up to three times per base-4 digit; here the matrices are
// DERIVED from the step function (67 squarings) and stored per base-16 digit value, so a jump costs one
// vector-matrix product per non-zero hex digit (6 for a 24-bit stream id instead of ~18).
//
// Table layout (host vector, uploaded as is): [level][digit-1][input bit 0..159][8 words], a row being
// the image of one unit vector (5 words + 3 words of padding so a row is two aligned 16-byte loads);
// levels 0..7 are the subsequence digits (ids < 2^32), levels 8..23 the offset digits (64-bit offsets).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace cr {
namespace xorwow {

constexpr int kBits = 160, kRowWords = 8, kMatrixWords = kBits * kRowWords;
constexpr int kSeqLevels = 8, kOffLevels = 16, kDigits = 15;
constexpr size_t kTableWords = static_cast<size_t>(kSeqLevels + kOffLevels) * kDigits * kMatrixWords;

struct Matrix { uint32_t row[kBits][5]; };

inline void stepV(uint32_t v[5])                      // curand_kernel.h:863-874 without the Weyl counter
{
    const uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}

inline void apply(const Matrix& m, const uint32_t in[5], uint32_t out[5])     // out = in * m
{
    uint32_t r[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++)
        for (uint32_t w = in[i]; w; w &= w - 1) {
            const uint32_t* row = m.row[i * 32 + __builtin_ctz(w)];
            for (int k = 0; k < 5; k++) r[k] ^= row[k];
        }
    memcpy(out, r, sizeof r);
}

inline Matrix multiply(const Matrix& a, const Matrix& b)                      // v*(a*b) == (v*a)*b
{
    Matrix c;
    for (int k = 0; k < kBits; k++) apply(b, a.row[k], c.row[k]);
    return c;
}

inline Matrix oneStep()
{
    Matrix t;
    for (int k = 0; k < kBits; k++) {
        uint32_t e[5] = {0, 0, 0, 0, 0};
        e[k >> 5] = 1u << (k & 31);
        stepV(e);
        memcpy(t.row[k], e, sizeof e);
    }
    return t;
}

inline void store(std::vector<uint32_t>& table, int level, int digit, const Matrix& m)
{
    uint32_t* dst = table.data() + (static_cast<size_t>(level) * kDigits + static_cast<size_t>(digit - 1)) * kMatrixWords;
    for (int k = 0; k < kBits; k++) {
        memcpy(dst + static_cast<size_t>(k) * kRowWords, m.row[k], 5 * sizeof(uint32_t));
        dst[static_cast<size_t>(k) * kRowWords + 5] = dst[static_cast<size_t>(k) * kRowWords + 6] = dst[static_cast<size_t>(k) * kRowWords + 7] = 0u;
    }
}

// base^(d * 16^level) for level < levels, d = 1..15, written at table levels firstLevel + level
inline void fillLevels(std::vector<uint32_t>& table, int firstLevel, int levels, Matrix base)
{
    for (int level = 0; level < levels; level++) {
        Matrix p = base;
        for (int d = 1; d <= kDigits; d++) {
            store(table, firstLevel + level, d, p);
            p = multiply(p, base);                 // after the loop: base^16
        }
        base = p;
    }
}

inline std::vector<uint32_t> buildTable()
{
    std::vector<uint32_t> table(kTableWords);
    const Matrix t = oneStep();
    Matrix seq = t;
    for (int i = 0; i < 67; i++) seq = multiply(seq, seq);     // T^(2^67): one whole subsequence
    fillLevels(table, 0, kSeqLevels, seq);
    fillLevels(table, kSeqLevels, kOffLevels, t);
    return table;
}

// Host restatement of the device routine (k_rngInit) for the CPU-side known-answer test.
inline void jumpHost(const std::vector<uint32_t>& table, uint32_t v[5], int firstLevel, unsigned long long n)
{
    for (int level = firstLevel; n != 0; level++, n >>= 4) {
        const int d = static_cast<int>(n & 15ull);
        if (!d) continue;
        const uint32_t* m = table.data() + (static_cast<size_t>(level) * kDigits + static_cast<size_t>(d - 1)) * kMatrixWords;
        uint32_t r[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 5; i++)
            for (int j = 0; j < 32; j++)
                if ((v[i] >> j) & 1u)
                    for (int k = 0; k < 5; k++) r[k] ^= m[static_cast<size_t>(i * 32 + j) * kRowWords + k];
        memcpy(v, r, sizeof r);
    }
}

// curand_init's seed scrambling (curand_kernel.h:800-812)
inline void seedState(unsigned long long seed, uint32_t& d, uint32_t v[5])
{
    const uint32_t s0 = static_cast<uint32_t>(seed) ^ 0xaad26b49u;
    const uint32_t s1 = static_cast<uint32_t>(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    d = 6615241u + t1 + t0;
    v[0] = 123456789u + t0; v[1] = 362436069u ^ t0; v[2] = 521288629u + t1; v[3] = 88675123u ^ t1; v[4] = 5783321u + t0;
}

}  // namespace xorwow
}  // namespace cr
