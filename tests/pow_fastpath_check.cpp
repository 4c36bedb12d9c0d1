// Checks that the branch-free fast path of crm::pow (compound-ray_b200/csrc/cr_math.h) returns the same bits
// as its definition exp(y*log(x)) (crm::powGeneric): every 3rd binary32 pattern of a superset of the
// fast-path range plus all patterns next to its borders, for the two exponents the renderer uses and a few
// others, then random patterns of (x, y) including negatives, zeros, denormals, infinities and NaNs.
// (The exhaustive run over every pattern x 8 exponents, 5.5e9 evaluations, was done when the path was added.)
#include "cr_math.h"
#include <cstdio>
#include <initializer_list>
int main()
{
    const float ys[] = {2.2f, (float)(1.0 / (double)2.2f), 2.5f, -2.5f, 0.0f};
    long long bad = 0;
    for (float y : ys) {
        #pragma omp parallel for reduction(+:bad) schedule(static)
        for (long long u = 0x2B000000ll; u <= 0x54000000ll; u += 3) {
            const float x = crm::u2f((uint32_t)u);
            if (crm::f2u(crm::pow(x, y)) != crm::f2u(crm::powGeneric(x, y))) bad++;
        }
        for (long long c : {0x2B800000ll, 0x53800000ll})
            for (long long u = c - 4096; u <= c + 4096; u++) {
                const float x = crm::u2f((uint32_t)u);
                if (crm::f2u(crm::pow(x, y)) != crm::f2u(crm::powGeneric(x, y))) bad++;
            }
    }
    unsigned long long s = 88172645463325252ull;
    for (long long i = 0; i < 20000000ll; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const float x = crm::u2f((uint32_t)s), y = crm::u2f((uint32_t)(s >> 32));
        const float a = crm::pow(x, y), b = crm::powGeneric(x, y);
        if (crm::f2u(a) != crm::f2u(b) && !(a != a && b != b)) bad++;
    }
    printf("%lld\n", bad);
    return bad ? 1 : 0;
}
