// oracle/kat/curand_kat.cpp -- TEST INFRASTRUCTURE ONLY.
// Known-answer generator for the XORWOW restatement: runs NVIDIA cuRAND's OWN host
// implementation (CUDA toolkit header curand_kernel.h, the pinned third-party dependency of
// libEyeRenderer3/shaders.cu:51,680-695) and prints JSON.  Built into oracle/_ref/ by the
// Makefile; its output is committed as tests/golden/xorwow_kat.json.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include <curand_kernel.h>

static unsigned f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

int main()
{
    const unsigned long long ids[] = {0ull, 1ull, 2ull, 3ull, 4ull, 31ull, 999ull, 1000ull, 32000ull,
                                      407935ull, 10000000ull, 10239999ull, 4294967295ull, 1ull << 40};
    const unsigned long long offs[] = {0ull, 1ull, 3ull, 4ull, 1000ull, 200000ull, 150000000ull};
    printf("{\n \"init\": [\n");
    bool first = true;
    for (unsigned long long id : ids)
        for (unsigned long long off : offs) {
            curandStateXORWOW_t st;
            curand_init(42ull, id, off, &st);
            printf("%s  {\"seed\": 42, \"subsequence\": %llu, \"offset\": %llu, \"d\": %u, \"v\": [%u, %u, %u, %u, %u], \"draws\": [",
                   first ? "" : ",\n", id, off, st.d, st.v[0], st.v[1], st.v[2], st.v[3], st.v[4]);
            first = false;
            for (int i = 0; i < 8; i++) printf("%s%u", i ? ", " : "", curand(&st));
            printf("]}");
        }
    printf("\n ],\n \"seed_variants\": [\n");
    const unsigned long long seeds[] = {0ull, 1ull, 42ull, 0xdeadbeefcafef00dull};
    first = true;
    for (unsigned long long sd : seeds) {
        curandStateXORWOW_t st;
        curand_init(sd, 5ull, 7ull, &st);
        printf("%s  {\"seed\": %llu, \"subsequence\": 5, \"offset\": 7, \"d\": %u, \"v\": [%u, %u, %u, %u, %u]}",
               first ? "" : ",\n", sd, st.d, st.v[0], st.v[1], st.v[2], st.v[3], st.v[4]);
        first = false;
    }
    // skipahead on a live state
    printf("\n ],\n \"skipahead\": [\n");
    first = true;
    for (unsigned long long n : offs) {
        curandStateXORWOW_t st;
        curand_init(42ull, 77ull, 0ull, &st);
        for (int i = 0; i < 5; i++) curand(&st);
        skipahead(n, &st);
        printf("%s  {\"n\": %llu, \"d\": %u, \"v\": [%u, %u, %u, %u, %u]}", first ? "" : ",\n", n, st.d,
               st.v[0], st.v[1], st.v[2], st.v[3], st.v[4]);
        first = false;
    }
    // the reference's per-frame draw pattern: normal then uniform, repeated (shaders.cu:690-692).
    // Host cuRAND uses libm sinf/cosf/logf, so the FLOAT values are reference-for-tolerance only;
    // the integer state after each frame is exact.
    printf("\n ],\n \"frames\": [\n");
    first = true;
    for (unsigned long long id : {0ull, 7ull, 31999ull}) {
        curandStateXORWOW_t st;
        curand_init(42ull, id, 0ull, &st);
        for (int f = 0; f < 4; f++) {
            float n = curand_normal(&st);
            float u = curand_uniform(&st);
            printf("%s  {\"id\": %llu, \"frame\": %d, \"normal_bits\": %u, \"uniform_bits\": %u, \"normal\": %.9g, \"uniform\": %.9g, \"d\": %u, \"v\": [%u, %u, %u, %u, %u], \"flag\": %d, \"extra_bits\": %u}",
                   first ? "" : ",\n", id, f, f2u(n), f2u(u), n, u, st.d, st.v[0], st.v[1], st.v[2], st.v[3], st.v[4],
                   st.boxmuller_flag, f2u(st.boxmuller_extra));
            first = false;
        }
    }
    printf("\n ]\n}\n");
    return 0;
}
