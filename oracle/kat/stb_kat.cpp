// oracle/kat/stb_kat.cpp -- TEST INFRASTRUCTURE ONLY.
// Decodes the JPEG fixtures with the reference's OWN decoder -- the vendored stb_image.h where it lies
// under /root/reference/support/tinygltf, exactly as tinygltf calls it (req_comp = 4) -- and writes the
// RGBA bytes to <out>.rgba next to a small JSON index on stdout.  Only buildable where /root/reference exists.
//   stb_kat <out_dir> <file.jpg> ...
#define STB_IMAGE_IMPLEMENTATION
#include <support/tinygltf/stb_image.h>

#include <cstdio>
#include <string>

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    const std::string out = argv[1];
    printf("{\n");
    for (int i = 2; i < argc; i++) {
        int w = 0, h = 0, comp = 0;
        unsigned char* px = stbi_load(argv[i], &w, &h, &comp, 4);
        std::string name = argv[i];
        name = name.substr(name.find_last_of('/') + 1);
        if (!px) { printf(" \"%s\": {\"error\": \"%s\"}%s\n", name.c_str(), stbi_failure_reason(), i + 1 < argc ? "," : ""); continue; }
        const std::string path = out + "/" + name + ".rgba";
        FILE* f = fopen(path.c_str(), "wb");
        fwrite(px, 1, (size_t)w * h * 4, f);
        fclose(f);
        unsigned long long sum = 0;
        for (size_t k = 0; k < (size_t)w * h * 4; k++) sum = sum * 1099511628211ull + px[k];
        printf(" \"%s\": {\"width\": %d, \"height\": %d, \"components\": %d, \"fnv\": %llu}%s\n", name.c_str(), w, h, comp, sum,
               i + 1 < argc ? "," : "");
        stbi_image_free(px);
    }
    printf("}\n");
    return 0;
}
