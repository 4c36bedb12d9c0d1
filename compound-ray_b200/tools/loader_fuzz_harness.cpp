// loader_fuzz_harness.cpp -- host-only driver for tools/loader_fuzz.py: feeds every file named on the command line
// to the glTF loader (*.gltf) or to the image decoders (anything else) and reports how many were accepted.
// Built with g++ -fsanitize=address,undefined against csrc/cr_scene.cpp alone (no CUDA involved): malformed
// input must end in a C++ exception (which the C ABI turns into an error message), never in a memory error.
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>
#include <vector>

#include "cr_jpeg.h"
#include "cr_scene.h"

int main(int argc, char** argv)
{
    int accepted = 0, rejected = 0;
    for (int i = 1; i < argc; i++) {
        const std::string path = argv[i];
        try {
            if (path.size() > 5 && path.compare(path.size() - 5, 5, ".gltf") == 0) {
                const cr::HostScene sc = cr::loadGltfScene(path, false);
                for (const cr::HitboxMesh& hb : sc.hitboxes) (void)cr::pointInsideHitbox(hb, {0.1f, 0.2f, 0.3f});
            } else {
                std::ifstream f(path, std::ios::binary);
                const std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
                (void)cr::decodeImage(bytes.data(), bytes.size());
            }
            accepted++;
        } catch (const std::exception&) {
            rejected++;
        }
    }
    std::cout << "accepted " << accepted << " rejected " << rejected << std::endl;
    return 0;
}
