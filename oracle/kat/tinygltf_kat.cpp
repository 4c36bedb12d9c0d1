// oracle/kat/tinygltf_kat.cpp -- TEST INFRASTRUCTURE ONLY.
// Known-answer generator for the scene loader: parses glTF files with the REFERENCE's own vendored tinygltf
// (support/tinygltf/tiny_gltf.h + json.hpp + stb_image.h, compiled where they lie) exactly as
// MulticamScene.cpp:531-545 does, walks the node forest in the reference's order (root nodes by index,
// :722-735; children depth-first, :519-525), builds node transforms with the reference's sutil classes (:173-205),
// reads accessors the way bufferViewFromGLTF does (:85-109: buffer + bufferView.byteOffset + accessor.byteOffset,
// bufferView.byteStride or the element size) and prints, per scene: the cameras in insertion order with pose bits,
// and per mesh primitive the triangle/vertex counts, material facts and FNV-1a hashes of the world-space
// triangles (v0, v1-v0, v2-v0), corner UVs and corner colours; per texture the size and a hash of the RGBA8 pixels
// tinygltf hands to the renderer.  Output committed as tests/golden/tinygltf_kat.json; tests/test_host.py computes
// the same digests from the product's loader.  Only buildable where /root/reference exists.
#define TINYGLTF_IMPLEMENTATION
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#include <support/tinygltf/tiny_gltf.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include <sutil/Matrix.h>
#include <sutil/Quaternion.h>
#include <sutil/vec_math.h>
using namespace sutil;

struct Fnv {
    unsigned long long h = 1469598103934665603ull;
    void bytes(const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } }
    void f(float v) { bytes(&v, 4); }
};
static unsigned f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

struct Out {
    std::string cameras, meshes;
    int nCams = 0, nMeshes = 0;
};

static bool extraTrue(const tinygltf::Value& extras, const char* key)   // MulticamScene.cpp:131-146
{
    if (!extras.IsObject() || !extras.Has(key)) return false;
    const tinygltf::Value& v = extras.Get(key);
    if (v.IsBool()) return v.Get<bool>();
    if (v.IsString()) { std::string s = v.Get<std::string>(); for (auto& c : s) c = (char)tolower(c); return s == "true"; }
    return false;
}

static const unsigned char* accessorBase(const tinygltf::Model& m, int acc, size_t& stride, size_t elem, size_t& count)
{
    const auto& a = m.accessors[acc];
    const auto& bv = m.bufferViews[a.bufferView];
    stride = bv.byteStride ? bv.byteStride : elem;
    count = a.count;
    return m.buffers[bv.buffer].data.data() + bv.byteOffset + a.byteOffset;
}

static void walk(const tinygltf::Model& model, const tinygltf::Node& node, const Matrix4x4& parent, const std::string& dir, Out& out)
{
    // MulticamScene.cpp:173-205
    const Matrix4x4 translation = node.translation.empty() ? Matrix4x4::identity()
        : Matrix4x4::translate(make_float3((float)node.translation[0], (float)node.translation[1], (float)node.translation[2]));
    const Matrix4x4 rotation = node.rotation.empty() ? Matrix4x4::identity()
        : Quaternion((float)node.rotation[3], (float)node.rotation[0], (float)node.rotation[1], (float)node.rotation[2]).rotationMatrix();
    const Matrix4x4 scale = node.scale.empty() ? Matrix4x4::identity()
        : Matrix4x4::scale(make_float3((float)node.scale[0], (float)node.scale[1], (float)node.scale[2]));
    std::vector<float> gm;
    for (double x : node.matrix) gm.push_back((float)x);
    const Matrix4x4 matrix = node.matrix.empty() ? Matrix4x4::identity() : Matrix4x4(reinterpret_cast<float*>(gm.data())).transpose();
    const Matrix4x4 xf = parent * matrix * translation * rotation * scale;
    char buf[1024];

    if (node.camera != -1) {                                             // :207-328
        const auto& cam = model.cameras[node.camera];
        const float3 eye = make_float3(xf * make_float4(0, 0, 0, 1));
        const float3 up = make_float3(xf * make_float4(0, 1, 0, 0));
        const float3 fwd = make_float3(xf * make_float4(0, 0, -1, 0));
        const float3 right = make_float3(xf * make_float4(1, 0, 0, 0));
        const char* kind = "perspective";
        if (cam.type == "orthographic") kind = "orthographic";
        else if (extraTrue(cam.extras, "panoramic")) kind = "panoramic";
        else if (extraTrue(cam.extras, "compound-eye")) {
            kind = "compound";
            const std::string eyePath = cam.extras.Get("compound-structure").Get<std::string>();
            std::ifstream a(eyePath), b(dir + eyePath);
            if (!a.is_open() && !b.is_open()) return;                     // "read cancelled": the camera is not added (:266-279)
        }
        snprintf(buf, sizeof buf, "%s   {\"name\": \"%s\", \"kind\": \"%s\", \"pose\": [%u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u, %u]}",
                 out.nCams ? ",\n" : "", cam.name.c_str(), kind, f2u(eye.x), f2u(eye.y), f2u(eye.z), f2u(right.x), f2u(right.y), f2u(right.z),
                 f2u(up.x), f2u(up.y), f2u(up.z), f2u(fwd.x), f2u(fwd.y), f2u(fwd.z));
        out.cameras += buf;
        out.nCams++;
        return;
    }
    if (node.mesh != -1) {
        const auto& mesh = model.meshes[node.mesh];
        if (extraTrue(mesh.extras, "hitbox")) return;                    // hitbox meshes are not rendered (:329-345)
        for (const auto& prim : mesh.primitives) {
            if (prim.mode != TINYGLTF_MODE_TRIANGLES) continue;
            size_t ps, pc;
            const unsigned char* pb = accessorBase(model, prim.attributes.at("POSITION"), ps, 12, pc);
            std::vector<unsigned> idx;
            if (prim.indices >= 0) {
                const int ct = model.accessors[prim.indices].componentType;
                size_t is, ic;
                const unsigned char* ib = accessorBase(model, prim.indices, is, ct == TINYGLTF_COMPONENT_TYPE_UNSIGNED_SHORT ? 2 : 4, ic);
                for (size_t i = 0; i < ic; i++) {
                    if (ct == TINYGLTF_COMPONENT_TYPE_UNSIGNED_SHORT) { unsigned short v; memcpy(&v, ib + i * is, 2); idx.push_back(v); }
                    else { unsigned v; memcpy(&v, ib + i * is, 4); idx.push_back(v); }
                }
            } else for (size_t i = 0; i < pc; i++) idx.push_back((unsigned)i);
            const size_t ntri = idx.size() / 3;
            Fnv ht, hu, hc;
            std::vector<float3> world(pc);
            for (size_t v = 0; v < pc; v++) { float p[3]; memcpy(p, pb + v * ps, 12); world[v] = make_float3(xf * make_float4(p[0], p[1], p[2], 1.0f)); }
            for (size_t t = 0; t < ntri; t++) {
                const float3 p0 = world[idx[3 * t]], p1 = world[idx[3 * t + 1]], p2 = world[idx[3 * t + 2]];
                const float3 e1 = p1 - p0, e2 = p2 - p0;
                ht.f(p0.x); ht.f(p0.y); ht.f(p0.z); ht.f(e1.x); ht.f(e1.y); ht.f(e1.z); ht.f(e2.x); ht.f(e2.y); ht.f(e2.z);
            }
            int hasUV = 0, colorType = -1;
            auto uvIt = prim.attributes.find("TEXCOORD_0");
            if (uvIt != prim.attributes.end() && model.accessors[uvIt->second].componentType == TINYGLTF_COMPONENT_TYPE_FLOAT) {
                hasUV = 1;
                size_t us, uc;
                const unsigned char* ub = accessorBase(model, uvIt->second, us, 8, uc);
                for (size_t t = 0; t < ntri * 3; t++) { float uv[2]; memcpy(uv, ub + idx[t] * us, 8); hu.f(uv[0]); hu.f(uv[1]); }
            }
            auto cIt = prim.attributes.find("COLOR_0");
            if (cIt != prim.attributes.end() && model.accessors[cIt->second].type == TINYGLTF_TYPE_VEC4) {   // :435-518
                colorType = model.accessors[cIt->second].componentType;
                const size_t cs = colorType == TINYGLTF_COMPONENT_TYPE_FLOAT ? 4 : (colorType == TINYGLTF_COMPONENT_TYPE_UNSIGNED_SHORT ? 2 : 1);
                size_t st, cc;
                const unsigned char* cb = accessorBase(model, cIt->second, st, 4 * cs, cc);
                for (size_t t = 0; t < ntri * 3; t++)
                    for (int k = 0; k < 4; k++) {
                        const unsigned char* p = cb + idx[t] * st + k * cs;
                        float f;
                        if (cs == 4) memcpy(&f, p, 4);
                        else if (cs == 2) { unsigned short v; memcpy(&v, p, 2); float4 c = make_float4((float)v); c /= 65535.0f; f = c.x; }   // LocalGeometry.h:131-133
                        else { float4 c = make_float4((float)*p); c /= 255.0f; f = c.x; }                                                    // :119-121
                        hc.f(f);
                    }
            }
            float base[4] = {1, 1, 1, 1};
            int tex = -1;
            if (prim.material >= 0) {                                       // :626-650
                const auto& mat = model.materials[prim.material];
                const auto it = mat.values.find("baseColorFactor");
                if (it != mat.values.end()) { const tinygltf::ColorValue c = it->second.ColorFactor(); for (int k = 0; k < 4; k++) base[k] = (float)c[k]; }
                const auto tt = mat.values.find("baseColorTexture");
                if (tt != mat.values.end()) tex = tt->second.TextureIndex();
            }
            snprintf(buf, sizeof buf, "%s   {\"name\": \"%s\", \"tris\": %zu, \"verts\": %zu, \"color_type\": %d, \"has_uv\": %d, \"tex\": %d, "
                     "\"base_color\": [%u, %u, %u, %u], \"tri_hash\": \"%016llx\", \"uv_hash\": \"%016llx\", \"col_hash\": \"%016llx\"}",
                     out.nMeshes ? ",\n" : "", mesh.name.c_str(), ntri, pc, colorType, hasUV, tex, f2u(base[0]), f2u(base[1]), f2u(base[2]), f2u(base[3]),
                     ht.h, hu.h, hc.h);
            out.meshes += buf;
            out.nMeshes++;
        }
        return;
    }
    for (int child : node.children) walk(model, model.nodes[child], xf, dir, out);
}

int main(int argc, char** argv)
{
    printf("{\n");
    for (int a = 1; a + 1 < argc; a += 2) {
        const std::string key = argv[a], path = argv[a + 1];
        tinygltf::Model model;
        tinygltf::TinyGLTF loader;
        std::string err, warn;
        if (!loader.LoadASCIIFromFile(&model, &err, &warn, path)) { fprintf(stderr, "failed to load %s: %s\n", path.c_str(), err.c_str()); return 1; }
        const size_t slash = path.find_last_of('/');
        const std::string dir = slash == std::string::npos ? "" : path.substr(0, slash + 1);
        Out out;
        std::vector<int> root(model.nodes.size(), 1);
        for (auto& n : model.nodes) for (int c : n.children) root[c] = 0;
        for (size_t i = 0; i < root.size(); i++) if (root[i]) walk(model, model.nodes[i], Matrix4x4::identity(), dir, out);
        printf(" \"%s\": {\n  \"cameras\": [\n%s\n  ],\n  \"meshes\": [\n%s\n  ],\n  \"textures\": [", key.c_str(), out.cameras.c_str(), out.meshes.c_str());
        for (size_t t = 0; t < model.textures.size(); t++) {
            const auto& img = model.images[model.textures[t].source];
            Fnv h;
            h.bytes(img.image.data(), img.image.size());
            printf("%s{\"width\": %d, \"height\": %d, \"component\": %d, \"bits\": %d, \"hash\": \"%016llx\"}", t ? ", " : "", img.width, img.height,
                   img.component, img.bits, h.h);
        }
        printf("]\n }%s\n", a + 3 < argc ? "," : "");
    }
    printf("}\n");
    return 0;
}
