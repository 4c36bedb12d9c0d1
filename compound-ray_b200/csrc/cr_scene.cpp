// cr_scene.cpp -- glTF subset loader, .eye parser, pose arithmetic, hit-geometry queries.
// See cr_scene.h for the reference functions this replaces.  All float arithmetic is written
// one rounding per operation (compile with -ffp-contract=off) so the flattened buffers are
// bit-comparable with the CPU checker's independent numpy loader.
#include "cr_scene.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "cr_jpeg.h"
#include "cr_json.h"
#include "cr_math.h"

namespace cr {

// ------------------------------------------------------------------------------------------
// 4x4 row-major float matrix, sutil conventions (sutil/Matrix.h:344-359, 472-490, 677-700)
// ------------------------------------------------------------------------------------------
namespace {

struct Mat4 {
    float m[16];
    static Mat4 identity()
    {
        Mat4 r;
        for (int i = 0; i < 16; i++) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
        return r;
    }
};

Mat4 mul(const Mat4& a, const Mat4& b)
{
    Mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float sum = 0.0f;
            for (int k = 0; k < 4; k++) {
                const float p = a.m[i * 4 + k] * b.m[k * 4 + j];
                sum = sum + p;
            }
            r.m[i * 4 + j] = sum;
        }
    return r;
}

// rows 0..2 of M * (x,y,z,w), summed left to right
Float3 mulPoint(const Mat4& M, float x, float y, float z, float w)
{
    Float3 r;
    r.x = ((M.m[0] * x + M.m[1] * y) + M.m[2] * z) + M.m[3] * w;
    r.y = ((M.m[4] * x + M.m[5] * y) + M.m[6] * z) + M.m[7] * w;
    r.z = ((M.m[8] * x + M.m[9] * y) + M.m[10] * z) + M.m[11] * w;
    return r;
}

// world = ((m0*x + m1*y) + m2*z) + m3  (the w = 1 product is exact, so this equals mulPoint(...,1))
Float3 xformVertex(const Mat4& M, float x, float y, float z)
{
    Float3 r;
    r.x = ((M.m[0] * x + M.m[1] * y) + M.m[2] * z) + M.m[3];
    r.y = ((M.m[4] * x + M.m[5] * y) + M.m[6] * z) + M.m[7];
    r.z = ((M.m[8] * x + M.m[9] * y) + M.m[10] * z) + M.m[11];
    return r;
}

// sutil/Quaternion.h:239-269 with the (w,x,y,z) constructor order of MulticamScene.cpp:183-188
Mat4 quatMatrix(float qw, float qx, float qy, float qz)
{
    Mat4 r = Mat4::identity();
    float* m = r.m;
    m[0] = 1.0f - 2.0f * qy * qy - 2.0f * qz * qz;
    m[1] = 2.0f * qx * qy - 2.0f * qz * qw;
    m[2] = 2.0f * qx * qz + 2.0f * qy * qw;
    m[4] = 2.0f * qx * qy + 2.0f * qz * qw;
    m[5] = 1.0f - 2.0f * qx * qx - 2.0f * qz * qz;
    m[6] = 2.0f * qy * qz - 2.0f * qx * qw;
    m[8] = 2.0f * qx * qz - 2.0f * qy * qw;
    m[9] = 2.0f * qy * qz + 2.0f * qx * qw;
    m[10] = 1.0f - 2.0f * qx * qx - 2.0f * qy * qy;
    return r;
}

Mat4 nodeTransform(const Mat4& parent, const Json& node)
{
    Mat4 T = Mat4::identity(), R = Mat4::identity(), S = Mat4::identity(), M = Mat4::identity();
    const Json& t = node["translation"];
    if (t.isArray() && t.size() == 3) {
        T.m[3] = static_cast<float>(t[0].number());
        T.m[7] = static_cast<float>(t[1].number());
        T.m[11] = static_cast<float>(t[2].number());
    }
    const Json& q = node["rotation"];
    if (q.isArray() && q.size() == 4)
        R = quatMatrix(static_cast<float>(q[3].number()), static_cast<float>(q[0].number()),
                       static_cast<float>(q[1].number()), static_cast<float>(q[2].number()));
    const Json& s = node["scale"];
    if (s.isArray() && s.size() == 3) {
        S.m[0] = static_cast<float>(s[0].number());
        S.m[5] = static_cast<float>(s[1].number());
        S.m[10] = static_cast<float>(s[2].number());
    }
    const Json& mj = node["matrix"];
    if (mj.isArray() && mj.size() == 16)   // glTF is column-major: transpose (MulticamScene.cpp:203)
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) M.m[r * 4 + c] = static_cast<float>(mj[static_cast<size_t>(c * 4 + r)].number());
    return mul(mul(mul(mul(parent, M), T), R), S);
}

void aabbTransform(const Mat4& M, const Float3& bmin, const Float3& bmax, Float3& omin, Float3& omax)
{
    // sutil/Aabb.h:351-371
    omin = {INFINITY, INFINITY, INFINITY};
    omax = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = 0; i < 8; i++) {
        const Float3 c = mulPoint(M, (i & 4) ? bmax.x : bmin.x, (i & 2) ? bmax.y : bmin.y, (i & 1) ? bmax.z : bmin.z, 1.0f);
        omin.x = fminf(omin.x, c.x); omin.y = fminf(omin.y, c.y); omin.z = fminf(omin.z, c.z);
        omax.x = fmaxf(omax.x, c.x); omax.y = fmaxf(omax.y, c.y); omax.z = fmaxf(omax.z, c.z);
    }
}

bool extraIsTrue(const Json& extras, const char* key)   // MulticamScene.cpp:131-146
{
    const Json& v = extras[key];
    if (v.isBool()) return v.b;
    if (v.isString()) {
        std::string s = v.str;
        std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
        return s == "true";
    }
    return false;
}

std::vector<uint8_t> readFileBytes(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open '" + path + "'");
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

std::vector<uint8_t> loadUri(const std::string& uri, const std::string& dir)
{
    if (uri.compare(0, 5, "data:") == 0) {
        const size_t comma = uri.find(',');
        if (comma == std::string::npos) throw std::runtime_error("malformed data: URI");
        return base64Decode(uri.data() + comma + 1, uri.size() - comma - 1);
    }
    return readFileBytes(dir + uri);
}

struct Model {
    Json j;
    std::string dir;
    std::vector<std::vector<uint8_t>> buffers;

    struct View {
        const uint8_t* base = nullptr;
        size_t count = 0, stride = 0;
        int componentType = 0, ncomp = 0;
        const Json* accessor = nullptr;
    };
    static int componentSize(int ct)
    {
        switch (ct) {
            case 5120: case 5121: return 1;
            case 5122: case 5123: return 2;
            case 5125: case 5126: return 4;
            default: throw std::runtime_error("gltf accessor component type not supported");
        }
    }
    static int typeComponents(const std::string& t)
    {
        if (t == "SCALAR") return 1;
        if (t == "VEC2") return 2;
        if (t == "VEC3") return 3;
        if (t == "VEC4") return 4;
        if (t == "MAT4") return 16;
        throw std::runtime_error("gltf accessor type not supported: " + t);
    }
    // A JSON number used as a size or offset: a non-negative integer below 2^53, or the file is rejected
    // (casting a negative or huge double to size_t is undefined, and would wrap the range checks below).
    static size_t sizeField(const Json& v, double dflt, const char* what)
    {
        const double d = v.number(dflt);
        if (!(d >= 0.0) || d > 9007199254740992.0 || d != static_cast<double>(static_cast<uint64_t>(d)))
            throw std::runtime_error(std::string("gltf: '") + what + "' is not a valid size");
        return static_cast<size_t>(d);
    }
    const std::vector<uint8_t>& bufferOf(const Json& bv) const
    {
        const int b = bv["buffer"].integer(-1);
        if (b < 0 || static_cast<size_t>(b) >= buffers.size()) throw std::runtime_error("gltf: bufferView without a valid buffer");
        return buffers[static_cast<size_t>(b)];
    }
    View view(int accessorIdx) const
    {
        if (accessorIdx < 0) throw std::runtime_error("bad accessor index");
        const Json& acc = j["accessors"][static_cast<size_t>(accessorIdx)];
        if (!acc.isObject()) throw std::runtime_error("bad accessor index");
        const int bvIdx = acc["bufferView"].integer(-1);
        if (bvIdx < 0) throw std::runtime_error("accessor without bufferView");
        const Json& bv = j["bufferViews"][static_cast<size_t>(bvIdx)];
        if (!bv.isObject()) throw std::runtime_error("accessor without bufferView");
        View v;
        v.accessor = &acc;
        v.componentType = acc["componentType"].integer();
        v.ncomp = typeComponents(acc["type"].string());
        v.count = sizeField(acc["count"], 0, "accessor.count");
        const size_t elem = static_cast<size_t>(componentSize(v.componentType) * v.ncomp);
        v.stride = sizeField(bv["byteStride"], 0, "bufferView.byteStride");
        if (!v.stride) v.stride = elem;
        const size_t bvOff = sizeField(bv["byteOffset"], 0, "bufferView.byteOffset");
        const size_t accOff = sizeField(acc["byteOffset"], 0, "accessor.byteOffset");
        const auto& buf = bufferOf(bv);
        // off + stride*(count-1) + elem <= size, evaluated without any sum that could wrap
        if (bvOff > buf.size() || accOff > buf.size() - bvOff) throw std::runtime_error("accessor exceeds buffer");
        const size_t off = bvOff + accOff;
        if (v.count) {
            if (elem > buf.size() - off) throw std::runtime_error("accessor exceeds buffer");
            if (v.count - 1 > (buf.size() - off - elem) / v.stride) throw std::runtime_error("accessor exceeds buffer");
        }
        v.base = buf.data() + off;
        return v;
    }
};

float readComponentAsFloat(const uint8_t* p, int ct)
{
    switch (ct) {
        case 5126: { float f; memcpy(&f, p, 4); return f; }
        case 5125: { uint32_t u; memcpy(&u, p, 4); return static_cast<float>(u); }
        case 5123: { uint16_t u; memcpy(&u, p, 2); return static_cast<float>(u); }
        case 5122: { int16_t u; memcpy(&u, p, 2); return static_cast<float>(u); }
        case 5121: return static_cast<float>(*p);
        default: return static_cast<float>(*reinterpret_cast<const int8_t*>(p));
    }
}
uint32_t readIndex(const uint8_t* p, int ct)
{
    switch (ct) {
        case 5125: { uint32_t u; memcpy(&u, p, 4); return u; }
        case 5123: { uint16_t u; memcpy(&u, p, 2); return u; }
        case 5121: return *p;
        default: throw std::runtime_error("index component type not supported");
    }
}

struct LoadCtx {
    const Model& md;
    HostScene& sc;
    bool verbose;
    struct Material { float baseColor[4] = {1, 1, 1, 1}; int tex = -1; };
    std::vector<Material> materials;
    size_t visits = 0;      // node visits so far: a glTF node graph is a forest, so this stays <= the node count
};

// glTF requires a strict tree; a file whose "children" form a cycle or a heavily shared DAG is rejected instead
// of recursing without bound.
constexpr int kMaxNodeDepth = 256;
constexpr size_t kMaxNodeVisits = size_t(1) << 22;

void processNode(LoadCtx& cx, const Json& node, const Mat4& parent, int depth = 0)
{
    if (depth > kMaxNodeDepth || ++cx.visits > kMaxNodeVisits) throw std::runtime_error("gltf: node hierarchy is not a tree (cycle or runaway sharing)");
    const Json& J = cx.md.j;
    const Mat4 xf = nodeTransform(parent, node);

    if (node.has("camera")) {                                            // MulticamScene.cpp:207-328
        const Json& cam = J["cameras"][static_cast<size_t>(node["camera"].integer())];
        HostCamera hc;
        hc.name = cam["name"].string();
        const Float3 up = mulPoint(xf, 0.0f, 1.0f, 0.0f, 0.0f);
        const Float3 fwd = mulPoint(xf, 0.0f, 0.0f, -1.0f, 0.0f);
        const Float3 right = mulPoint(xf, 1.0f, 0.0f, 0.0f, 0.0f);
        hc.pose.pos = mulPoint(xf, 0.0f, 0.0f, 0.0f, 1.0f);
        hc.pose.ax = right; hc.pose.ay = up; hc.pose.az = fwd;           // setLocalSpace(right, up, forward)
        const Json& extras = cam["extras"];
        if (cx.verbose) std::cout << "[PyEye] camera '" << hc.name << "' type " << cam["type"].string() << std::endl;
        if (cam["type"].string() == "orthographic") {
            hc.kind = CAM_ORTHOGRAPHIC;
            hc.scale[0] = static_cast<float>(cam["orthographic"]["xmag"].number());
            hc.scale[1] = static_cast<float>(cam["orthographic"]["ymag"].number());
            hc.scale[2] = 0.0f;
            cx.sc.cameras.push_back(hc);
            return;
        }
        if (extraIsTrue(extras, "panoramic")) {
            hc.kind = CAM_PANORAMIC;
            hc.scale[0] = 0.0f; hc.scale[1] = 0.0f; hc.scale[2] = 0.0f;   // startRadius 0 (PanoramicCamera.cpp:16)
            cx.sc.cameras.push_back(hc);
            return;
        }
        if (extraIsTrue(extras, "compound-eye")) {
            const std::string eyePath = extras["compound-structure"].string();
            const std::string proj = extras["compound-projection"].string();
            if (eyePath.empty()) { std::cerr << "ERROR: Eye data path empty or non-existant." << std::endl; return; }
            if (proj.empty()) { std::cerr << "ERROR: Projection shader specifier empty or non-existant." << std::endl; return; }
            std::string used = eyePath;
            {
                std::ifstream probe(used);
                if (!probe.is_open()) {
                    std::cerr << "WARNING: Unable to open \"" << eyePath << "\", attempting to open at relative address..." << std::endl;
                    used = cx.md.dir + eyePath;
                    std::ifstream probe2(used);
                    if (!probe2.is_open()) {
                        std::cerr << "ERROR: Unable to open \"" << used << "\", read cancelled." << std::endl;
                        return;                                           // camera silently not added (:276-279)
                    }
                }
            }
            hc.ommatidia = readEyeFile(used);
            if (hc.ommatidia.empty()) { std::cerr << "  ERROR: Zero ommatidia loaded." << std::endl; return; }
            hc.kind = CAM_COMPOUND;
            hc.projection = proj;
            hc.eyePath = used;
            cx.sc.cameras.push_back(hc);
            return;
        }
        // perspective: ctor scale (10,10,1) then setYFOV (cameras/PerspectiveCamera.cpp:3-23)
        hc.kind = CAM_PERSPECTIVE;
        const float yfovDeg = static_cast<float>(cam["perspective"]["yfov"].number()) * 180.0f / static_cast<float>(M_PI);
        const float yfov = yfovDeg / 180 * crm::kPi;
        hc.scale[2] = 1.0f;
        hc.scale[1] = tanf(yfov / 2.0f) * hc.scale[2];
        hc.scale[0] = hc.scale[1] * 1.0f;
        cx.sc.cameras.push_back(hc);
        return;
    }

    if (node.has("mesh")) {
        const Json& gm = J["meshes"][static_cast<size_t>(node["mesh"].integer())];
        const bool isHitbox = extraIsTrue(gm["extras"], "hitbox");       // MulticamScene.cpp:329
        HitboxMesh hb;
        if (isHitbox) {
            hb.name = gm["name"].string();
            memcpy(hb.xform, xf.m, sizeof hb.xform);
        }
        const Json& prims = gm["primitives"];
        for (size_t pi = 0; pi < prims.size(); pi++) {
            const Json& prim = prims[pi];
            if (prim.has("mode") && prim["mode"].integer() != 4) {
                std::cerr << "\tNon-triangle primitive: skipping\n";
                continue;
            }
            const Json& attrs = prim["attributes"];
            if (!attrs.has("POSITION")) throw std::runtime_error("primitive without POSITION");
            const Model::View pos = cx.md.view(attrs["POSITION"].integer());
            if (pos.componentType != 5126 || pos.ncomp != 3) throw std::runtime_error("POSITION must be float VEC3");
            std::vector<uint32_t> idx;
            if (prim.has("indices")) {
                const Model::View iv = cx.md.view(prim["indices"].integer());
                idx.resize(iv.count);
                for (size_t i = 0; i < iv.count; i++) idx[i] = readIndex(iv.base + i * iv.stride, iv.componentType);
            } else {
                idx.resize(pos.count);
                for (size_t i = 0; i < pos.count; i++) idx[i] = static_cast<uint32_t>(i);
            }
            const size_t ntri = idx.size() / 3;
            for (size_t i = 0; i < ntri * 3; i++)
                if (idx[i] >= pos.count) throw std::runtime_error("index out of range");

            if (isHitbox) {
                for (size_t i = 0; i < ntri * 3; i++) {
                    float p[3];
                    memcpy(p, pos.base + idx[i] * pos.stride, 12);
                    hb.tris.insert(hb.tris.end(), p, p + 3);
                }
                continue;
            }

            MeshGroup mg;
            mg.name = gm["name"].string();
            mg.firstTri = static_cast<uint32_t>(cx.sc.triangleCount());
            mg.nTris = static_cast<uint32_t>(ntri);
            mg.firstVert = static_cast<uint32_t>(cx.sc.vertexCount());
            mg.nVerts = static_cast<uint32_t>(pos.count);
            mg.material = prim.has("material") ? prim["material"].integer(-1) : -1;
            if (mg.material >= 0 && static_cast<size_t>(mg.material) < cx.materials.size()) {
                memcpy(mg.baseColor, cx.materials[static_cast<size_t>(mg.material)].baseColor, sizeof mg.baseColor);
                mg.texture = cx.materials[static_cast<size_t>(mg.material)].tex;
            }
            // vertices -> world space
            Float3 omin{INFINITY, INFINITY, INFINITY}, omax{-INFINITY, -INFINITY, -INFINITY};
            for (size_t v = 0; v < pos.count; v++) {
                float p[3];
                memcpy(p, pos.base + v * pos.stride, 12);
                omin.x = fminf(omin.x, p[0]); omin.y = fminf(omin.y, p[1]); omin.z = fminf(omin.z, p[2]);
                omax.x = fmaxf(omax.x, p[0]); omax.y = fmaxf(omax.y, p[1]); omax.z = fmaxf(omax.z, p[2]);
                const Float3 w = xformVertex(xf, p[0], p[1], p[2]);
                cx.sc.positions.push_back(w.x); cx.sc.positions.push_back(w.y); cx.sc.positions.push_back(w.z);
            }
            // accessor min/max when present (MulticamScene.cpp:376-389), else computed
            const Json& amin = (*pos.accessor)["min"];
            const Json& amax = (*pos.accessor)["max"];
            if (amin.isArray() && amin.size() == 3 && amax.isArray() && amax.size() == 3) {
                omin = {static_cast<float>(amin[0].number()), static_cast<float>(amin[1].number()), static_cast<float>(amin[2].number())};
                omax = {static_cast<float>(amax[0].number()), static_cast<float>(amax[1].number()), static_cast<float>(amax[2].number())};
            }
            aabbTransform(xf, omin, omax, mg.wmin, mg.wmax);
            // TEXCOORD_0
            if (attrs.has("TEXCOORD_0")) {
                const Model::View uv = cx.md.view(attrs["TEXCOORD_0"].integer());
                if (uv.componentType == 5126 && uv.ncomp == 2 && uv.count >= pos.count) {
                    mg.hasUV = 1;
                    cx.sc.anyUV = true;
                    for (size_t v = 0; v < pos.count; v++) {
                        float t[2];
                        memcpy(t, uv.base + v * uv.stride, 8);
                        cx.sc.uvs.push_back(t[0]); cx.sc.uvs.push_back(t[1]);
                    }
                }
            }
            if (!mg.hasUV) cx.sc.uvs.resize(cx.sc.uvs.size() + 2 * pos.count, 0.0f);
            // COLOR_0: VEC4 of float / u16 / u8 (MulticamScene.cpp:435-518, cuda/LocalGeometry.h:107-150).
            // The reference's u8 path is unreachable (bufferViewFromGLTF throws, :93-99); it is accepted
            // here with the scaling its shader intends (x * 1/255).
            if (attrs.has("COLOR_0")) {
                const Model::View cv = cx.md.view(attrs["COLOR_0"].integer());
                if (cv.ncomp != 4) std::cerr << "\t\t\tWarning: Vertex colours are not of type vec4. Ignoring vertex colours.\n";
                else if ((cv.componentType == 5126 || cv.componentType == 5123 || cv.componentType == 5121) && cv.count >= pos.count) {
                    mg.colorType = cv.componentType;
                    cx.sc.anyColor = true;
                    const float inv = cv.componentType == 5123 ? 1.0f / 65535.0f : (cv.componentType == 5121 ? 1.0f / 255.0f : 1.0f);
                    const int cs = Model::componentSize(cv.componentType);
                    for (size_t v = 0; v < pos.count; v++)
                        for (int c = 0; c < 4; c++) {
                            float f = readComponentAsFloat(cv.base + v * cv.stride + static_cast<size_t>(c * cs), cv.componentType);
                            if (cv.componentType != 5126) f = f * inv;    // sutil `/=` multiplies by the reciprocal
                            cx.sc.colors.push_back(f);
                        }
                }
            }
            if (mg.colorType == -1) cx.sc.colors.resize(cx.sc.colors.size() + 4 * pos.count, 0.0f);
            for (size_t i = 0; i < ntri * 3; i++) cx.sc.indices.push_back(mg.firstVert + idx[i]);
            cx.sc.triMesh.resize(cx.sc.triMesh.size() + ntri, static_cast<uint32_t>(cx.sc.meshes.size()));
            if (cx.verbose)
                std::cout << "[PyEye] mesh '" << mg.name << "': " << ntri << " triangles, colours " << mg.colorType
                          << ", uv " << mg.hasUV << ", texture " << mg.texture << std::endl;
            cx.sc.meshes.push_back(mg);
        }
        if (isHitbox && !hb.tris.empty()) {
            hb.omin = {hb.tris[0], hb.tris[1], hb.tris[2]};
            hb.omax = hb.omin;
            for (size_t i = 0; i < hb.tris.size(); i += 3) {
                hb.omin.x = fminf(hb.omin.x, hb.tris[i]); hb.omin.y = fminf(hb.omin.y, hb.tris[i + 1]); hb.omin.z = fminf(hb.omin.z, hb.tris[i + 2]);
                hb.omax.x = fmaxf(hb.omax.x, hb.tris[i]); hb.omax.y = fmaxf(hb.omax.y, hb.tris[i + 1]); hb.omax.z = fmaxf(hb.omax.z, hb.tris[i + 2]);
            }
            aabbTransform(xf, hb.omin, hb.omax, hb.wmin, hb.wmax);
            cx.sc.hitboxes.push_back(std::move(hb));
        }
        return;
    }

    const Json& children = node["children"];                            // MulticamScene.cpp:519-525
    for (size_t i = 0; i < children.size(); i++)
        processNode(cx, J["nodes"][static_cast<size_t>(children[i].integer())], xf, depth + 1);
}

}  // namespace

// ------------------------------------------------------------------------------------------
std::vector<Ommatidium> readEyeFile(const std::string& path)
{
    std::ifstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open eye file '" + path + "'");
    std::vector<Ommatidium> out;
    std::string line;
    while (std::getline(f, line)) {
        // split on single spaces, dropping empty pieces (MulticamScene.cpp:147-163), then 8 x stof
        float v[8];
        int n = 0;
        size_t i = 0;
        while (i < line.size() && n < 8) {
            while (i < line.size() && line[i] == ' ') i++;
            if (i >= line.size()) break;
            size_t e = line.find(' ', i);
            if (e == std::string::npos) e = line.size();
            v[n++] = std::stof(line.substr(i, e - i));
            i = e;
        }
        if (n < 8) throw std::runtime_error("malformed .eye line (need 8 floats): '" + line + "'");
        out.push_back({v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]});
    }
    return out;
}

ImageRGBA8 decodeImageFile(const std::string& path)
{
    const std::vector<uint8_t> bytes = readFileBytes(path);
    return decodeImage(bytes.data(), bytes.size());
}

HostScene loadGltfScene(const std::string& path, bool verbose)
{
    Model md;
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("Failed to load GLTF scene '" + path + "': file not found");
        std::stringstream ss;
        ss << f.rdbuf();
        md.j = Json::parse(ss.str());
    }
    const size_t slash = path.find_last_of("/\\");
    md.dir = (slash == std::string::npos) ? std::string() : path.substr(0, slash + 1);

    HostScene sc;
    sc.path = path;
    const Json& J = md.j;

    const Json& scenes = J["scenes"];
    for (size_t i = 0; i < scenes.size(); i++) {
        const std::string bg = scenes[i]["extras"]["background-shader"].string();
        if (!bg.empty()) sc.missShaderName = bg;
    }
    if (sc.missShaderName == "simple_sky") sc.missShader = 1;
    else if (sc.missShaderName == "default_background") sc.missShader = 0;
    else throw std::runtime_error("unknown background-shader '" + sc.missShaderName + "'");

    const Json& buffers = J["buffers"];
    for (size_t i = 0; i < buffers.size(); i++) md.buffers.push_back(loadUri(buffers[i]["uri"].string(), md.dir));

    std::vector<ImageRGBA8> images;
    const Json& imgs = J["images"];
    for (size_t i = 0; i < imgs.size(); i++) {
        std::vector<uint8_t> bytes;
        if (imgs[i].has("bufferView")) {
            const int bvIdx = imgs[i]["bufferView"].integer(-1);
            if (bvIdx < 0) throw std::runtime_error("image with a bad bufferView index");
            const Json& bv = J["bufferViews"][static_cast<size_t>(bvIdx)];
            if (!bv.isObject()) throw std::runtime_error("image with a bad bufferView index");
            const auto& buf = md.bufferOf(bv);
            const size_t off = Model::sizeField(bv["byteOffset"], 0, "bufferView.byteOffset");
            const size_t len = Model::sizeField(bv["byteLength"], 0, "bufferView.byteLength");
            if (off > buf.size() || len > buf.size() - off) throw std::runtime_error("image bufferView exceeds buffer");
            bytes.assign(buf.begin() + static_cast<long>(off), buf.begin() + static_cast<long>(off + len));
        } else {
            bytes = loadUri(imgs[i]["uri"].string(), md.dir);
        }
        images.push_back(decodeImage(bytes.data(), bytes.size()));
    }
    const Json& texs = J["textures"];
    for (size_t i = 0; i < texs.size(); i++) {
        const int src = texs[i]["source"].integer(-1);
        if (src < 0 || static_cast<size_t>(src) >= images.size()) throw std::runtime_error("texture without image source");
        sc.textures.push_back(images[static_cast<size_t>(src)]);
    }

    LoadCtx cx{md, sc, verbose, {}, 0};
    const Json& mats = J["materials"];
    for (size_t i = 0; i < mats.size(); i++) {
        LoadCtx::Material m;
        const Json& pbr = mats[i]["pbrMetallicRoughness"];
        const Json& bc = pbr["baseColorFactor"];
        if (bc.isArray() && bc.size() == 4)
            for (size_t c = 0; c < 4; c++) m.baseColor[c] = static_cast<float>(bc[c].number());
        if (pbr.has("baseColorTexture")) m.tex = pbr["baseColorTexture"]["index"].integer(-1);
        cx.materials.push_back(m);
    }

    const Json& nodes = J["nodes"];
    std::vector<char> isRoot(nodes.size(), 1);
    for (size_t i = 0; i < nodes.size(); i++) {
        const Json& ch = nodes[i]["children"];
        for (size_t c = 0; c < ch.size(); c++) {
            const int k = ch[c].integer(-1);
            if (k >= 0 && static_cast<size_t>(k) < nodes.size()) isRoot[static_cast<size_t>(k)] = 0;
        }
    }
    for (size_t i = 0; i < nodes.size(); i++)
        if (isRoot[i]) processNode(cx, nodes[i], Mat4::identity());
    return sc;
}

// ------------------------------------------------------------------------------------------
// Pose arithmetic (cameras/DataRecordCamera.h:49-87)
// ------------------------------------------------------------------------------------------
static inline Float3 v3(float x, float y, float z) { return Float3{x, y, z}; }
static inline Float3 vadd(Float3 a, Float3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline Float3 vscale(Float3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline float vdot(Float3 a, Float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline Float3 vcross(Float3 a, Float3 b)
{ return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline Float3 vnormalize(Float3 v) { const float inv = 1.0f / sqrtf(vdot(v, v)); return vscale(v, inv); }

void poseReset(Pose& p) { p = Pose(); }

Float3 poseTransformToLocal(const Pose& p, Float3 v)
{ return vadd(vadd(vscale(p.ax, v.x), vscale(p.ay, v.y)), vscale(p.az, v.z)); }

static Float3 rotatePointNormalised(Float3 pt, float angle, Float3 axis)
{
    const Float3 n = vnormalize(axis);
    float sn, cs;
    crm::sincos(angle, sn, cs);
    return vadd(vadd(vscale(pt, cs), vscale(vcross(n, pt), sn)), vscale(n, (1 - cs) * vdot(n, pt)));
}
void poseRotateAround(Pose& p, float angle, Float3 axis)
{
    p.ax = rotatePointNormalised(p.ax, angle, axis);
    p.ay = rotatePointNormalised(p.ay, angle, axis);
    p.az = rotatePointNormalised(p.az, angle, axis);
}
void poseRotateLocallyAround(Pose& p, float angle, Float3 localAxis)
{ poseRotateAround(p, angle, poseTransformToLocal(p, localAxis)); }
void poseMove(Pose& p, Float3 step) { p.pos = vadd(p.pos, step); }
void poseMoveLocally(Pose& p, Float3 localStep) { poseMove(p, poseTransformToLocal(p, localStep)); }

// ------------------------------------------------------------------------------------------
// Point-in-hitbox parity test (sutil/hitscanprocessing.cpp:20-83): cast along +x from just outside
// the object AABB to the query point (object space) and count crossings.  The world point is taken
// to object space with w = 0 (make_float4(float3) sets w = 0, sutil/vec_math.h:596-599), so the
// node translation is ignored -- reference quirk kept.
// ------------------------------------------------------------------------------------------
static bool invert4(const float* m, float* dst)
{
    // general cofactor inverse (sutil/Matrix.h:596-640 computes the same adjugate / determinant)
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    if (det == 0.0f) return false;
    const float d = 1.0f / det;
    for (int i = 0; i < 16; i++) dst[i] = inv[i] * d;
    return true;
}

bool pointInsideHitbox(const HitboxMesh& hb, Float3 wp)
{
    float inv[16];
    if (!invert4(hb.xform, inv)) return false;
    Mat4 I;
    memcpy(I.m, inv, sizeof inv);
    const Float3 target = mulPoint(I, wp.x, wp.y, wp.z, 0.0f);
    Float3 start = target;
    start.x = hb.omin.x - 1.0f;
    const Float3 dir = vnormalize(v3(target.x - start.x, target.y - start.y, target.z - start.z));
    unsigned crossings = 0;
    for (size_t i = 0; i + 8 < hb.tris.size(); i += 9) {
        const Float3 p0 = v3(hb.tris[i], hb.tris[i + 1], hb.tris[i + 2]);
        const Float3 p1 = v3(hb.tris[i + 3], hb.tris[i + 4], hb.tris[i + 5]);
        const Float3 p2 = v3(hb.tris[i + 6], hb.tris[i + 7], hb.tris[i + 8]);
        const Float3 nrm = vnormalize(vcross(v3(p1.x - p0.x, p1.y - p0.y, p1.z - p0.z), v3(p2.x - p0.x, p2.y - p0.y, p2.z - p0.z)));
        const float denom = vdot(nrm, dir);
        if (denom == 0) continue;
        const float dist = vdot(v3(p0.x - start.x, p0.y - start.y, p0.z - start.z), nrm) / denom;
        if (dist == 0) continue;
        const Float3 hit = vadd(start, vscale(dir, dist));
        if (dist < 0 || hit.x > target.x) continue;
        const Float3 pts[3] = {p0, p1, p2};
        bool inside = true;
        for (int e = 0; e < 3 && inside; e++) {
            const Float3 a = pts[e], b = pts[(e + 1) % 3];
            const Float3 edge = v3(b.x - a.x, b.y - a.y, b.z - a.z);
            const Float3 rel = v3(hit.x - a.x, hit.y - a.y, hit.z - a.z);
            if (vdot(nrm, vcross(edge, rel)) < 0) inside = false;
        }
        if (inside) crossings++;
    }
    return (crossings % 2u) == 1u;
}

}  // namespace cr
