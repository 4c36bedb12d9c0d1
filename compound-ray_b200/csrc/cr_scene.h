// cr_scene.h -- host-side scene model: glTF subset -> flattened world-space SoA buffers.
//
// Replaces libEyeRenderer3/MulticamScene.{h,cpp} loadScene/processGLTFNode
// (MulticamScene.cpp:165-526, 531-736) and the camera classes' pose state
// (cameras/DataRecordCamera.h:24-101).  Instead of uploading raw glTF buffers and building one
// OptiX GAS per mesh + an IAS (MulticamScene.cpp:1052-1428), static instances are flattened into
// ONE world-space triangle soup at load so a single LBVH serves every ray.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "cr_image.h"

namespace cr {

struct Float3 { float x, y, z; };

// cameras/CompoundEyeDataTypes.h:22-28 == libEyeRenderer.h:8-14 (OmmatidiumPacket) == one .eye line
struct Ommatidium {
    float px, py, pz;
    float dx, dy, dz;
    float acceptance;   // FWHM, radians
    float focal;        // focalPointOffset
};
static_assert(sizeof(Ommatidium) == 32, "Ommatidium must stay 32 bytes (ABI)");

enum CameraKind : int { CAM_PERSPECTIVE = 0, CAM_PANORAMIC = 1, CAM_ORTHOGRAPHIC = 2, CAM_COMPOUND = 3 };

// RaygenPosedContainer pose (cameras/GenericCameraDataTypes.h:17-42)
struct Pose {
    Float3 pos{0.f, 0.f, 0.f};
    Float3 ax{1.f, 0.f, 0.f}, ay{0.f, 1.f, 0.f}, az{0.f, 0.f, 1.f};
};

struct HostCamera {
    std::string name;
    CameraKind kind = CAM_PERSPECTIVE;
    Pose pose;
    float scale[3] = {10.f, 10.f, 1.f};   // perspective: scale xyz; panoramic: [0] = startRadius; ortho: xy
    std::string projection;               // suffix after "__raygen__compound_projection_"
    std::string eyePath;
    std::vector<Ommatidium> ommatidia;
};

struct MeshGroup {
    std::string name;
    uint32_t firstTri = 0, nTris = 0;
    uint32_t firstVert = 0, nVerts = 0;
    int colorType = -1;        // glTF component type of COLOR_0 (5126/5123/5121) or -1
    int hasUV = 0;
    int material = -1;
    int texture = -1;          // glTF texture index of baseColorTexture or -1
    float baseColor[4] = {1.f, 1.f, 1.f, 1.f};
    Float3 wmin{0, 0, 0}, wmax{0, 0, 0};   // world AABB (accessor min/max through the node transform)
};

struct HitboxMesh {            // MulticamScene.cpp:329-345; sutil/hitscanprocessing.cpp
    std::string name;
    float xform[16];           // row-major node transform
    std::vector<float> tris;   // object space, 9 floats per triangle (p0, p1, p2)
    Float3 omin, omax, wmin, wmax;
};

struct HostScene {
    std::vector<HostCamera> cameras;
    std::vector<MeshGroup> meshes;
    std::vector<HitboxMesh> hitboxes;
    std::vector<float> positions;     // world xyz per vertex
    std::vector<float> uvs;           // 2 per vertex (zeros when the mesh has none)
    std::vector<float> colors;        // 4 per vertex, already scaled to float (zeros when none)
    std::vector<uint32_t> indices;    // 3 global vertex ids per triangle
    std::vector<uint32_t> triMesh;    // mesh group of each triangle
    std::vector<ImageRGBA8> textures; // per glTF *texture* index (sampler is always wrap+bilinear:
                                      // MulticamScene.cpp:801-834)
    int missShader = 0;               // 0 default_background, 1 simple_sky (MulticamScene.cpp:555-565)
    std::string missShaderName = "default_background";
    bool anyUV = false, anyColor = false;
    std::string path;

    size_t triangleCount() const { return indices.size() / 3; }
    size_t vertexCount() const { return positions.size() / 3; }
};

// Loads an ASCII .gltf (embedded base64 or external buffers/images).  Throws std::runtime_error.
HostScene loadGltfScene(const std::string& path, bool verbose);

// Decodes a PNG / baseline-JPEG file to RGBA8 (the texture path of the loader).
ImageRGBA8 decodeImageFile(const std::string& path);

// .eye parsing (MulticamScene.cpp:290-299; data/eyes/eye-specification.txt)
std::vector<Ommatidium> readEyeFile(const std::string& path);

// Host pose arithmetic (cameras/DataRecordCamera.h:49-87)
void poseReset(Pose& p);
void poseRotateAround(Pose& p, float angle, Float3 axis);          // axis normalised inside
void poseRotateLocallyAround(Pose& p, float angle, Float3 localAxis);
void poseMove(Pose& p, Float3 step);
void poseMoveLocally(Pose& p, Float3 localStep);
Float3 poseTransformToLocal(const Pose& p, Float3 v);

// Hit-geometry queries (MulticamScene.cpp:1757-1818; sutil/hitscanprocessing.cpp:20-83)
bool pointInsideHitbox(const HitboxMesh& hb, Float3 worldPoint);

}  // namespace cr
