// oracle/kat/hitscan_kat.cpp -- TEST INFRASTRUCTURE ONLY.
// Known-answer generator for isInsideHitGeometry / getGeometry{Min,Max}Bounds: links the REFERENCE's own
// sutil/hitscanprocessing.cpp where it lies (isPointWithinMesh, calculateObjectAabb,
// calculateWorldAabbUsingTransformAndObjectAabb) and evaluates it on small closed meshes under node
// transforms built exactly as MulticamScene.cpp:173-205 builds them.  The meshes, transforms, query points and
// the reference's answers are committed as tests/golden/hitscan_kat.json; tests/test_host.py feeds the same
// meshes to the product as glTF hitbox scenes.  Only buildable where /root/reference exists.
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include <sutil/hitscanprocessing.h>
#include <sutil/Quaternion.h>
using namespace sutil;

static unsigned f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

struct Case {
    const char* name;
    std::vector<float3> verts;
    std::vector<unsigned> idx;
    double t[3], r[4], s[3];
};

static Matrix4x4 node_xform(const Case& n)   // MulticamScene.cpp:173-205 with an identity parent
{
    const Matrix4x4 translation = Matrix4x4::translate(make_float3((float)n.t[0], (float)n.t[1], (float)n.t[2]));
    const Matrix4x4 rotation = Quaternion((float)n.r[3], (float)n.r[0], (float)n.r[1], (float)n.r[2]).rotationMatrix();
    const Matrix4x4 scale = Matrix4x4::scale(make_float3((float)n.s[0], (float)n.s[1], (float)n.s[2]));
    return Matrix4x4::identity() * Matrix4x4::identity() * translation * rotation * scale;
}

int main()
{
    std::vector<Case> cases;
    {   // axis-aligned cube, scaled and translated (the translation is ignored by the query: w = 0)
        Case c{"cube", {}, {}, {10, 0, 0}, {0, 0, 0, 1}, {2, 2, 2}};
        for (int x = -1; x <= 1; x += 2) for (int y = -1; y <= 1; y += 2) for (int z = -1; z <= 1; z += 2) c.verts.push_back(make_float3(x, y, z));
        const unsigned q[6][4] = {{0, 1, 3, 2}, {4, 6, 7, 5}, {0, 4, 5, 1}, {2, 3, 7, 6}, {0, 2, 6, 4}, {1, 5, 7, 3}};
        for (auto& f : q) { unsigned t[6] = {f[0], f[1], f[2], f[0], f[2], f[3]}; c.idx.insert(c.idx.end(), t, t + 6); }
        cases.push_back(c);
    }
    {   // tetrahedron under a rotation and a non-uniform scale
        Case c{"tetra", {make_float3(1, 1, 1), make_float3(1, -1, -1), make_float3(-1, 1, -1), make_float3(-1, -1, 1)},
               {0, 1, 2, 0, 3, 1, 0, 2, 3, 1, 3, 2}, {0.5, -2, 3}, {0.18257418583505536, 0.3651483716701107, 0.5477225575051661, 0.7302967433402214}, {2, 0.5, 3}};
        cases.push_back(c);
    }
    {   // L-shaped prism (not convex): outline extruded along z
        Case c{"ell", {}, {}, {0, 0, 0}, {0, 0.26009199023246765, 0, 0.965583860874176}, {1.5, 1.5, 1.5}};
        const float o[6][2] = {{0, 0}, {2, 0}, {2, 1}, {1, 1}, {1, 2}, {0, 2}};
        for (int k = 0; k < 2; k++) for (auto& p : o) c.verts.push_back(make_float3(p[0] - 1.0f, p[1] - 1.0f, k ? 0.5f : -0.5f));
        const unsigned cap[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 4}, {0, 4, 5}};
        for (auto& t : cap) { c.idx.insert(c.idx.end(), {t[0], t[2], t[1]}); c.idx.insert(c.idx.end(), {t[0] + 6, t[1] + 6, t[2] + 6}); }
        for (unsigned i = 0; i < 6; i++) { const unsigned j = (i + 1) % 6; c.idx.insert(c.idx.end(), {i, j, j + 6, i, j + 6, i + 6}); }
        cases.push_back(c);
    }
    std::mt19937 rng(12345);
    printf("{\n");
    for (size_t ci = 0; ci < cases.size(); ci++) {
        const Case& c = cases[ci];
        hitscan::TriangleMesh tm;
        tm.name = c.name;
        tm.transform = node_xform(c);
        for (size_t i = 0; i + 2 < c.idx.size(); i += 3) tm.triangles.push_back({c.verts[c.idx[i]], c.verts[c.idx[i + 1]], c.verts[c.idx[i + 2]]});
        hitscan::calculateObjectAabb(tm);
        hitscan::calculateWorldAabbUsingTransformAndObjectAabb(tm);
        printf(" \"%s\": {\n  \"t\": [%.17g, %.17g, %.17g], \"r\": [%.17g, %.17g, %.17g, %.17g], \"s\": [%.17g, %.17g, %.17g],\n", c.name,
               c.t[0], c.t[1], c.t[2], c.r[0], c.r[1], c.r[2], c.r[3], c.s[0], c.s[1], c.s[2]);
        printf("  \"verts\": [");
        for (size_t i = 0; i < c.verts.size(); i++) printf("%s[%.9g, %.9g, %.9g]", i ? ", " : "", c.verts[i].x, c.verts[i].y, c.verts[i].z);
        printf("],\n  \"idx\": [");
        for (size_t i = 0; i < c.idx.size(); i++) printf("%s%u", i ? ", " : "", c.idx[i]);
        printf("],\n  \"world_min\": [%u, %u, %u], \"world_max\": [%u, %u, %u],\n", f2u(tm.worldAabb.m_min.x), f2u(tm.worldAabb.m_min.y),
               f2u(tm.worldAabb.m_min.z), f2u(tm.worldAabb.m_max.x), f2u(tm.worldAabb.m_max.y), f2u(tm.worldAabb.m_max.z));
        // query points: world-space points whose w = 0 image covers the object box and a margin around it
        std::uniform_real_distribution<float> u(-1.6f, 1.6f);
        printf("  \"points\": [");
        std::vector<int> inside;
        for (int k = 0; k < 400; k++) {
            const float3 obj = make_float3(u(rng), u(rng), u(rng));
            float3 w = make_float3(tm.transform * make_float4(obj.x, obj.y, obj.z, 0.0f));      // inverse of the query's w = 0 transform
            printf("%s[%u, %u, %u]", k ? ", " : "", f2u(w.x), f2u(w.y), f2u(w.z));
            inside.push_back(hitscan::isPointWithinMesh(tm, w) ? 1 : 0);
        }
        printf("],\n  \"inside\": [");
        for (size_t k = 0; k < inside.size(); k++) printf("%s%d", k ? ", " : "", inside[k]);
        printf("]\n }%s\n", ci + 1 < cases.size() ? "," : "");
    }
    printf("}\n");
    return 0;
}
