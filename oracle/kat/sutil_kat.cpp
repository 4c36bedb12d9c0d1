// oracle/kat/sutil_kat.cpp -- TEST INFRASTRUCTURE ONLY.
// Known-answer generator for the loader / pose conventions: compiles against the REFERENCE's own
// sutil headers where they lie (/root/reference/sutil/{Matrix,Quaternion,Aabb,vec_math}.h) and
// evaluates exactly the expressions of libEyeRenderer3/MulticamScene.cpp:165-219 (node transform,
// camera axes) and cameras/DataRecordCamera.h:66-87 (pose rotation) on fixed inputs.  Output is
// committed as tests/golden/sutil_kat.json.  Only buildable where /root/reference exists.
#include <cstdio>
#include <cstring>
#include <vector>
#include <sutil/Matrix.h>
#include <sutil/Quaternion.h>
#include <sutil/Aabb.h>
#include <sutil/vec_math.h>
using namespace sutil;

static unsigned f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static void pm(const char* name, const Matrix4x4& m, bool comma = true)
{
    printf("   \"%s\": [", name);
    for (int i = 0; i < 16; i++) printf("%s%u", i ? ", " : "", f2u(m[i]));
    printf("]%s\n", comma ? "," : "");
}
static void pv(const char* name, float3 v, bool comma = true)
{ printf("   \"%s\": [%u, %u, %u]%s\n", name, f2u(v.x), f2u(v.y), f2u(v.z), comma ? "," : ""); }

struct NodeIn { double t[3]; double r[4]; double s[3]; };

static Matrix4x4 node_xform(const Matrix4x4& parent, const NodeIn& n)
{
    // MulticamScene.cpp:173-205
    const Matrix4x4 translation = Matrix4x4::translate(make_float3((float)n.t[0], (float)n.t[1], (float)n.t[2]));
    const Matrix4x4 rotation = Quaternion((float)n.r[3], (float)n.r[0], (float)n.r[1], (float)n.r[2]).rotationMatrix();
    const Matrix4x4 scale = Matrix4x4::scale(make_float3((float)n.s[0], (float)n.s[1], (float)n.s[2]));
    const Matrix4x4 matrix = Matrix4x4::identity();
    return parent * matrix * translation * rotation * scale;
}

int main()
{
    // inputs: the test-scene "Camera" parent/child pair, the Suzanne node, the natural-standin Plane
    // node and one synthetic non-uniform case
    const NodeIn nodes[] = {
        {{7.358891487121582, 4.958309173583984, 6.925790786743164}, {0.483536034822464, 0.33687159419059753, -0.20870360732078552, 0.7804827094078064}, {1, 1, 1}},
        {{0, 0, 0}, {-0.7071067690849304, 0, 0, 0.7071067690849304}, {1, 1, 1}},
        {{-4.120373725891113, 0, 2.064225435256958}, {0, 0.26009199023246765, 0, 0.965583860874176}, {1, 1, 1}},
        {{0, -14.353710174560547, 0}, {0, 0, 0, 1}, {496.10650634765625, 496.10650634765625, 496.10650634765625}},
        {{1.5, -2.25, 3.125}, {0.18257418583505536, 0.3651483716701107, 0.5477225575051661, 0.7302967433402214}, {2, 0.5, 3}},
    };
    printf("{\n \"nodes\": [\n");
    for (int i = 0; i < 5; i++) {
        Matrix4x4 m = node_xform(Matrix4x4::identity(), nodes[i]);
        printf("  {\n   \"t\": [%.17g, %.17g, %.17g], \"r\": [%.17g, %.17g, %.17g, %.17g], \"s\": [%.17g, %.17g, %.17g],\n",
               nodes[i].t[0], nodes[i].t[1], nodes[i].t[2], nodes[i].r[0], nodes[i].r[1], nodes[i].r[2], nodes[i].r[3],
               nodes[i].s[0], nodes[i].s[1], nodes[i].s[2]);
        pm("xform", m);
        float3 p = make_float3(m * make_float4(0.25f, -1.5f, 3.0f, 1.0f));
        pv("point", p, false);
        printf("  }%s\n", i < 4 ? "," : "");
    }
    printf(" ],\n \"camera\": {\n");
    {
        Matrix4x4 parent = node_xform(Matrix4x4::identity(), nodes[0]);
        Matrix4x4 m = node_xform(parent, nodes[1]);
        pm("xform", m);
        // MulticamScene.cpp:215-219
        pv("up", make_float3(m * make_float4(0.0f, 1.0f, 0.0f, 0.0f)));
        pv("forward", make_float3(m * make_float4(0.0f, 0.0f, -1.0f, 0.0f)));
        pv("right", make_float3(m * make_float4(1.0f, 0.0f, 0.0f, 0.0f)));
        pv("eye", make_float3(m * make_float4(0.0f, 0.0f, 0.0f, 1.0f)), false);
    }
    printf(" },\n \"aabb\": {\n");
    {
        Matrix4x4 m = node_xform(Matrix4x4::identity(), nodes[4]);
        Aabb bb(make_float3(-1.0f, -0.5f, -2.0f), make_float3(1.5f, 0.75f, 0.25f));
        bb.transform(m);
        pv("min", bb.m_min);
        pv("max", bb.m_max, false);
    }
    printf(" },\n \"rotate\": [\n");
    {
        // cameras/DataRecordCamera.h:83-87 (normalised axis) applied as in setCameraPose
        // (libEyeRenderer.cpp:380-388): reset; rotX; rotY; rotZ; move.
        const float angs[3][3] = {{0.3f, -1.1f, 2.5f}, {0.0f, 1.5707964f, 0.0f}, {-0.7f, 0.2f, 0.05f}};
        for (int k = 0; k < 3; k++) {
            float3 ax[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            const float3 waxes[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            for (int w = 0; w < 3; w++)
                for (int a = 0; a < 3; a++) {
                    const float3 na = normalize(waxes[w]);
                    const float ang = angs[k][w];
                    const float3 pt = ax[a];
                    // host libm cos/sin here: values are reference-for-tolerance
                    ax[a] = cosf(ang) * pt + sinf(ang) * cross(na, pt) + (1 - cosf(ang)) * dot(na, pt) * na;
                }
            printf("  {\"angles\": [%.9g, %.9g, %.9g], \"x\": [%.9g, %.9g, %.9g], \"y\": [%.9g, %.9g, %.9g], \"z\": [%.9g, %.9g, %.9g]}%s\n",
                   angs[k][0], angs[k][1], angs[k][2], ax[0].x, ax[0].y, ax[0].z, ax[1].x, ax[1].y, ax[1].z,
                   ax[2].x, ax[2].y, ax[2].z, k < 2 ? "," : "");
        }
    }
    printf(" ]\n}\n");
    return 0;
}
