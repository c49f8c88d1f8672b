// Device-side scene layout, exact-arithmetic vector math, keyed-transform
// evaluation and the primitive intersectors of the render core.
//
// PARITY CONTRACT.  Hit decisions must be bit-identical to
// the reference CPU code, so this translation unit is compiled with -fmad=false
// (no FMA contraction), IEEE division and square root, no flush-to-zero, and
//   * every expression keeps the reference's association (dot = (xx' + yy') + zz'),
//   * normalisation DIVIDES by the length (never multiplies by a reciprocal),
//   * min/max use std::min/std::max semantics (first argument wins on NaN),
//   * comparisons keep the reference's polarity so NaNs take the same branch,
//   * identity transforms are still applied (they turn -0.0 into +0.0).
// Each function cites the reference lines it stands for (Rayito_Stage7_QT/...).
#ifndef RAYITO_B200_RT_DEVICE_CUH
#define RAYITO_B200_RT_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "rayito_b200.h"

#define RT_RAY_TMIN 0.0001f      /* RRay.h:23 */
#define RT_RAY_TMAX 1.0e30f      /* RRay.h:28 */

#define RT_NODE_LEAF 0x4u        /* RAccel.h:124 */
#define RT_NODE_AXIS 0x3u        /* RAccel.h:123 */
#define RT_TOKEN_SHAPE 0x80000000u   /* stack entry that names a shape directly (linear-list mode) */

struct V3
{
    float x, y, z;
};

// ---------------------------------------------------------------------------
// Device scene.  All arrays live in one HBM arena owned by RtScene (rt_core.cu).
// ---------------------------------------------------------------------------

// BVH node, reference layout (RAccel.h:136-145), read as two 128-bit loads:
//   q0 = (min.x, min.y, min.z, max.x)   q1 = (max.y, max.z, child|prim, flags)
// Mesh leaves are re-encoded at upload: word 6 = first triangle record of the
// face, flags = LEAF | (triangle count << 3); top-level leaves keep prim = shape.
struct DNode
{
    float4 q0;
    float4 q1;
};

// Transform classes decided at upload (exactly, from the key values):
//   GENERAL    anything
//   TRANSLATE  every rotation key is exactly (1;0,0,0) and every scale key exactly
//              (1,1,1): then rotation(t) and scaling(t) are exactly identity for
//              EVERY t -- lerp(1,1,t) = fl(fl(1-t)+t) = 1 for t in [0,1], and the
//              normalised lerp of identity quaternions is w/sqrt(w*w) = 1 -- so only
//              translation(t) needs evaluating
//   STATIC     TRANSLATE with at most one key: the translation is a constant, kept
//              in the shape record itself
//   RIGID      every scale key exactly (1,1,1), rotations arbitrary: scaling(t) is exactly (1,1,1)
//              for every t by the same argument, so only translation(t) and rotation(t) are evaluated
#define RT_XF_GENERAL 0u
#define RT_XF_TRANSLATE 1u
#define RT_XF_STATIC 2u
#define RT_XF_RIGID 3u

// Per-sample transform cache (rt_render.cuh).  Every ray of a pixel sample carries the sample's
// time, and translation(t) / rotation(t) / scaling(t) are pure functions of it, so the renderer
// evaluates every ANIMATED transform (two or more keys) once per sample, all lanes together, when
// the camera ray is generated, and the traversal / shading kernels read the values back instead of
// re-running the key search, the lerps and the normalised quaternion lerp (a square root and four
// IEEE divisions) every time one of the sample's rays enters the shape -- which they did at 4-12
// of 32 lanes, half of all warp instructions of the top-level passes (profiles/README.md, r02).
// A sample's row holds, per animated transform in upload order: TRANSLATE one float4 (t.xyz, -);
// RIGID two (t.xyz, qw) (qv.xyz, -); GENERAL three (t.xyz, qw) (qv.xyz, s.x) (s.y, s.z, -, -).
#define RT_XF_CACHE_MAX_STRIDE 32u     /* float4 per sample; scenes with more animated data evaluate directly */

struct DShapeMem         // 32 bytes in HBM, two 128-bit loads
{
    uint32_t type_kind;  // RT_SHAPE_* | RT_XF_* << 8 | (traversal kernels may read the cache entry) << 10
                         // | (offset in the per-sample transform cache + 1, 0 = not cached) << 16
    uint32_t geom, xform, material;
    int32_t light;
    float tx, ty, tz;    // translation of a STATIC transform
};

struct DShape            // the same, decoded into registers (load_shape)
{
    uint32_t type, geom, xform, material;
    int32_t light;
    uint32_t xkind;
    uint32_t cache_slot; // offset in the per-sample transform cache + 1, 0 = evaluate directly
    uint32_t trav_cached;// the traversal kernels read the entry too (shapes that are small in the scene)
    float tx, ty, tz;
};

struct DXform
{
    uint32_t first_key, num_keys;
    uint32_t kind;       // RT_XF_*
    uint32_t pad;
};

struct DPlane            // 32 bytes
{
    float px, py, pz;    // m_position
    float nx, ny, nz;    // m_normal (normalised by the constructor)
    float pos_dot_n;     // dot(m_position, m_normal), RScene.h:308
    uint32_t bullseye;
};

struct DSphere           // 16 bytes
{
    float px, py, pz, radius;
};

// Rectangle light with the per-call constants of RLight.h:66,84-87 evaluated once
// at upload with the same float operations (host SSE2 arithmetic is IEEE and
// unfused, so the bits equal what the reference recomputes on every call).
struct DRect             // 80 bytes
{
    float px, py, pz;            // m_position
    float nx, ny, nz;            // cross(side1, side2).normalized()
    float s1x, s1y, s1z;         // side1 normalised
    float s2x, s2y, s2z;         // side2 normalised
    float len1, len2;            // side lengths
    float pos_dot_n;             // dot(m_position, normal)
    float r1x, r1y, r1z;         // raw side1
    float r2x, r2y, r2z;         // raw side2 (padding to 80 B follows)
};

struct DMesh
{
    uint32_t first_node, num_nodes;
    uint32_t first_tri;          // first triangle record of the mesh
    uint32_t first_face;         // global index of face 0
    uint32_t num_faces;
    uint32_t first_cdf;
    float total_area;
    uint32_t pad;
};

// Fan triangle: three float4 (xyz + one id word each) and, separately, its three
// global normal indices.
//   v0.w = face index (mesh-local), v1.w = triangle index inside the face,
//   v2.w = 1 if the face has vertex normals
// One step of the top-level BVH's depth-first walk for one direction octant (rt_split.cuh,
// trace_top_static).  The order in which Bvh::intersect pops nodes depends only on the signs
// of the ray direction (RAccel.h:481-486,540-556), so it is tabulated per octant at upload.
#define RT_WALK_MAX_DEPTH 8      /* deepest node a tabulated walk may contain (= RT_SPLIT_TOPCAP) */
#define RT_WALK_MAX_STEPS 96
#define RT_WALK_TOKEN 0xffffffffu
struct DTopStep
{
    uint32_t node;           // top-level node popped at this step (RT_WALK_TOKEN: linear shape list entry)
    uint32_t word;           // interior: first child; leaf: shape index
    uint32_t flags;          // bits 0-2 node flags (axis, leaf) | depth << 8 | pending entries << 16
    uint32_t pending_depth;  // 4 bits per pending stack entry: its depth
    uint32_t pending_node[RT_WALK_MAX_DEPTH];   // the explicit stack below this node, bottom first
    uint32_t pad[4];
};

struct DScene
{
    uint32_t set_xform;
    uint32_t num_finite, num_infinite;
    uint32_t num_top_nodes;
    uint32_t num_lights;
    uint32_t stage6;             // RT_SEMANTICS_STAGE6 rules (no transforms, first fan triangle wins, ...)

    const DShapeMem* shapes;
    const DNode* top_nodes;
    const DNode* mesh_nodes;
    const float4* tris;          // 3 per triangle
    const uint4* tri_normals;    // 1 per triangle: n0, n1, n2, unused
    const uint32_t* face_first_tri; // per global face (+1): first triangle record
    const float* normals;        // xyz, all meshes
    const DXform* xforms;
    const float* key_time;
    const float* key_scale;
    const float* key_rot;        // w, x, y, z
    const float* key_trans;
    const DPlane* planes;
    const DSphere* spheres;
    const DRect* rects;
    const DMesh* meshes;
    const float* face_area_cdf;
    const RtMaterial* materials;
    const uint32_t* lights;      // shape index per light
    const DTopStep* top_walk;    // [8 octants][top_walk_steps], NULL if the top level is too deep to tabulate
    uint32_t top_walk_steps;
    uint32_t top_walk_levels;    // deepest node of the walk + 2: rows of the per-lane range table
    // per-sample transform cache: the animated transforms in row order (xform index, row offset, kind)
    const uint4* anim;           // x = xform, y = row offset (float4), z = RT_XF_*
    uint32_t num_anim;
    uint32_t anim_stride;        // float4 per sample (0 = no cache)
};

// ---------------------------------------------------------------------------
// Vector math with the reference's operation order (RMath.h:180-360)
// ---------------------------------------------------------------------------

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator/(V3 a, V3 b) { return mk(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float f) { return mk(f * a.x, f * a.y, f * a.z); }
__device__ __forceinline__ V3 operator*(float f, V3 a) { return mk(f * a.x, f * a.y, f * a.z); }
__device__ __forceinline__ V3 operator/(V3 a, float f) { return mk(a.x / f, a.y / f, a.z / f); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }

__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b)
{
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ float length3(V3 a) { return sqrtf(length2(a)); }

// Vector::normalize (RMath.h:194): divide by the length when it is positive;
// returns the old length through *len when asked.
// (out of line: one sqrt and three IEEE divisions, used at a dozen call sites of the
// shading kernels)
__device__ __noinline__ V3 normalized3(V3 a, float* len = nullptr)
{
    float l = length3(a);
    if (len) *len = l;
    if (l > 0) { a.x /= l; a.y /= l; a.z /= l; }
    return a;
}

// std::min / std::max on floats: (b < a) ? b : a and (a < b) ? b : a
__device__ __forceinline__ float std_min(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float std_max(float a, float b) { return (a < b) ? b : a; }

// ---------------------------------------------------------------------------
// Keyed transform (RMath.h:619-941)
// ---------------------------------------------------------------------------

struct TRS
{
    V3 t;        // translation(time)
    float qw;    // rotation(time)
    V3 qv;
    V3 s;        // scaling(time)
    bool absent; // Stage 6: the shape has no transform at all (nothing is applied, -0.0 stays -0.0)
};

// q * v = v + w*t + cross(qv, t), t = 2 cross(qv, v)   (RMath.h:536-549)
__device__ __noinline__ V3 quat_rotate(float qw, V3 qv, V3 v)
{
    V3 t = 2.0f * cross3(qv, v);
    return v + t * qw + cross3(qv, t);
}

// Transform::timeIndex (RMath.h:850-884): binary search for the key before `time`
// and the mix factor towards the next key; mix == 0 selects the key verbatim.
__device__ __forceinline__ uint32_t xform_bracket(const float* __restrict__ times, uint32_t n, float time, float& mix)
{
    uint32_t lo = 0, hi = n - 1;
    if (times[hi] <= time) lo = hi;
    else if (times[lo] >= time) hi = lo;
    while (hi - lo > 0)
    {
        uint32_t mid = (lo + hi) / 2;
        if (time < times[mid]) hi = mid;
        else if (mid > lo) lo = mid;
        else break;
    }
    if (lo == n - 1) mix = 0.0f;
    else if (times[lo] >= time) mix = 0.0f;
    else mix = (time - times[lo]) / (times[lo + 1] - times[lo]);
    return lo;
}

// translation(time), scaling(time), rotation(time) (RMath.h:681-715) evaluated
// ONCE per (ray, shape); the reference re-evaluates them per use, but each is a
// pure function of time so the values are the same.
__device__ __noinline__ TRS xform_eval(const DScene& sc, uint32_t xform, float time)
{
    TRS r;
    r.absent = sc.stage6 != 0;
    DXform x = sc.xforms[xform];
    if (x.num_keys == 0)
    {
        r.t = mk(0.0f, 0.0f, 0.0f);
        r.s = mk(1.0f, 1.0f, 1.0f);
        r.qw = 1.0f;
        r.qv = mk(0.0f, 0.0f, 0.0f);
        return r;
    }
    if (x.kind == RT_XF_TRANSLATE || x.kind == RT_XF_STATIC)
    {
        // translation-only transform: rotation(t) and scaling(t) are exactly identity
        r.s = mk(1.0f, 1.0f, 1.0f);
        r.qw = 1.0f;
        r.qv = mk(0.0f, 0.0f, 0.0f);
        if (x.num_keys == 1)
        {
            const float* T1 = sc.key_trans + 3 * x.first_key;
            r.t = mk(T1[0], T1[1], T1[2]);
            return r;
        }
        float m;
        uint32_t k = x.first_key + xform_bracket(sc.key_time + x.first_key, x.num_keys, time, m);
        const float* Tk = sc.key_trans + 3 * k;
        if (m == 0.0f)
            r.t = mk(Tk[0], Tk[1], Tk[2]);
        else
            r.t = mk(Tk[0], Tk[1], Tk[2]) * (1.0f - m) + mk(Tk[3], Tk[4], Tk[5]) * m;
        return r;
    }
    float mix;
    uint32_t i = x.first_key + xform_bracket(sc.key_time + x.first_key, x.num_keys, time, mix);
    const float* T = sc.key_trans + 3 * i;
    const float* S = sc.key_scale + 3 * i;
    const float* R = sc.key_rot + 4 * i;
    const bool rigid = x.kind == RT_XF_RIGID;       // scaling(t) == (1,1,1) exactly, whatever t
    if (mix == 0.0f)
    {
        r.t = mk(T[0], T[1], T[2]);
        r.s = rigid ? mk(1.0f, 1.0f, 1.0f) : mk(S[0], S[1], S[2]);
        r.qw = R[0];
        r.qv = mk(R[1], R[2], R[3]);
    }
    else
    {
        float om = 1.0f - mix;
        r.t = mk(T[0], T[1], T[2]) * om + mk(T[3], T[4], T[5]) * mix;
        r.s = rigid ? mk(1.0f, 1.0f, 1.0f) : mk(S[0], S[1], S[2]) * om + mk(S[3], S[4], S[5]) * mix;
        // lerp(q1, q2, t) = (q1*(1-t) + q2*t).normalized()   (RMath.h:576-580)
        float w = om * R[0] + mix * R[4];
        V3 v = mk(R[1], R[2], R[3]) * om + mk(R[5], R[6], R[7]) * mix;
        float len = sqrtf(w * w + length2(v));
        if (len > 0) { w /= len; v = v / len; }
        r.qw = w;
        r.qv = v;
    }
    return r;
}

// The transform cache: write one transform's values into a sample's row / read them back
__device__ __forceinline__ void xform_cache_store(float4* row, uint32_t off, uint32_t kind, const TRS& r)
{
    if (kind == RT_XF_TRANSLATE)
    {
        row[off] = make_float4(r.t.x, r.t.y, r.t.z, 0.0f);
        return;
    }
    row[off] = make_float4(r.t.x, r.t.y, r.t.z, r.qw);
    row[off + 1] = make_float4(r.qv.x, r.qv.y, r.qv.z, r.s.x);
    if (kind != RT_XF_RIGID)
        row[off + 2] = make_float4(r.s.y, r.s.z, 0.0f, 0.0f);
}
__device__ __forceinline__ TRS xform_cache_load(const float4* row, uint32_t off, uint32_t kind, bool stage6)
{
    TRS r;
    r.absent = stage6;
    const float4 a = __ldg(row + off);
    r.t = mk(a.x, a.y, a.z);
    r.s = mk(1.0f, 1.0f, 1.0f);
    r.qw = 1.0f;
    r.qv = mk(0.0f, 0.0f, 0.0f);
    if (kind != RT_XF_TRANSLATE)
    {
        const float4 b = __ldg(row + off + 1);
        r.qw = a.w;
        r.qv = mk(b.x, b.y, b.z);
        if (kind != RT_XF_RIGID)
        {
            const float4 c = __ldg(row + off + 2);
            r.s = mk(b.w, c.x, c.y);
        }
    }
    return r;
}

// Exact shortcuts.  Most transforms in a scene have an identity rotation and unit
// scale (the ShapeSet itself, every translated-only shape), yet the reference still
// runs q*v and the division for them.  Skipping that arithmetic is bit-identical
// when it provably cannot change a bit:
//   * dividing by exactly (1,1,1) returns the operand;
//   * rotating by exactly (1; 0,0,0): t = 2 cross(qv, v) is a vector of signed
//     zeros, so v + t*w + cross(qv, t) == v for every component of v that is not
//     itself a zero (x + (+-0) == x).  A zero component may change sign (-0 -> +0,
//     which flips 1/x and the traversal order), so those rays take the full path.
// (Non-finite components are not shortcut-safe either; rays are finite.)
__device__ __forceinline__ bool trs_identity_rotation(const TRS& x)
{
    return x.qw == 1.0f && x.qv.x == 0.0f && x.qv.y == 0.0f && x.qv.z == 0.0f;
}
__device__ __forceinline__ bool trs_unit_scale(const TRS& x)
{
    return x.s.x == 1.0f && x.s.y == 1.0f && x.s.z == 1.0f;
}
__device__ __forceinline__ bool no_zero_component(V3 v)
{
    return v.x != 0.0f && v.y != 0.0f && v.z != 0.0f;
}

__device__ __forceinline__ V3 rotate_exact(float qw, V3 qv, V3 v, bool identity, bool absent = false)
{
    if (absent || (identity && no_zero_component(v)))
        return v;
    return quat_rotate(qw, qv, v);
}

// Component-wise division by a non-unit scale: rare, kept out of line so the three
// IEEE division sequences exist once per kernel
__device__ __noinline__ V3 scale_divide(V3 r, V3 s) { return r / s; }

// Transform::toLocalPoint / toLocalVector (RMath.h:814-827)
__device__ __forceinline__ V3 to_local_point(const TRS& x, V3 p)
{
    V3 r = rotate_exact(x.qw, -x.qv, p - x.t, trs_identity_rotation(x), x.absent);
    return trs_unit_scale(x) ? r : scale_divide(r, x.s);
}
__device__ __forceinline__ V3 to_local_vector(const TRS& x, V3 v)
{
    V3 r = rotate_exact(x.qw, -x.qv, v, trs_identity_rotation(x), x.absent);
    return trs_unit_scale(x) ? r : scale_divide(r, x.s);
}
// fromLocalPoint / fromLocalVector / fromLocalNormal (RMath.h:819-842)
__device__ __forceinline__ V3 from_local_point(const TRS& x, V3 p)
{
    V3 v = trs_unit_scale(x) ? p : p * x.s;        // p * 1 == p
    return rotate_exact(x.qw, x.qv, v, trs_identity_rotation(x), x.absent) + x.t;
}
__device__ __forceinline__ V3 from_local_vector(const TRS& x, V3 v)
{
    V3 w = trs_unit_scale(x) ? v : v * x.s;
    return rotate_exact(x.qw, x.qv, w, trs_identity_rotation(x), x.absent);
}
__device__ __forceinline__ V3 from_local_normal(const TRS& x, V3 n) { return rotate_exact(x.qw, x.qv, n, trs_identity_rotation(x), x.absent); }
__device__ __forceinline__ V3 to_local_normal(const TRS& x, V3 n) { return rotate_exact(x.qw, -x.qv, n, trs_identity_rotation(x), x.absent); }

// ---------------------------------------------------------------------------
// Slab test (BBox::intersects, RAccel.h:47-59).  t0/t1 are clipped in place.
// ---------------------------------------------------------------------------
// The two halves of the slab test: the part that depends on the node and the ray only ...
__device__ __forceinline__ void box_slabs(float4 q0, float4 q1, V3 o, V3 inv, float& bmin, float& bmax)
{
    V3 a = (mk(q0.x, q0.y, q0.z) - o) * inv;     // vt0 = (m_min - origin) * invDir
    V3 b = (mk(q0.w, q1.x, q1.y) - o) * inv;     // vt1 = (m_max - origin) * invDir
    V3 nr = mk(std_min(a.x, b.x), std_min(a.y, b.y), std_min(a.z, b.z));
    V3 fr = mk(std_max(a.x, b.x), std_max(a.y, b.y), std_max(a.z, b.z));
    bmin = std_max(std_max(nr.x, nr.y), nr.z);   // vtNear.maxComponent()
    bmax = std_min(std_min(fr.x, fr.y), fr.z);   // vtFar.minComponent()
}

// ... and the clipping of the inherited range, which also depends on when the node is popped
__device__ __forceinline__ bool box_clip(float bmin, float bmax, float& t0, float& t1)
{
    t0 = std_max(bmin, t0);
    t1 = std_min(bmax, t1);
    return t0 <= t1;
}

__device__ __forceinline__ bool box_test(float4 q0, float4 q1, V3 o, V3 inv, float& t0, float& t1)
{
    float bmin, bmax;
    box_slabs(q0, q1, o, inv, bmin, bmax);
    return box_clip(bmin, bmax, t0, t1);
}

// The same test for rays whose slab products cannot be NaN (finite non-zero invDir, finite
// origin; node boxes are finite): without NaNs std::min/std::max and the hardware min/max
// return the same number -- they can differ only in the sign of a zero, and every value here
// is consumed by comparisons and further min/max only, never divided by or stored in a hit.
// One FMNMX per min/max instead of a compare and a select (12 of them per node).
__device__ __forceinline__ bool box_test_plain(float4 q0, float4 q1, V3 o, V3 inv, float& t0, float& t1)
{
    V3 a = (mk(q0.x, q0.y, q0.z) - o) * inv;
    V3 b = (mk(q0.w, q1.x, q1.y) - o) * inv;
    float bmin = fmaxf(fmaxf(fminf(a.x, b.x), fminf(a.y, b.y)), fminf(a.z, b.z));
    float bmax = fminf(fminf(fmaxf(a.x, b.x), fmaxf(a.y, b.y)), fmaxf(a.z, b.z));
    t0 = fmaxf(bmin, t0);
    t1 = fminf(bmax, t1);
    return t0 <= t1;
}

// ---------------------------------------------------------------------------
// Primitive tests.  The *_closest forms return the accepted t (and leave tbest
// untouched on a miss); the *_any forms implement the doesIntersect variants,
// which differ in detail from the closest-hit code and are kept separate.
// ---------------------------------------------------------------------------

// Mesh::intersectTri up to the acceptance test (RMesh.h:261-303).  beta/gamma are
// returned for the shading normal.
__device__ __forceinline__ bool tri_closest(V3 o, V3 d, V3 p0, V3 p1, V3 p2, float tbest,
                                            float& t_out, float& beta_out, float& gamma_out)
{
    V3 e1 = p1 - p0;
    V3 e2 = p2 - p0;
    V3 g = cross3(e1, e2);
    float det = -dot3(d, g);
    if (det == 0.0f)
        return false;
    V3 r0 = p0 - o;
    V3 rvc = cross3(d, r0);
    V3 r1 = p1 - o;
    float inv_det = 1.0f / det;
    float gamma = -dot3(r1, rvc) * inv_det;
    if (gamma < 0.0f || gamma > 1.0f)
        return false;
    V3 r2 = p2 - o;
    float beta = dot3(r2, rvc) * inv_det;
    if (beta < 0.0f || beta + gamma > 1.0f)
        return false;
    float t = -dot3(r0, g) * inv_det;
    if (t < RT_RAY_TMIN || t >= tbest)
        return false;
    t_out = t;
    beta_out = beta;
    gamma_out = gamma;
    return true;
}

// Mesh::doesIntersectTri (RMesh.h:338-379) is the same arithmetic against tMax
__device__ __forceinline__ bool tri_any(V3 o, V3 d, V3 p0, V3 p1, V3 p2, float tmax)
{
    float t, b, g;
    return tri_closest(o, d, p0, p1, p2, tmax, t, b, g);
}

// Sphere::intersect (RScene.h:397-455).  lo = local origin minus the centre.
__device__ __forceinline__ bool sphere_closest(V3 lo, V3 ld, float radius, float tbest, float& t_out)
{
    float a = length2(ld);
    float b = 2.0f * dot3(ld, lo);
    float c = length2(lo) - radius * radius;
    float disc = b * b - 4.0f * a * c;
    if (disc < 0.0f)
        return false;
    disc = sqrtf(disc);
    float q = (b < 0.0f) ? (-0.5f * (b - disc)) : (-0.5f * (b + disc));
    float t0 = q / a;
    float t1 = (q != 0.0f) ? (c / q) : tbest;
    if (t0 > t1)
    {
        float tmp = t1;
        t1 = t0;
        t0 = tmp;
    }
    if (t0 >= RT_RAY_TMIN && t0 < tbest) { t_out = t0; return true; }
    if (t1 >= RT_RAY_TMIN && t1 < tbest) { t_out = t1; return true; }
    return false;
}

// Sphere::doesIntersect (RScene.h:468-512): no root swap, different guard order
__device__ __forceinline__ bool sphere_any(V3 lo, V3 ld, float radius, float tmax)
{
    float a = length2(ld);
    float b = 2.0f * dot3(ld, lo);
    float c = length2(lo) - radius * radius;
    float disc = b * b - 4.0f * a * c;
    if (disc < 0.0f)
        return false;
    disc = sqrtf(disc);
    float q = (b < 0.0f) ? (-0.5f * (b - disc)) : (-0.5f * (b + disc));
    float t0 = q / a;
    if (t0 >= RT_RAY_TMIN && t0 < tmax)
        return true;
    float t1 = c / q;
    if (q != 0.0f && t1 < tmax && t1 >= RT_RAY_TMIN)
        return true;
    return false;
}

// Plane::intersect / doesIntersect up to the acceptance test (RScene.h:288-316,
// 333-363): one-sided, so nDotD >= 0 misses.
__device__ __forceinline__ bool plane_test(const DPlane& pl, V3 lo, V3 ld, float tlimit, float& t_out)
{
    V3 n = mk(pl.nx, pl.ny, pl.nz);
    float n_dot_d = dot3(n, ld);
    if (n_dot_d >= 0.0f)
        return false;
    float t = (pl.pos_dot_n - dot3(lo, n)) / n_dot_d;
    if (t >= tlimit || t < RT_RAY_TMIN)
        return false;
    t_out = t;
    return true;
}

// RectangleLight::intersect / doesIntersect up to acceptance (RLight.h:58-100,118-160)
__device__ __forceinline__ bool rect_test(const DRect& rc, V3 lo, V3 ld, float tlimit, float& t_out)
{
    V3 n = mk(rc.nx, rc.ny, rc.nz);
    float n_dot_d = dot3(n, ld);
    if (n_dot_d == 0.0f)
        return false;
    float t = (rc.pos_dot_n - dot3(lo, n)) / n_dot_d;
    if (t >= tlimit || t < RT_RAY_TMIN)
        return false;
    V3 point = lo + t * ld;
    V3 rel = point - mk(rc.px, rc.py, rc.pz);
    float u = dot3(rel, mk(rc.s1x, rc.s1y, rc.s1z));
    float v = dot3(rel, mk(rc.s2x, rc.s2y, rc.s2z));
    if (u < 0.0f || u > rc.len1 || v < 0.0f || v > rc.len2)
        return false;
    t_out = t;
    return true;
}

#endif // RAYITO_B200_RT_DEVICE_CUH
