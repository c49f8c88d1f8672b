// Building blocks of the two-level BVH traversal (record loads, the per-lane ray
// state, hit finalisation) shared by the unified wave kernel (rt_wave.cuh) and the
// split top-level / mesh passes (rt_split.cuh).
//
// Reference semantics reproduced here (Rayito_Stage7_QT):
//   ShapeSet::intersect / doesIntersect            RScene.h:120-184
//   Bvh<T>::intersect / doesIntersect              RAccel.h:389-563
//   Mesh::intersect / doesIntersect (+ face loop)  RMesh.h:62-81, 226-249
// i.e. the ray is warped into set-local space, infinite planes are tested first in
// list order, then either the top-level BVH or (<= 2 finite shapes) the linear
// list; a mesh leaf warps the ray into mesh-local space at the ray's time and
// walks the face BVH.  Node order, (t0,t1) inheritance, the "t0 >= m_t" pre-check,
// the strict "t < m_t" acceptance and the un-box-tested leaves are all kept, so
// the accepted primitive and t are bit-identical to the CPU code.
//
// B200 mapping: nodes fetched as two 128-bit loads, triangles as three, the
// (node, t0, t1) stack in thread-local memory (L1-resident, lane-interleaved).  The
// scene of configs C3/C4 is ~3 MB and stays in L2.
#ifndef RAYITO_B200_RT_TRACE_CUH
#define RAYITO_B200_RT_TRACE_CUH

#include "rt_device.cuh"


// Work counters for the algorithmic-bytes figure
struct WorkCount
{
    uint32_t node_pops, tri_tests, shape_tests, xform_evals;
    uint32_t xform_keyed;    // ... of which on a transform with at least one key (the reference reads key data)
    uint32_t xform_pairs;    // ... of which on a transform with two or more keys (a key pair may be read)
};
#define RT_WORK_COUNTERS 6
#define RT_WORK_ZERO { 0, 0, 0, 0, 0, 0 }

struct LocalRay
{
    V3 o, d, inv;
    uint32_t neg;     // bit a set when inv[a] < 0   (dirSigns, RAccel.h:481-486)
    bool plain;       // the slab arithmetic of this ray cannot produce a NaN (box_test_plain)
};

__device__ __forceinline__ bool finite_nonzero(float v) { float a = fabsf(v); return a > 0.0f && a < __int_as_float(0x7f800000); }
__device__ __forceinline__ bool finite_value(float v) { return fabsf(v) < __int_as_float(0x7f800000); }

__device__ __forceinline__ void local_ray_finish(LocalRay& r)
{
    r.inv = mk(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    r.neg = (r.inv.x < 0.0f ? 1u : 0u) | (r.inv.y < 0.0f ? 2u : 0u) | (r.inv.z < 0.0f ? 4u : 0u);
    // (box - origin) * invDir is NaN only as 0 * inf or inf - inf: a zero or denormal direction
    // component (invDir infinite) or a non-finite origin.  Every other ray is "plain".
    r.plain = finite_nonzero(r.inv.x) && finite_nonzero(r.inv.y) && finite_nonzero(r.inv.z) &&
              finite_value(r.o.x) && finite_value(r.o.y) && finite_value(r.o.z);
}

__device__ __forceinline__ V3 xyz4(float4 v) { return mk(v.x, v.y, v.z); }

struct ClosestHit
{
    float t;            // Intersection::m_t
    int32_t shape;      // -1 = miss
    int32_t tri_rec;    // global triangle record of the winner, -1 for analytic shapes
};

// Load a triangle record
__device__ __forceinline__ void load_tri(const DScene& sc, uint32_t rec, V3& p0, V3& p1, V3& p2,
                                         uint32_t& w0, uint32_t& w1, uint32_t& w2)
{
    float4 a = __ldg(sc.tris + 3 * (size_t)rec + 0);
    float4 b = __ldg(sc.tris + 3 * (size_t)rec + 1);
    float4 c = __ldg(sc.tris + 3 * (size_t)rec + 2);
    p0 = mk(a.x, a.y, a.z);
    p1 = mk(b.x, b.y, b.z);
    p2 = mk(c.x, c.y, c.z);
    w0 = __float_as_uint(a.w);
    w1 = __float_as_uint(b.w);
    w2 = __float_as_uint(c.w);
}

__device__ __forceinline__ DNode load_node(const DNode* nodes, uint32_t i)
{
    DNode n;
    const float4* p = reinterpret_cast<const float4*>(nodes + i);
    n.q0 = __ldg(p);
    n.q1 = __ldg(p + 1);
    return n;
}

__device__ __forceinline__ DShape load_shape(const DScene& sc, uint32_t i)
{
    DShape s;
    const uint4* p = reinterpret_cast<const uint4*>(sc.shapes + i);
    uint4 a = __ldg(p);
    uint4 b = __ldg(p + 1);
    s.type = a.x & 0xffu; s.xkind = (a.x >> 8) & 0x3u; s.trav_cached = (a.x >> 10) & 1u; s.cache_slot = a.x >> 16;
    s.geom = a.y; s.xform = a.z; s.material = a.w;
    s.light = (int32_t)b.x;
    s.tx = __uint_as_float(b.y); s.ty = __uint_as_float(b.z); s.tz = __uint_as_float(b.w);
    return s;
}

// The shape's transform at `time`.  STATIC transforms come straight from the shape
// record (no key loads, no search); the rest go through xform_eval.
// `row`: this sample's row of the per-sample transform cache, or NULL (ray-batch entry points).
__device__ __forceinline__ TRS shape_xform(const DScene& sc, const DShape& sh, float time, const float4* row = nullptr)
{
    if (row != nullptr && sh.cache_slot != 0u)
        return xform_cache_load(row, sh.cache_slot - 1u, sh.xkind, sc.stage6 != 0);
    if (sh.xkind == RT_XF_STATIC)
    {
        TRS r;
        r.absent = sc.stage6 != 0;
        r.t = mk(sh.tx, sh.ty, sh.tz);
        r.s = mk(1.0f, 1.0f, 1.0f);
        r.qw = 1.0f;
        r.qv = mk(0.0f, 0.0f, 0.0f);
        return r;
    }
    return xform_eval(sc, sh.xform, time);
}

// One Ray::transformToLocal of the reference (RRay.h:78-81) on transform `xform`
template <bool COUNT>
__device__ __forceinline__ void count_xform(const DScene& sc, uint32_t xform, WorkCount& wc)
{
    if (COUNT)
    {
        wc.xform_evals++;
        const uint32_t nk = sc.xforms[xform].num_keys;
        wc.xform_keyed += nk >= 1u ? 1u : 0u;
        wc.xform_pairs += nk >= 2u ? 1u : 0u;
    }
}

// The same for the traversal kernels, which read the cache only for shapes marked for it (rt_scene.cuh)
__device__ __forceinline__ TRS shape_xform_trav(const DScene& sc, const DShape& sh, float time, const float4* row)
{
    return shape_xform(sc, sh, time, sh.trav_cached ? row : nullptr);
}

// Shading inputs of the winning hit: Intersection::m_normal and m_colorModifier
// as the reference leaves them after ShapeSet::intersect returns.  Recomputed
// once from the hit identity instead of at every tentative acceptance; the
// arithmetic per shape is the reference's (RScene.h:321-328, 457-463;
// RLight.h:107-113; RMesh.h:305-333, 70-71; RScene.h:152-153).
__device__ __forceinline__ void hit_shading_inputs(const DScene& sc, const LocalRay& r0, float time,
                                                   const ClosestHit& hit, V3& normal, float& color_mod,
                                                   const float4* row = nullptr, int known_type = -1)
{
    normal = mk(0.0f, 0.0f, 0.0f);
    color_mod = 1.0f;
    if (hit.shape < 0)
        return;
    DShape sh = load_shape(sc, (uint32_t)hit.shape);
    if (known_type >= 0)
        sh.type = (uint32_t)known_type;     // a compile-time constant at the call site: the other kinds compile away
    TRS trs = shape_xform(sc, sh, time, row);
    V3 lo = to_local_point(trs, r0.o);
    V3 ld = to_local_vector(trs, r0.d);
    V3 n;
    if (sh.type == RT_SHAPE_PLANE)
    {
        DPlane pl = sc.planes[sh.geom];
        n = from_local_normal(trs, mk(pl.nx, pl.ny, pl.nz));
        if (pl.bullseye)
        {
            V3 rel = (lo + hit.t * ld) - mk(pl.px, pl.py, pl.pz);
            if (fmodf(length3(rel) * 0.25f, 1.0f) > 0.5f)
                color_mod = 0.2f;       // Color(1,1,1) *= 0.2f
        }
    }
    else if (sh.type == RT_SHAPE_SPHERE)
    {
        DSphere s = sc.spheres[sh.geom];
        V3 c = lo - mk(s.px, s.py, s.pz);
        V3 local_norm = c + hit.t * ld;
        n = normalized3(from_local_normal(trs, local_norm));
    }
    else if (sh.type == RT_SHAPE_RECT)
    {
        DRect rc = sc.rects[sh.geom];
        n = from_local_normal(trs, mk(rc.nx, rc.ny, rc.nz));
        if (dot3(n, r0.d) > 0.0f)
            n = n * -1.0f;
    }
    else
    {
        V3 p0, p1, p2;
        uint32_t w0, w1, w2;
        load_tri(sc, (uint32_t)hit.tri_rec, p0, p1, p2, w0, w1, w2);
        V3 e1 = p1 - p0;
        V3 e2 = p2 - p0;
        V3 g = cross3(e1, e2);
        V3 sn;
        if (w2 != 0)
        {
            // barycentrics exactly as intersectTri computes them
            float det = -dot3(ld, g);
            V3 rr0 = p0 - lo;
            V3 rvc = cross3(ld, rr0);
            V3 rr1 = p1 - lo;
            float inv_det = 1.0f / det;
            float gamma = -dot3(rr1, rvc) * inv_det;
            V3 rr2 = p2 - lo;
            float beta = dot3(rr2, rvc) * inv_det;
            float alpha = 1.0f - beta - gamma;
            uint4 ni = __ldg(sc.tri_normals + hit.tri_rec);
            const float* N = sc.normals;
            V3 n0 = mk(N[3 * (size_t)ni.x], N[3 * (size_t)ni.x + 1], N[3 * (size_t)ni.x + 2]);
            V3 n1 = mk(N[3 * (size_t)ni.y], N[3 * (size_t)ni.y + 1], N[3 * (size_t)ni.y + 2]);
            V3 n2 = mk(N[3 * (size_t)ni.z], N[3 * (size_t)ni.z + 1], N[3 * (size_t)ni.z + 2]);
            sn = normalized3((n0 * alpha) + (n1 * beta) + (n2 * gamma));
        }
        else
        {
            sn = sc.stage6 ? g : normalized3(g);      // Stage 6 keeps the raw cross product (S6 RMesh.h:298)
        }
        n = from_local_normal(trs, sn);
    }
    // ShapeSet::intersect: normal out of set-local space (RScene.h:152-153)
    TRS set_trs = xform_eval(sc, sc.set_xform, time);
    normal = from_local_normal(set_trs, n);
}

#endif // RAYITO_B200_RT_TRACE_CUH
