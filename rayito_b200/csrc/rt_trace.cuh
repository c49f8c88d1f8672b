// Two-level BVH traversal: closest hit (ShapeSet::intersect) and any hit
// (ShapeSet::doesIntersect) for one ray, as device functions shared by the
// ray-batch kernels and the wavefront renderer.
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
// B200 mapping: one ray per thread, ONE traversal loop for both BVH levels (lanes
// in the top-level tree and lanes inside a mesh share the slab-test instructions
// instead of serialising two nested loops), nodes fetched as two 128-bit loads,
// triangles as three, the (node, t0, t1) stack in thread-local memory (L1-resident,
// lane-interleaved).  The scene of configs C3/C4 is ~4 MB and stays in L2.
#ifndef RAYITO_B200_RT_TRACE_CUH
#define RAYITO_B200_RT_TRACE_CUH

#include "rt_device.cuh"

#define RT_TOKEN_SHAPE 0x80000000u   /* stack entry that names a shape directly (linear-list mode) */

// Work counters for the algorithmic-bytes figure
struct WorkCount
{
    uint32_t node_pops, tri_tests, shape_tests, xform_evals;
};

struct LocalRay
{
    V3 o, d, inv;
    uint32_t neg;     // bit a set when inv[a] < 0   (dirSigns, RAccel.h:481-486)
};

__device__ __forceinline__ void local_ray_finish(LocalRay& r)
{
    r.inv = mk(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    r.neg = (r.inv.x < 0.0f ? 1u : 0u) | (r.inv.y < 0.0f ? 2u : 0u) | (r.inv.z < 0.0f ? 4u : 0u);
}

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
    s.type = a.x; s.geom = a.y; s.xform = a.z; s.material = a.w;
    s.light = (int32_t)b.x; s.pad0 = s.pad1 = s.pad2 = 0;
    return s;
}

// ---------------------------------------------------------------------------
// Closest hit
// ---------------------------------------------------------------------------
template <int CAP, bool COUNT>
__device__ __forceinline__ ClosestHit trace_closest(const DScene& sc, V3 origin, V3 direction,
                                                    float tmax, float time, LocalRay& set_ray, WorkCount& wc)
{
    uint32_t stk_node[CAP];
    float stk_t0[CAP];
    float stk_t1[CAP];

    // ShapeSet::intersect: ray into set-local space (RScene.h:123-124)
    TRS set_trs = xform_eval(sc, sc.set_xform, time);
    if (COUNT) wc.xform_evals++;
    LocalRay r0;
    r0.o = to_local_point(set_trs, origin);
    r0.d = to_local_vector(set_trs, direction);
    local_ray_finish(r0);
    set_ray = r0;

    ClosestHit hit;
    hit.t = tmax;          // Intersection(ray): m_t = ray.m_tMax
    hit.shape = -1;
    hit.tri_rec = -1;

    // Infinite shapes, list order (RScene.h:126-133)
    for (uint32_t i = 0; i < sc.num_infinite; ++i)
    {
        uint32_t sid = sc.num_finite + i;
        DShape sh = load_shape(sc, sid);
        TRS trs = xform_eval(sc, sh.xform, time);
        V3 lo = to_local_point(trs, r0.o);
        V3 ld = to_local_vector(trs, r0.d);
        if (COUNT) { wc.xform_evals++; wc.shape_tests++; }
        float t;
        if (plane_test(sc.planes[sh.geom], lo, ld, hit.t, t))
        {
            hit.t = t;
            hit.shape = (int32_t)sid;
            hit.tri_rec = -1;
        }
    }

    int sp = 0;
    if (sc.num_top_nodes > 0)
    {
        // Bvh<ShapeSet>::intersect: root with [kRayTMin, m_t] (RAccel.h:495-498)
        stk_node[0] = 0;
        stk_t0[0] = RT_RAY_TMIN;
        stk_t1[0] = hit.t;
        sp = 1;
    }
    else
    {
        // Linear list in insertion order (RScene.h:142-149); pushed in reverse
        for (uint32_t i = sc.num_finite; i > 0; --i)
        {
            stk_node[sp] = RT_TOKEN_SHAPE | (i - 1);
            stk_t0[sp] = 0.0f;
            stk_t1[sp] = 0.0f;
            ++sp;
        }
    }

    LocalRay r1 = r0;             // mesh-local ray while inside a mesh
    int mesh_base = -1;           // stack height below the current mesh's entries
    uint32_t mesh_shape = 0;
    const DNode* mesh_nodes = sc.mesh_nodes;

    while (sp > 0)
    {
        if (sp == mesh_base)
            mesh_base = -1;       // the mesh's BVH is drained: Mesh::intersect returns
        const bool in_mesh = mesh_base >= 0;

        --sp;
        uint32_t node_id = stk_node[sp];
        float t0 = stk_t0[sp];
        float t1 = stk_t1[sp];

        uint32_t shape_id = 0xffffffffu;
        if (!in_mesh && (node_id & RT_TOKEN_SHAPE))
        {
            shape_id = node_id & ~RT_TOKEN_SHAPE;
        }
        else
        {
            DNode nd = load_node(in_mesh ? mesh_nodes : sc.top_nodes, node_id);
            if (COUNT) wc.node_pops++;
            uint32_t flags = __float_as_uint(nd.q1.w);
            uint32_t word = __float_as_uint(nd.q1.z);
            if (flags & RT_NODE_LEAF)
            {
                if (in_mesh)
                {
                    // Mesh::intersect(isect, face): every fan triangle is tested,
                    // later ones against the already shortened m_t (RMesh.h:226-238)
                    uint32_t count = flags >> 3;
                    for (uint32_t k = 0; k < count; ++k)
                    {
                        V3 p0, p1, p2;
                        uint32_t w0, w1, w2;
                        load_tri(sc, word + k, p0, p1, p2, w0, w1, w2);
                        if (COUNT) wc.tri_tests++;
                        float t, beta, gamma;
                        if (tri_closest(r1.o, r1.d, p0, p1, p2, hit.t, t, beta, gamma))
                        {
                            hit.t = t;
                            hit.shape = (int32_t)mesh_shape;
                            hit.tri_rec = (int32_t)(word + k);
                        }
                    }
                    continue;
                }
                shape_id = word;   // top-level leaf: m_shapes[prim]->intersect (RScene.h:262)
            }
            else
            {
                // Interior node (RAccel.h:521-560)
                if (t0 >= hit.t)
                    continue;
                if (t1 > hit.t)
                    t1 = hit.t;
                const LocalRay& r = in_mesh ? r1 : r0;
                if (!box_test(nd.q0, nd.q1, r.o, r.inv, t0, t1))
                    continue;
                uint32_t axis = flags & RT_NODE_AXIS;
                bool neg = (r.neg >> axis) & 1u;
                // left child holds the high side: it is the near one for negative directions
                uint32_t near_id = neg ? word : word + 1;
                uint32_t far_id = neg ? word + 1 : word;
                stk_node[sp] = far_id;  stk_t0[sp] = t0; stk_t1[sp] = t1;
                ++sp;
                stk_node[sp] = near_id; stk_t0[sp] = t0; stk_t1[sp] = t1;
                ++sp;
                continue;
            }
        }

        // A finite shape of the set
        DShape sh = load_shape(sc, shape_id);
        TRS trs = xform_eval(sc, sh.xform, time);
        if (COUNT) wc.xform_evals++;
        V3 lo = to_local_point(trs, r0.o);
        V3 ld = to_local_vector(trs, r0.d);
        if (sh.type == RT_SHAPE_MESH)
        {
            // Mesh::intersect (RMesh.h:62-74): local ray, face BVH with [kRayTMin, m_t]
            DMesh m = sc.meshes[sh.geom];
            if (m.num_nodes > 0)
            {
                r1.o = lo;
                r1.d = ld;
                local_ray_finish(r1);
                mesh_nodes = sc.mesh_nodes + m.first_node;
                mesh_shape = shape_id;
                mesh_base = sp;
                stk_node[sp] = 0;
                stk_t0[sp] = RT_RAY_TMIN;
                stk_t1[sp] = hit.t;
                ++sp;
            }
        }
        else if (sh.type == RT_SHAPE_SPHERE)
        {
            DSphere s = sc.spheres[sh.geom];
            if (COUNT) wc.shape_tests++;
            float t;
            if (sphere_closest(lo - mk(s.px, s.py, s.pz), ld, s.radius, hit.t, t))
            {
                hit.t = t;
                hit.shape = (int32_t)shape_id;
                hit.tri_rec = -1;
            }
        }
        else if (sh.type == RT_SHAPE_RECT)
        {
            if (COUNT) wc.shape_tests++;
            float t;
            if (rect_test(sc.rects[sh.geom], lo, ld, hit.t, t))
            {
                hit.t = t;
                hit.shape = (int32_t)shape_id;
                hit.tri_rec = -1;
            }
        }
    }
    return hit;
}

// Shading inputs of the winning hit: Intersection::m_normal and m_colorModifier
// as the reference leaves them after ShapeSet::intersect returns.  Recomputed
// once from the hit identity instead of at every tentative acceptance; the
// arithmetic per shape is the reference's (RScene.h:321-328, 457-463;
// RLight.h:107-113; RMesh.h:305-333, 70-71; RScene.h:152-153).
__device__ __forceinline__ void hit_shading_inputs(const DScene& sc, const LocalRay& r0, float time,
                                                   const ClosestHit& hit, V3& normal, float& color_mod)
{
    normal = mk(0.0f, 0.0f, 0.0f);
    color_mod = 1.0f;
    if (hit.shape < 0)
        return;
    DShape sh = load_shape(sc, (uint32_t)hit.shape);
    TRS trs = xform_eval(sc, sh.xform, time);
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
            sn = normalized3(g);
        }
        n = from_local_normal(trs, sn);
    }
    // ShapeSet::intersect: normal out of set-local space (RScene.h:152-153)
    TRS set_trs = xform_eval(sc, sc.set_xform, time);
    normal = from_local_normal(set_trs, n);
}

// ---------------------------------------------------------------------------
// Any hit
// ---------------------------------------------------------------------------
template <int CAP, bool COUNT>
__device__ __forceinline__ bool trace_any(const DScene& sc, V3 origin, V3 direction,
                                          float tmax, float time, WorkCount& wc)
{
    uint32_t stk_node[CAP];
    float stk_t0[CAP];
    float stk_t1[CAP];

    TRS set_trs = xform_eval(sc, sc.set_xform, time);
    if (COUNT) wc.xform_evals++;
    LocalRay r0;
    r0.o = to_local_point(set_trs, origin);
    r0.d = to_local_vector(set_trs, direction);
    local_ray_finish(r0);

    for (uint32_t i = 0; i < sc.num_infinite; ++i)
    {
        DShape sh = load_shape(sc, sc.num_finite + i);
        TRS trs = xform_eval(sc, sh.xform, time);
        V3 lo = to_local_point(trs, r0.o);
        V3 ld = to_local_vector(trs, r0.d);
        if (COUNT) { wc.xform_evals++; wc.shape_tests++; }
        float t;
        if (plane_test(sc.planes[sh.geom], lo, ld, tmax, t))
            return true;
    }

    int sp = 0;
    if (sc.num_top_nodes > 0)
    {
        stk_node[0] = 0;
        stk_t0[0] = RT_RAY_TMIN;
        stk_t1[0] = tmax;
        sp = 1;
    }
    else
    {
        for (uint32_t i = sc.num_finite; i > 0; --i)
        {
            stk_node[sp] = RT_TOKEN_SHAPE | (i - 1);
            stk_t0[sp] = 0.0f;
            stk_t1[sp] = 0.0f;
            ++sp;
        }
    }

    LocalRay r1 = r0;
    int mesh_base = -1;
    const DNode* mesh_nodes = sc.mesh_nodes;

    while (sp > 0)
    {
        if (sp == mesh_base)
            mesh_base = -1;
        const bool in_mesh = mesh_base >= 0;

        --sp;
        uint32_t node_id = stk_node[sp];
        float t0 = stk_t0[sp];
        float t1 = stk_t1[sp];

        uint32_t shape_id = 0xffffffffu;
        if (!in_mesh && (node_id & RT_TOKEN_SHAPE))
        {
            shape_id = node_id & ~RT_TOKEN_SHAPE;
        }
        else
        {
            DNode nd = load_node(in_mesh ? mesh_nodes : sc.top_nodes, node_id);
            if (COUNT) wc.node_pops++;
            uint32_t flags = __float_as_uint(nd.q1.w);
            uint32_t word = __float_as_uint(nd.q1.z);
            if (flags & RT_NODE_LEAF)
            {
                if (in_mesh)
                {
                    uint32_t count = flags >> 3;
                    for (uint32_t k = 0; k < count; ++k)
                    {
                        V3 p0, p1, p2;
                        uint32_t w0, w1, w2;
                        load_tri(sc, word + k, p0, p1, p2, w0, w1, w2);
                        if (COUNT) wc.tri_tests++;
                        if (tri_any(r1.o, r1.d, p0, p1, p2, tmax))
                            return true;
                    }
                    continue;
                }
                shape_id = word;
            }
            else
            {
                // No distance culling for shadow rays (RAccel.h:431-440)
                const LocalRay& r = in_mesh ? r1 : r0;
                if (!box_test(nd.q0, nd.q1, r.o, r.inv, t0, t1))
                    continue;
                uint32_t axis = flags & RT_NODE_AXIS;
                bool neg = (r.neg >> axis) & 1u;
                uint32_t near_id = neg ? word : word + 1;
                uint32_t far_id = neg ? word + 1 : word;
                stk_node[sp] = far_id;  stk_t0[sp] = t0; stk_t1[sp] = t1;
                ++sp;
                stk_node[sp] = near_id; stk_t0[sp] = t0; stk_t1[sp] = t1;
                ++sp;
                continue;
            }
        }

        DShape sh = load_shape(sc, shape_id);
        TRS trs = xform_eval(sc, sh.xform, time);
        if (COUNT) wc.xform_evals++;
        V3 lo = to_local_point(trs, r0.o);
        V3 ld = to_local_vector(trs, r0.d);
        if (sh.type == RT_SHAPE_MESH)
        {
            DMesh m = sc.meshes[sh.geom];
            if (m.num_nodes > 0)
            {
                r1.o = lo;
                r1.d = ld;
                local_ray_finish(r1);
                mesh_nodes = sc.mesh_nodes + m.first_node;
                mesh_base = sp;
                stk_node[sp] = 0;
                stk_t0[sp] = RT_RAY_TMIN;
                stk_t1[sp] = tmax;
                ++sp;
            }
        }
        else if (sh.type == RT_SHAPE_SPHERE)
        {
            DSphere s = sc.spheres[sh.geom];
            if (COUNT) wc.shape_tests++;
            if (sphere_any(lo - mk(s.px, s.py, s.pz), ld, s.radius, tmax))
                return true;
        }
        else if (sh.type == RT_SHAPE_RECT)
        {
            if (COUNT) wc.shape_tests++;
            float t;
            if (rect_test(sc.rects[sh.geom], lo, ld, tmax, t))
                return true;
        }
    }
    return false;
}

#endif // RAYITO_B200_RT_TRACE_CUH
