// Warp-persistent two-level BVH traversal ("wave" traversal).
//
// Same per-ray semantics as the reference (see rt_trace.cuh for the citations:
// ShapeSet::intersect/doesIntersect RScene.h:120-184, Bvh<T>::intersect/doesIntersect
// RAccel.h:389-563, Mesh RMesh.h:62-81,226-249): every lane pops ITS stack in
// exactly the reference's order, so accepted primitives and t stay bit-identical.
// What changes is how the 32 lanes of a warp are scheduled, because the first
// profile (profiles/r01_*) showed 6 of 32 lanes active on secondary rays:
//
//   * persistent warps pull rays from the stage's queue through one atomic cursor
//     and REFILL lanes whose ray finished (warp-coherent ray compaction), instead
//     of idling until the slowest ray of the warp is done;
//   * a lane that pops a leaf PARKS there; the warp first lets every lane run down
//     interior nodes (the cheap slab test, executed by all lanes together), then
//     services all parked triangle leaves together and all parked shape leaves
//     together (keyed-transform warp + sphere / rectangle test / mesh entry: the
//     expensive, formerly most divergent code).  Parking never reorders a lane's
//     own pops, so culling decisions see exactly the reference's m_t.
//   * hit finalisation (normals etc.) is NOT done here; callers run it in a fully
//     converged kernel from the (t, shape, triangle) record.
#ifndef RAYITO_B200_RT_WAVE_CUH
#define RAYITO_B200_RT_WAVE_CUH

#include "rt_trace.cuh"

#ifndef RT_REFILL_MIN
#define RT_REFILL_MIN 12         /* refill when at least this many lanes are idle */
#endif
#ifndef RT_ADVANCE_STEPS
#define RT_ADVANCE_STEPS 4       /* node pops per lane and scheduling round */
#endif
#ifndef RT_SERVICE_MIN_TRI
#define RT_SERVICE_MIN_TRI 4     /* lanes parked at triangle leaves before they are serviced */
#endif
#ifndef RT_SERVICE_MIN_SHAPE
#define RT_SERVICE_MIN_SHAPE 6   /* lanes parked at shape leaves before they are serviced */
#endif

enum { PARK_NONE = 0, PARK_TRI = 1, PARK_SHAPE = 2 };

struct WaveResult
{
    float t;
    int32_t shape;
    int32_t tri_rec;
    bool any_hit;
};

// IO policy contract:
//   bool  load(uint32_t j, V3& o, V3& d, float& tmax, float& time, uint32_t& tag)
//   void  store(uint32_t tag, const WaveResult& r)
//   const float4* xf_row(uint32_t tag)      the ray's row of the per-sample transform cache, or NULL
template <int CAP, bool ANY, bool COUNT, class IO>
__device__ __forceinline__ void trace_wave(const DScene& sc, const IO& io, uint32_t n, uint32_t* cursor, WorkCount& wc)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1;

    uint32_t stk_node[CAP];
    float stk_t0[CAP];
    float stk_t1[CAP];

    bool active = false;
    bool exhausted = false;            // warp-uniform: the queue has no more rays
    uint32_t tag = 0;
    float time = 0.0f, tmax = 0.0f;
    LocalRay r0, r1;
    r0.o = r0.d = r0.inv = mk(0.0f, 0.0f, 0.0f);
    r0.neg = 0;
    r1 = r0;
    WaveResult res;
    res.t = 0.0f; res.shape = -1; res.tri_rec = -1; res.any_hit = false;
    int sp = 0;
    int mesh_base = -1;
    uint32_t mesh_shape = 0;
    const DNode* mesh_nodes = sc.mesh_nodes;
    int parked = PARK_NONE;
    uint32_t park_word = 0, park_count = 0;

    for (;;)
    {
        // ------------------------------------------------------------ refill
        uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= RT_REFILL_MIN))
        {
            uint32_t need = __popc(idle);
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(cursor, need);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + need >= n)
                exhausted = true;
            if (!active)
            {
                uint32_t j = base + __popc(idle & lt_mask);
                V3 o, d;
                if (j < n && io.load(j, o, d, tmax, time, tag))
                {
                    active = true;
                    parked = PARK_NONE;
                    mesh_base = -1;
                    sp = 0;
                    res.t = tmax;
                    res.shape = -1;
                    res.tri_rec = -1;
                    res.any_hit = false;
                    // ray into set-local space (RScene.h:123-124 / :161)
                    TRS set_trs = xform_eval(sc, sc.set_xform, time);
                    count_xform<COUNT>(sc, sc.set_xform, wc);
                    r0.o = to_local_point(set_trs, o);
                    r0.d = to_local_vector(set_trs, d);
                    local_ray_finish(r0);
                    // infinite shapes first, in list order (RScene.h:126-133 / :162-169)
                    for (uint32_t k = 0; k < sc.num_infinite; ++k)
                    {
                        uint32_t sid = sc.num_finite + k;
                        DShape sh = load_shape(sc, sid);
                        TRS trs = shape_xform_trav(sc, sh, time, io.xf_row(tag));
                        V3 lo = to_local_point(trs, r0.o);
                        V3 ld = to_local_vector(trs, r0.d);
                        count_xform<COUNT>(sc, sh.xform, wc);
                        if (COUNT) wc.shape_tests++;
                        float t;
                        if (plane_test(sc.planes[sh.geom], lo, ld, res.t, t))
                        {
                            if (ANY) { res.any_hit = true; break; }
                            res.t = t;
                            res.shape = (int32_t)sid;
                        }
                    }
                    if (!(ANY && res.any_hit))
                    {
                        if (sc.num_top_nodes > 0)
                        {
                            stk_node[0] = 0;
                            stk_t0[0] = RT_RAY_TMIN;
                            stk_t1[0] = res.t;         // m_t (closest) or ray.m_tMax (any): both are res.t here
                            sp = 1;
                        }
                        else
                        {
                            for (uint32_t k = sc.num_finite; k > 0; --k)
                            {
                                stk_node[sp] = RT_TOKEN_SHAPE | (k - 1);
                                stk_t0[sp] = 0.0f;
                                stk_t1[sp] = 0.0f;
                                ++sp;
                            }
                        }
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0)
        {
            if (exhausted)
                break;
            continue;
        }

        // ----------------------------------------------------------- advance
        // Up to RT_ADVANCE_STEPS pops per lane and round: interior nodes are slab-
        // tested and their children pushed; reaching a leaf parks the lane.  The
        // bound keeps lanes parked at cheap top-level leaves from waiting for a lane
        // that is descending a deep mesh tree.
        #pragma unroll 1
        for (int it = 0; it < RT_ADVANCE_STEPS && active && parked == PARK_NONE && sp > 0; ++it)
        {
            if (sp == mesh_base)
                mesh_base = -1;                 // the mesh's BVH is drained: Mesh::intersect returns
            const bool in_mesh = mesh_base >= 0;
            --sp;
            uint32_t node_id = stk_node[sp];
            if (!in_mesh && (node_id & RT_TOKEN_SHAPE))
            {
                parked = PARK_SHAPE;
                park_word = node_id & ~RT_TOKEN_SHAPE;
                break;
            }
            DNode nd = load_node(in_mesh ? mesh_nodes : sc.top_nodes, node_id);
            if (COUNT) wc.node_pops++;
            uint32_t flags = __float_as_uint(nd.q1.w);
            uint32_t word = __float_as_uint(nd.q1.z);
            if (flags & RT_NODE_LEAF)
            {
                parked = in_mesh ? PARK_TRI : PARK_SHAPE;
                park_word = word;
                park_count = flags >> 3;
                break;
            }
            float t0 = stk_t0[sp];
            float t1 = stk_t1[sp];
            if (!ANY)
            {
                // closest hit culls against the current m_t (RAccel.h:523-530)
                if (t0 >= res.t)
                    continue;
                if (t1 > res.t)
                    t1 = res.t;
            }
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
        }

        // Service a leaf kind only when enough lanes wait for it (so the expensive
        // code runs well filled), or when no lane can advance any more, or when the
        // queue is drained and waiting would only idle the warp.
        const uint32_t m_tri = __ballot_sync(0xffffffffu, parked == PARK_TRI);
        const uint32_t m_shape = __ballot_sync(0xffffffffu, parked == PARK_SHAPE);
        const uint32_t m_adv = __ballot_sync(0xffffffffu, active && parked == PARK_NONE && sp > 0);
        const bool flush_all = m_adv == 0 || exhausted;
        const bool do_tri = m_tri != 0 && (flush_all || __popc(m_tri) >= RT_SERVICE_MIN_TRI);
        const bool do_shape = m_shape != 0 && (flush_all || __popc(m_shape) >= RT_SERVICE_MIN_SHAPE);

        // ----------------------------------------------------------- service
        if (do_tri && parked == PARK_TRI)
        {
            // Mesh::intersect(isect, face) / doesIntersect(ray, face) (RMesh.h:226-249)
            for (uint32_t k = 0; k < park_count; ++k)
            {
                V3 p0, p1, p2;
                uint32_t w0, w1, w2;
                load_tri(sc, park_word + k, p0, p1, p2, w0, w1, w2);
                if (COUNT) wc.tri_tests++;
                float t, beta, gamma;
                if (tri_closest(r1.o, r1.d, p0, p1, p2, ANY ? tmax : res.t, t, beta, gamma))
                {
                    if (ANY) { res.any_hit = true; sp = 0; break; }
                    res.t = t;
                    res.shape = (int32_t)mesh_shape;
                    res.tri_rec = (int32_t)(park_word + k);
                    if (sc.stage6) break;        // S6 RMesh.h:204-209: the first fan triangle to hit claims the face
                }
            }
            parked = PARK_NONE;
        }
        if (do_shape && parked == PARK_SHAPE)
        {
            DShape sh = load_shape(sc, park_word);
            TRS trs = shape_xform_trav(sc, sh, time, io.xf_row(tag));
            count_xform<COUNT>(sc, sh.xform, wc);
            V3 lo = to_local_point(trs, r0.o);
            V3 ld = to_local_vector(trs, r0.d);
            if (sh.type == RT_SHAPE_MESH)
            {
                // Mesh::intersect / doesIntersect (RMesh.h:62-81)
                DMesh m = sc.meshes[sh.geom];
                if (m.num_nodes > 0)
                {
                    r1.o = lo;
                    r1.d = ld;
                    local_ray_finish(r1);
                    mesh_nodes = sc.mesh_nodes + m.first_node;
                    mesh_shape = park_word;
                    mesh_base = sp;
                    stk_node[sp] = 0;
                    stk_t0[sp] = RT_RAY_TMIN;
                    stk_t1[sp] = ANY ? tmax : res.t;
                    ++sp;
                }
            }
            else if (sh.type == RT_SHAPE_SPHERE)
            {
                DSphere s = sc.spheres[sh.geom];
                if (COUNT) wc.shape_tests++;
                V3 c = lo - mk(s.px, s.py, s.pz);
                if (ANY)
                {
                    if (sphere_any(c, ld, s.radius, tmax)) { res.any_hit = true; sp = 0; }
                }
                else
                {
                    float t;
                    if (sphere_closest(c, ld, s.radius, res.t, t))
                    {
                        res.t = t;
                        res.shape = (int32_t)park_word;
                        res.tri_rec = -1;
                    }
                }
            }
            else if (sh.type == RT_SHAPE_RECT)
            {
                if (COUNT) wc.shape_tests++;
                float t;
                if (rect_test(sc.rects[sh.geom], lo, ld, ANY ? tmax : res.t, t))
                {
                    if (ANY) { res.any_hit = true; sp = 0; }
                    else
                    {
                        res.t = t;
                        res.shape = (int32_t)park_word;
                        res.tri_rec = -1;
                    }
                }
            }
            parked = PARK_NONE;
        }

        // ------------------------------------------------------------- retire
        if (active && parked == PARK_NONE && sp == 0)
        {
            io.store(tag, res);
            active = false;
        }
    }
}

#endif // RAYITO_B200_RT_WAVE_CUH
