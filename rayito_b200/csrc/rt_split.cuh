// Split two-level traversal: top-level passes and mesh passes as separate
// warp-persistent kernels.
//
// The unified wave kernel (rt_wave.cuh) still mixed two very different kinds of
// work in one warp: short top-level walks (5 pops, 2-3 analytic shapes with their
// keyed-transform warps) and long face-BVH walks inside a mesh (tens of pops and
// triangle tests).  ncu showed 6-10 of 32 lanes active: the few lanes deep inside a
// mesh dragged every scheduling round while the short rays came and went.
//
// Here a ray that reaches a mesh leaf of the top-level tree is SUSPENDED: its set-
// local ray, current m_t / winner and its (tiny) top-level stack are written to
// per-slot arrays and its slot is queued for a mesh pass.  A mesh pass warps the
// ray into mesh-local space at the ray's time, walks the face BVH with
// [kRayTMin, m_t] exactly as Mesh::intersect does (RMesh.h:62-81), updates m_t and
// queues the slot for a resume pass, which reloads the stack and continues the
// top-level walk where it stopped.  A mesh is one leaf of the top-level tree, so a
// ray suspends at most once per mesh: 1 + 2*M passes cover every ray.
//
// Per-ray pop order, (t0,t1) inheritance, culling against the current m_t and
// acceptance tests are untouched -- a suspended ray sees exactly the state it would
// have had -- so results stay bit-identical to the unified kernel and to the
// reference (the tests compare work counters and images between the two modes).
#ifndef RAYITO_B200_RT_SPLIT_CUH
#define RAYITO_B200_RT_SPLIT_CUH

#include <cuda_pipeline.h>
#include "rt_wave.cuh"

#ifndef RT_TOP_PLAIN_SLABS
#define RT_TOP_PLAIN_SLABS 1   /* hardware min/max in the tabulated top-level walk's slab tests (box_test_plain, see trace_mesh): C4 +1.2 % frame, +1.9 % traversal */
#endif
#ifndef RT_TOP_REFILL_MIN
#define RT_TOP_REFILL_MIN 12
#endif
#ifndef RT_TOP_ADVANCE_STEPS
#define RT_TOP_ADVANCE_STEPS 4
#endif
#ifndef RT_TOP_SERVICE_MIN
#define RT_TOP_SERVICE_MIN 6
#endif
#ifndef RT_MESH_REFILL_MIN
#define RT_MESH_REFILL_MIN 12
#endif
#ifndef RT_MESH_ADVANCE_STEPS
#define RT_MESH_ADVANCE_STEPS 4
#endif
#ifndef RT_MESH_SERVICE_MIN
#define RT_MESH_SERVICE_MIN 4
#endif

#define RT_SPLIT_TOPCAP 8        /* top-level stack entries a suspended ray can carry */
#define RT_SPLIT_DIRECT 0x80000000u
#ifndef RT_MESH_DIRECT_FINISH
#define RT_MESH_DIRECT_FINISH 1     /* 0: every suspended ray goes through a resume pass (A/B runs) */
#endif
#define RT_SPLIT_MAX_MESHES 12   /* more mesh shapes than this: use the unified kernel */

// Per-slot suspended-ray state (slot = path sample index / ray index): one 128-byte record per slot,
// of which a suspension writes and a later pass reads only the 32-byte SECTORS it needs, each sector
// always whole.  (Separate 16-byte arrays made every store a partial-sector write that L2 has to
// fill from HBM first -- half the DRAM reads of the top-level passes on the 10 M-triangle scene --
// and a record written whole regardless of its content costs the small scenes more than it saves.)
// The set-local ray itself is NOT saved: a resume pass recomputes it from the stage's own ray record
// (same inputs, same bits).
//   sector 0  rec[0] mesh-local origin xyz (computed by the top pass at mesh entry), m_t (closest hit) or tMax (any hit)
//             rec[1] mesh-local direction xyz, mesh shape to enter | RT_SPLIT_DIRECT when no top-level work is pending
//                    (the mesh pass then finishes the ray itself instead of queueing it for a resume pass)
//   sector 1  rec[2] m_t, winner shape, winner triangle record, (mesh shape to enter | sp << 24)
//             rec[3] top-level stack entry 0: node, t0, t1, -
//   sector 2  rec[4], rec[5] stack entries 1, 2        (written when sp > 1)
//   sector 3  rec[6], rec[7] stack entries 3, 4        (written when sp > 3)
//   stack[(RT_SPLIT_TOPCAP - 5) * slot + k - 5]   stack entries k >= 5
#define RT_SPLIT_REC 8          /* float4 per record */
struct SplitBufs
{
    float4* rec;
    float4* stack;
};
__device__ __forceinline__ float4* split_rec(const SplitBufs& sb, uint32_t slot) { return sb.rec + RT_SPLIT_REC * (size_t)slot; }
__device__ __forceinline__ float4* split_stack_entry(const SplitBufs& sb, uint32_t slot, int k)
{
    return k < 5 ? sb.rec + RT_SPLIT_REC * (size_t)slot + 3 + k : sb.stack + (size_t)slot * (RT_SPLIT_TOPCAP - 5) + (k - 5);
}
// After entries 0 .. sp-1 were written: fill the other half of a half-written sector
__device__ __forceinline__ void split_stack_pad(const SplitBufs& sb, uint32_t slot, int sp)
{
    if (sp == 0) *split_stack_entry(sb, slot, 0) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (sp == 2 || sp == 4) *split_stack_entry(sb, slot, sp) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// One pass: where the work comes from, where suspended / resumed slots go, and
// which counters this kernel zeroes for the passes after it (none of which it uses)
struct SplitPass
{
    const uint32_t* in_queue;    // resume / mesh passes: slots to process
    const uint32_t* in_count;    // number of entries (device memory)
    uint32_t* cursor;            // atomic work cursor of this pass
    uint32_t* out_queue;         // top pass: mesh queue; mesh pass: resume queue
    uint32_t* out_count;
    uint32_t* zero[4];           // counters to reset (may be NULL)
};

__device__ __forceinline__ void split_zero(const SplitPass& ps)
{
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int k = 0; k < 4; ++k)
            if (ps.zero[k]) *ps.zero[k] = 0;
}

// Append to a queue: one atomicAdd per warp.  All 32 lanes must call.
__device__ __forceinline__ void warp_queue_push(uint32_t* queue, uint32_t* counter, bool want, uint32_t value)
{
    uint32_t mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0)
        return;
    uint32_t lane = threadIdx.x & 31;
    uint32_t leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader)
        base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want)
        queue[base + __popc(mask & ((1u << lane) - 1))] = value;
}

// ---------------------------------------------------------------------------
// Top-level pass.  FRESH: rays come from the stage's IO (queue of path slots);
// otherwise they are resumed from their suspended state.
// ---------------------------------------------------------------------------
template <bool ANY, bool COUNT, bool FRESH, class IO>
__device__ __forceinline__ void trace_top(const DScene& sc, const IO& io, const SplitBufs& sb, const SplitPass& ps, WorkCount& wc)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1;
    const uint32_t n = FRESH ? io.count() : *ps.in_count;

    uint32_t stk_node[RT_SPLIT_TOPCAP];
    float stk_t0[RT_SPLIT_TOPCAP];
    float stk_t1[RT_SPLIT_TOPCAP];

    bool active = false;
    bool exhausted = false;
    uint32_t tag = 0;
    float time = 0.0f, tmax = 0.0f;
    LocalRay r0;
    r0.o = r0.d = r0.inv = mk(0.0f, 0.0f, 0.0f);
    r0.neg = 0;
    WaveResult res;
    res.t = 0.0f; res.shape = -1; res.tri_rec = -1; res.any_hit = false;
    int sp = 0;
    bool parked = false;
    uint32_t park_shape = 0;

    for (;;)
    {
        // ------------------------------------------------------------ refill
        uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= RT_TOP_REFILL_MIN))
        {
            uint32_t need = __popc(idle);
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(ps.cursor, need);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + need >= n)
                exhausted = true;
            uint32_t j = base + __popc(idle & lt_mask);
            if (!active && j < n)
            {
                active = true;
                parked = false;
                sp = 0;
                if (FRESH)
                {
                    V3 o, d;
                    io.load(j, o, d, tmax, time, tag);
                    res.t = tmax;
                    res.shape = -1;
                    res.tri_rec = -1;
                    res.any_hit = false;
                    TRS set_trs = xform_eval(sc, sc.set_xform, time);
                    count_xform<COUNT>(sc, sc.set_xform, wc);
                    r0.o = to_local_point(set_trs, o);
                    r0.d = to_local_vector(set_trs, d);
                    local_ray_finish(r0);
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
                            stk_t1[0] = res.t;
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
                else
                {
                    // resume a suspended ray: the mesh pass has updated m_t / the winner; the set-local ray is
                    // recomputed from the stage's ray record exactly as the fresh pass computed it
                    tag = ps.in_queue[j];
                    float4 h = split_rec(sb, tag)[2];
                    {
                        V3 o, d;
                        io.load_tag(tag, o, d, tmax, time);
                        TRS set_trs = xform_eval(sc, sc.set_xform, time);
                        r0.o = to_local_point(set_trs, o);
                        r0.d = to_local_vector(set_trs, d);
                    }
                    local_ray_finish(r0);
                    res.t = h.x;
                    res.shape = __float_as_int(h.y);
                    res.tri_rec = __float_as_int(h.z);
                    res.any_hit = false;
                    sp = (int)(__float_as_uint(h.w) >> 24);
                    #pragma unroll
                    for (int k = 0; k < RT_SPLIT_TOPCAP; ++k)
                    {
                        if (k < sp)
                        {
                            float4 e = *split_stack_entry(sb, tag, k);
                            stk_node[k] = __float_as_uint(e.x);
                            stk_t0[k] = e.y;
                            stk_t1[k] = e.z;
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
        #pragma unroll 1
        for (int it = 0; it < RT_TOP_ADVANCE_STEPS && active && !parked && sp > 0; ++it)
        {
            --sp;
            uint32_t node_id = stk_node[sp];
            if (node_id & RT_TOKEN_SHAPE)
            {
                parked = true;
                park_shape = node_id & ~RT_TOKEN_SHAPE;
                break;
            }
            DNode nd = load_node(sc.top_nodes, node_id);
            if (COUNT) wc.node_pops++;
            uint32_t flags = __float_as_uint(nd.q1.w);
            uint32_t word = __float_as_uint(nd.q1.z);
            if (flags & RT_NODE_LEAF)
            {
                parked = true;
                park_shape = word;
                break;
            }
            float t0 = stk_t0[sp];
            float t1 = stk_t1[sp];
            if (!ANY)
            {
                if (t0 >= res.t)
                    continue;
                if (t1 > res.t)
                    t1 = res.t;
            }
            if (!box_test(nd.q0, nd.q1, r0.o, r0.inv, t0, t1))
                continue;
            uint32_t axis = flags & RT_NODE_AXIS;
            bool neg = (r0.neg >> axis) & 1u;
            uint32_t near_id = neg ? word : word + 1;
            uint32_t far_id = neg ? word + 1 : word;
            stk_node[sp] = far_id;  stk_t0[sp] = t0; stk_t1[sp] = t1;
            ++sp;
            stk_node[sp] = near_id; stk_t0[sp] = t0; stk_t1[sp] = t1;
            ++sp;
        }

        const uint32_t m_shape = __ballot_sync(0xffffffffu, parked);
        const uint32_t m_adv = __ballot_sync(0xffffffffu, active && !parked && sp > 0);
        const bool do_shape = m_shape != 0 && (m_adv == 0 || exhausted || __popc(m_shape) >= RT_TOP_SERVICE_MIN);

        // ----------------------------------------------------------- service
        bool suspend = false;
        if (do_shape && parked)
        {
            DShape sh = load_shape(sc, park_shape);
            if (sh.type == RT_SHAPE_MESH)
            {
                // Mesh::intersect / doesIntersect (RMesh.h:62-81): warp the ray into
                // mesh-local space and pop the face BVH's root here; most rays that
                // reach a mesh leaf miss the mesh's root box and never need a mesh pass.
                DMesh m = sc.meshes[sh.geom];
                if (m.num_nodes > 0)
                {
                    TRS trs = shape_xform_trav(sc, sh, time, io.xf_row(tag));
                    count_xform<COUNT>(sc, sh.xform, wc);
                    LocalRay rm;
                    rm.o = to_local_point(trs, r0.o);
                    rm.d = to_local_vector(trs, r0.d);
                    local_ray_finish(rm);
                    DNode root = load_node(sc.mesh_nodes + m.first_node, 0);
                    if (COUNT) wc.node_pops++;
                    bool enter = true;
                    if (!(__float_as_uint(root.q1.w) & RT_NODE_LEAF))
                    {
                        float t0 = RT_RAY_TMIN, t1 = ANY ? tmax : res.t;
                        if (!ANY && t0 >= res.t)
                            enter = false;
                        else
                            enter = box_test(root.q0, root.q1, rm.o, rm.inv, t0, t1);
                    }
                    if (enter)
                    {
                        suspend = true;
                        float4* rec = split_rec(sb, tag);
                        rec[0] = make_float4(rm.o.x, rm.o.y, rm.o.z, ANY ? tmax : res.t);
                        rec[1] = make_float4(rm.d.x, rm.d.y, rm.d.z, __uint_as_float(park_shape | (sp == 0 && RT_MESH_DIRECT_FINISH ? RT_SPLIT_DIRECT : 0u)));
                    }
                }
            }
            else
            {
                TRS trs = shape_xform_trav(sc, sh, time, io.xf_row(tag));
                count_xform<COUNT>(sc, sh.xform, wc);
                V3 lo = to_local_point(trs, r0.o);
                V3 ld = to_local_vector(trs, r0.d);
                if (sh.type == RT_SHAPE_SPHERE)
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
                            res.shape = (int32_t)park_shape;
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
                            res.shape = (int32_t)park_shape;
                            res.tri_rec = -1;
                        }
                    }
                }
            }
            parked = false;
        }
        if (__ballot_sync(0xffffffffu, suspend))
        {
            if (suspend)
            {
                float4* rec = split_rec(sb, tag);
                rec[2] = make_float4(res.t, __int_as_float(res.shape), __int_as_float(res.tri_rec),
                                     __uint_as_float(park_shape | ((uint32_t)sp << 24)));
                #pragma unroll
                for (int k = 0; k < RT_SPLIT_TOPCAP; ++k)
                    if (k < sp)
                        *split_stack_entry(sb, tag, k) = make_float4(__uint_as_float(stk_node[k]), stk_t0[k], stk_t1[k], 0.0f);
                split_stack_pad(sb, tag, sp);       // sectors go out whole
                active = false;
            }
            warp_queue_push(ps.out_queue, ps.out_count, suspend, tag);
        }

        // ------------------------------------------------------------- retire
        if (active && !parked && sp == 0)
        {
            io.store(tag, res);
            active = false;
        }
    }
}

// ---------------------------------------------------------------------------
// Fresh top-level pass, tabulated walk.
//
// The top level is a handful of shapes (8 in the Stage 7 scene: a 15-node BVH), yet the
// dynamic pass above spends more time in it than the mesh passes spend in a 49 151-node face
// BVH: every lane is at a different node or shape, so a warp executes the union of all
// their code paths with 12-20 lanes active.  The ORDER in which Bvh::intersect pops nodes
// depends only on the direction signs, so here all lanes of a warp that share an octant step
// through the same tabulated pop sequence (DTopStep) together; what differs per lane is only
// whether a pop happens at all (its parent passed its slab test) and with which inherited
// range, kept per lane and per depth.  A node popped by no lane is skipped by the whole warp.
// Every lane performs exactly the pops, checks and shape tests of the reference, in the same
// order, on the same values; rays entering a mesh are suspended with the explicit stack the
// dynamic pass would hold at that point, so the mesh and resume passes are unchanged.
// ---------------------------------------------------------------------------
#ifndef RT_STATIC_PREFETCH
#define RT_STATIC_PREFETCH 1      /* +2.4 % frame, +4.6 % traversal in same-session A/B */
#endif
template <bool ANY, bool COUNT, class IO>
__device__ __forceinline__ void trace_top_static(const DScene& sc, const IO& io, const SplitBufs& sb, const SplitPass& ps,
                                                 WorkCount& wc, float* lane_t0, float* lane_t1, float4* stage_rec)
{
    // lane_t0 / lane_t1: [RT_WALK_MAX_DEPTH + 1][blockDim.x] shared floats
    const uint32_t lane = threadIdx.x & 31, tid = threadIdx.x, stride = blockDim.x;
    const uint32_t n = io.count();
    const uint32_t nsteps = sc.top_walk_steps;

#if RT_STATIC_PREFETCH
    // Three-deep hand-out pipeline: while chunk k is walked, the ray records of chunk k+1 are on
    // their way into shared memory (cp.async), the queue entries of chunk k+2 into registers, and
    // the cursor atomic of chunk k+3 is in flight; a refill then waits for none of them.
    uint32_t base_cur, tag_cur = 0, base_nxt, tag_nxt = 0, base_nn = 0;
    {
        uint32_t b0 = 0;
        if (lane == 0) b0 = atomicAdd(ps.cursor, 32u);
        base_cur = __shfl_sync(0xffffffffu, b0, 0);
        if (base_cur + lane < n)
        {
            tag_cur = io.tag_at(base_cur + lane);
            __pipeline_memcpy_async(stage_rec + 0 * RT_BLOCK + tid, io.rec_a(tag_cur), 16);
            __pipeline_memcpy_async(stage_rec + 1 * RT_BLOCK + tid, io.rec_b(tag_cur), 16);
        }
        __pipeline_commit();
        if (lane == 0) b0 = atomicAdd(ps.cursor, 32u);
        base_nxt = __shfl_sync(0xffffffffu, b0, 0);
        if (base_nxt + lane < n)
            tag_nxt = io.tag_at(base_nxt + lane);
        if (lane == 0) base_nn = atomicAdd(ps.cursor, 32u);
    }
    for (uint32_t chunk = 0;; ++chunk)
    {
        if (base_cur >= n)
            break;
        const uint32_t stage = chunk & 1u;
        if (base_nxt + lane < n)
        {
            __pipeline_memcpy_async(stage_rec + ((stage ^ 1u) * 2 + 0) * RT_BLOCK + tid, io.rec_a(tag_nxt), 16);
            __pipeline_memcpy_async(stage_rec + ((stage ^ 1u) * 2 + 1) * RT_BLOCK + tid, io.rec_b(tag_nxt), 16);
        }
        __pipeline_commit();
        __pipeline_wait_prior(1);           // this chunk's records have landed
        const uint32_t j = base_cur + lane;
        const bool live = j < n;
#else
    for (;;)
    {
        uint32_t base = 0;
        if (lane == 0)
            base = atomicAdd(ps.cursor, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n)
            break;
        const uint32_t j = base + lane;
        const bool live = j < n;
#endif

        uint32_t tag = 0;
        float time = 0.0f, tmax = 0.0f;
        LocalRay r0;
        r0.o = r0.d = r0.inv = mk(0.0f, 0.0f, 0.0f);
        r0.neg = 0;
        r0.plain = false;
        WaveResult res;
        res.t = 0.0f; res.shape = -1; res.tri_rec = -1; res.any_hit = false;
        bool open = false;          // still walking
        bool suspended = false;
        if (live)
        {
            V3 o, d;
#if RT_STATIC_PREFETCH
            tag = tag_cur;
            io.decode(stage_rec[(stage * 2 + 0) * RT_BLOCK + tid], stage_rec[(stage * 2 + 1) * RT_BLOCK + tid], o, d, tmax, time);
#else
            io.load(j, o, d, tmax, time, tag);
#endif
            res.t = tmax;
            TRS set_trs = xform_eval(sc, sc.set_xform, time);
            count_xform<COUNT>(sc, sc.set_xform, wc);
            r0.o = to_local_point(set_trs, o);
            r0.d = to_local_vector(set_trs, d);
            local_ray_finish(r0);
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
            open = !(ANY && res.any_hit);
        }
        lane_t0[tid] = RT_RAY_TMIN;         // depth 0: the root's range (RAccel.h:402-404)
        lane_t1[tid] = res.t;
        uint32_t alive = 1u;                // bit d: the node at depth d on the current path was pushed

#if RT_TOP_PLAIN_SLABS
        // all rays of this warp's group are NaN-free: hardware min/max in the slab tests (box_test_plain)
        const bool warp_plain = __all_sync(0xffffffffu, !open || r0.plain);
#endif
        uint32_t todo = __ballot_sync(0xffffffffu, open);
        while (todo)
        {
            const uint32_t oct = __shfl_sync(0xffffffffu, r0.neg, __ffs(todo) - 1);
            const uint32_t grp = __ballot_sync(0xffffffffu, open && r0.neg == oct);
            todo &= ~grp;
            const bool in_grp = (grp >> lane) & 1u;
            const DTopStep* steps = sc.top_walk + (size_t)oct * nsteps;
            #pragma unroll 1
            for (uint32_t s = 0; s < nsteps; ++s)
            {
                const uint4 hd = __ldg(reinterpret_cast<const uint4*>(steps + s));
                const uint32_t node = hd.x, word = hd.y, flags = hd.z;
                const uint32_t depth = (flags >> 8) & 0xffu;
                const bool here = in_grp && open && ((alive >> depth) & 1u);
                const uint32_t m_here = __ballot_sync(0xffffffffu, here);
                if (!(flags & RT_NODE_LEAF))
                {
                    bool pass = false;
                    if (m_here)
                    {
                        DNode nd = load_node(sc.top_nodes, node);
                        if (here)
                        {
                            if (COUNT) wc.node_pops++;
                            float t0 = lane_t0[depth * stride + tid];
                            float t1 = lane_t1[depth * stride + tid];
                            bool go = true;
                            if (!ANY)
                            {
                                if (t0 >= res.t)
                                    go = false;
                                else if (t1 > res.t)
                                    t1 = res.t;
                            }
#if RT_TOP_PLAIN_SLABS
                            if (go && (warp_plain ? box_test_plain(nd.q0, nd.q1, r0.o, r0.inv, t0, t1)
                                                  : box_test(nd.q0, nd.q1, r0.o, r0.inv, t0, t1)))
#else
                            if (go && box_test(nd.q0, nd.q1, r0.o, r0.inv, t0, t1))
#endif
                            {
                                pass = true;
                                lane_t0[(depth + 1) * stride + tid] = t0;
                                lane_t1[(depth + 1) * stride + tid] = t1;
                            }
                        }
                    }
                    if (in_grp)
                        alive = pass ? (alive | (2u << depth)) : (alive & ~(2u << depth));
                    continue;
                }
                if (m_here == 0)
                    continue;
                if (COUNT && here && node != RT_WALK_TOKEN) wc.node_pops++;

                // ---- a shape leaf, popped by the lanes in m_here (same code as trace_top's service)
                const uint32_t shape_id = word;
                DShape sh = load_shape(sc, shape_id);
                bool suspend = false;
                if (sh.type == RT_SHAPE_MESH)
                {
                    DMesh m = sc.meshes[sh.geom];
                    if (here && m.num_nodes > 0)
                    {
                        TRS trs = shape_xform_trav(sc, sh, time, io.xf_row(tag));
                        count_xform<COUNT>(sc, sh.xform, wc);
                        LocalRay rm;
                        rm.o = to_local_point(trs, r0.o);
                        rm.d = to_local_vector(trs, r0.d);
                        local_ray_finish(rm);
                        DNode root = load_node(sc.mesh_nodes + m.first_node, 0);
                        if (COUNT) wc.node_pops++;
                        bool enter = true;
                        if (!(__float_as_uint(root.q1.w) & RT_NODE_LEAF))
                        {
                            float t0 = RT_RAY_TMIN, t1 = ANY ? tmax : res.t;
                            if (!ANY && t0 >= res.t)
                                enter = false;
                            else
                                enter = box_test(root.q0, root.q1, rm.o, rm.inv, t0, t1);
                        }
                        if (enter)
                        {
                            suspend = true;
                            float4* rec = split_rec(sb, tag);
                            rec[0] = make_float4(rm.o.x, rm.o.y, rm.o.z, ANY ? tmax : res.t);

                            // the explicit stack of the dynamic pass at this point: pending far
                            // children whose parents passed, with the ranges those parents pushed
                            const uint4 p0 = __ldg(reinterpret_cast<const uint4*>(steps + s) + 1);
                            const uint4 p1 = __ldg(reinterpret_cast<const uint4*>(steps + s) + 2);
                            const uint32_t pn[8] = { p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w };
                            const uint32_t npend = (flags >> 16) & 0xffu;
                            uint32_t sp = 0;
                            #pragma unroll
                            for (uint32_t k = 0; k < RT_WALK_MAX_DEPTH; ++k)
                            {
                                if (k < npend)
                                {
                                    uint32_t dk = (hd.w >> (4 * k)) & 0xfu;
                                    if ((alive >> dk) & 1u)
                                    {
                                        *split_stack_entry(sb, tag, (int)sp) = make_float4(__uint_as_float(pn[k]), lane_t0[dk * stride + tid],
                                                                                           lane_t1[dk * stride + tid], 0.0f);
                                        ++sp;
                                    }
                                }
                            }
                            split_stack_pad(sb, tag, (int)sp);      // sectors go out whole
                            rec[2] = make_float4(res.t, __int_as_float(res.shape), __int_as_float(res.tri_rec),
                                                 __uint_as_float(shape_id | (sp << 24)));
                            rec[1] = make_float4(rm.d.x, rm.d.y, rm.d.z, __uint_as_float(shape_id | (sp == 0 && RT_MESH_DIRECT_FINISH ? RT_SPLIT_DIRECT : 0u)));
                            open = false;
                            suspended = true;
                        }
                    }
                }
                else if (here)
                {
                    TRS trs = shape_xform_trav(sc, sh, time, io.xf_row(tag));
                    count_xform<COUNT>(sc, sh.xform, wc);
                    V3 lo = to_local_point(trs, r0.o);
                    V3 ld = to_local_vector(trs, r0.d);
                    if (sh.type == RT_SHAPE_SPHERE)
                    {
                        DSphere sp = sc.spheres[sh.geom];
                        if (COUNT) wc.shape_tests++;
                        V3 c = lo - mk(sp.px, sp.py, sp.pz);
                        if (ANY)
                        {
                            if (sphere_any(c, ld, sp.radius, tmax)) { res.any_hit = true; open = false; }
                        }
                        else
                        {
                            float t;
                            if (sphere_closest(c, ld, sp.radius, res.t, t))
                            {
                                res.t = t;
                                res.shape = (int32_t)shape_id;
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
                            if (ANY) { res.any_hit = true; open = false; }
                            else
                            {
                                res.t = t;
                                res.shape = (int32_t)shape_id;
                                res.tri_rec = -1;
                            }
                        }
                    }
                }
                (void)suspend;
            }
        }
        // rays that entered a mesh: one queue append per chunk (a lane suspends at most once per walk)
        warp_queue_push(ps.out_queue, ps.out_count, suspended, tag);
        if (live && !suspended)
            io.store(tag, res);
#if RT_STATIC_PREFETCH
        base_cur = base_nxt;
        tag_cur = tag_nxt;
        base_nxt = __shfl_sync(0xffffffffu, base_nn, 0);
        tag_nxt = 0;
        if (base_nxt + lane < n)
            tag_nxt = io.tag_at(base_nxt + lane);
        if (lane == 0 && base_nxt < n)
            base_nn = atomicAdd(ps.cursor, 32u);
#endif
    }
#if RT_STATIC_PREFETCH
    __pipeline_wait_prior(0);
#endif
}

// ---------------------------------------------------------------------------
// Mesh pass: Mesh::intersect / doesIntersect for suspended rays
// ---------------------------------------------------------------------------
#ifndef RT_MESH_PREFETCH
#define RT_MESH_PREFETCH 0     /* 1: L2 prefetch of a leaf's triangle records when the lane parks on it */
#endif
#ifndef RT_MESH_PLAIN_SLABS
#define RT_MESH_PLAIN_SLABS 1  /* hardware min/max in the slab test of NaN-free rays (box_test_plain): C5 +1.0 %, C4 +0.5 % */
#endif
#ifndef RT_MESH_TOPCACHE
#define RT_MESH_TOPCACHE 1     /* newest far entry cached in registers: +0.6 % C4, +2.6 % C5 */
#endif
// The first RT_MESH_SMEM_STACK stack slots of every lane live in shared memory ([slot][thread],
// conflict free); only deeper ones fall back to local memory, whose loads were the top stall of this
// pass (18 % of samples, profiles/README.md v9).
#ifndef RT_MESH_SMEM_STACK
#define RT_MESH_SMEM_STACK 10     /* same-session A/B: +0.3 % on C4, +1.7 % on C5 (14: C4 -0.3 %, C5 +3.5 %) */
#endif
// Trees deeper than 32 (the CAP = 64 instantiation: the 10 M-triangle mesh is 37 deep) keep more
// slots there: +2.2 % on C5 in the r1b same-session A/B, and the shallow instantiation that C3/C4
// use is unchanged.
#ifndef RT_MESH_SMEM_STACK_DEEP
#define RT_MESH_SMEM_STACK_DEEP 14
#endif
#define RT_MESH_STK_PUT(slot, n_, a_, b_)                                                        \
    { if ((slot) < kSmemStack) { sm_node[(slot) * RT_BLOCK + threadIdx.x] = (n_);                \
          sm_t0[(slot) * RT_BLOCK + threadIdx.x] = (a_); sm_t1[(slot) * RT_BLOCK + threadIdx.x] = (b_); } \
      else { stk_node[(slot) - kSmemStack] = (n_); stk_t0[(slot) - kSmemStack] = (a_);           \
             stk_t1[(slot) - kSmemStack] = (b_); } }
#define RT_MESH_STK_GET(slot, n_, a_, b_)                                                        \
    { if ((slot) < kSmemStack) { (n_) = sm_node[(slot) * RT_BLOCK + threadIdx.x];                \
          (a_) = sm_t0[(slot) * RT_BLOCK + threadIdx.x]; (b_) = sm_t1[(slot) * RT_BLOCK + threadIdx.x]; } \
      else { (n_) = stk_node[(slot) - kSmemStack]; (a_) = stk_t0[(slot) - kSmemStack];           \
             (b_) = stk_t1[(slot) - kSmemStack]; } }
template <int CAP, bool ANY, bool COUNT, class IO>
__device__ __forceinline__ void trace_mesh(const DScene& sc, const IO& io, const SplitBufs& sb, const SplitPass& ps, WorkCount& wc)
{
    constexpr int kSmemStack = CAP > 32 ? RT_MESH_SMEM_STACK_DEEP : RT_MESH_SMEM_STACK;
    static_assert(kSmemStack >= 1 && kSmemStack < CAP, "shared-memory stack slots");
    __shared__ uint32_t sm_node[kSmemStack * RT_BLOCK];
    __shared__ float sm_t0[kSmemStack * RT_BLOCK];
    __shared__ float sm_t1[kSmemStack * RT_BLOCK];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1;
    const uint32_t n = *ps.in_count;

    uint32_t stk_node[CAP];
    float stk_t0[CAP];
    float stk_t1[CAP];

    bool active = false;
    bool exhausted = false;
    uint32_t tag = 0;
    float tmax = 0.0f;
    LocalRay r1;
    r1.o = r1.d = r1.inv = mk(0.0f, 0.0f, 0.0f);
    r1.neg = 0;
    r1.plain = false;
    float best = 0.0f;           // m_t
    int32_t best_rec = -1;       // triangle record accepted in this mesh, if any
    bool any_hit = false;
    uint32_t mesh_shape = 0, meta = 0;
    const DNode* mesh_nodes = sc.mesh_nodes;
    int sp = 0;
    bool have_cur = false;       // cur_*: the next node to pop, when it did not go through the stack
    uint32_t cur_node = 0;
    float cur_t0 = 0.0f, cur_t1 = 0.0f;
    bool top_valid = false;      // RT_MESH_TOPCACHE: the newest stack entry lives in top_* (slot sp - 1)
    uint32_t top_node = 0;
    float top_t0 = 0.0f, top_t1 = 0.0f;
    bool parked = false;
    uint32_t park_word = 0, park_count = 0;

    for (;;)
    {
        uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= RT_MESH_REFILL_MIN))
        {
            uint32_t need = __popc(idle);
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(ps.cursor, need);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + need >= n)
                exhausted = true;
            uint32_t j = base + __popc(idle & lt_mask);
            if (!active && j < n)
            {
                active = true;
                parked = false;
                have_cur = false;
                top_valid = false;
                tag = ps.in_queue[j];
                float4 mo = split_rec(sb, tag)[0], md = split_rec(sb, tag)[1];
                tmax = mo.w;            // (any hit)
                best = mo.w;            // (closest hit)
                best_rec = -1;
                any_hit = false;
                meta = __float_as_uint(md.w);
                mesh_shape = meta & 0xffffffu;
                DShape sh = load_shape(sc, mesh_shape);
                DMesh m = sc.meshes[sh.geom];
                // the top pass already warped the ray into mesh-local space and found
                // that it enters the root; redo the root's slab test (same inputs, same
                // bits) only to get the clipped range its children inherit
                r1.o = xyz4(mo);
                r1.d = xyz4(md);
                local_ray_finish(r1);
                mesh_nodes = sc.mesh_nodes + m.first_node;
                DNode root = load_node(mesh_nodes, 0);
                uint32_t rflags = __float_as_uint(root.q1.w);
                uint32_t rword = __float_as_uint(root.q1.z);
                sp = 0;
                if (rflags & RT_NODE_LEAF)
                {
                    parked = true;
                    park_word = rword;
                    park_count = rflags >> 3;
                }
                else
                {
                    float t0 = RT_RAY_TMIN, t1 = ANY ? tmax : best;
                    box_test(root.q0, root.q1, r1.o, r1.inv, t0, t1);
                    bool neg = (r1.neg >> (rflags & RT_NODE_AXIS)) & 1u;
#if RT_MESH_TOPCACHE
                    top_node = neg ? rword + 1 : rword; top_t0 = t0; top_t1 = t1;
                    top_valid = true;
#else
                    RT_MESH_STK_PUT(0, neg ? rword + 1 : rword, t0, t1);
#endif
                    sp = 1;
                    cur_node = neg ? rword : rword + 1; cur_t0 = t0; cur_t1 = t1;
                    have_cur = true;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0)
        {
            if (exhausted)
                break;
            continue;
        }
#if RT_MESH_PLAIN_SLABS
        // warp-uniform: all rays in flight are NaN-free (practically always; an axis-parallel
        // ray anywhere in the warp sends the whole warp through the std::min/max form)
        const bool warp_plain = __all_sync(0xffffffffu, !active || r1.plain);
#endif

        #pragma unroll 1
        for (int it = 0; it < RT_MESH_ADVANCE_STEPS && active && !parked && (have_cur || sp > 0); ++it)
        {
            // the near child of the node just entered is popped next: it stays in registers
            // instead of going through the (local-memory) stack
            uint32_t node_id;
            float t0, t1;
            if (have_cur)
            {
                node_id = cur_node; t0 = cur_t0; t1 = cur_t1;
                have_cur = false;
            }
            else
            {
                --sp;
#if RT_MESH_TOPCACHE
                if (top_valid)
                {
                    node_id = top_node; t0 = top_t0; t1 = top_t1;       // the newest entry never left the registers
                    top_valid = false;
                }
                else
#endif
                {
                    RT_MESH_STK_GET(sp, node_id, t0, t1);
                }
            }
            DNode nd = load_node(mesh_nodes, node_id);
            if (COUNT) wc.node_pops++;
            uint32_t flags = __float_as_uint(nd.q1.w);
            uint32_t word = __float_as_uint(nd.q1.z);
            if (flags & RT_NODE_LEAF)
            {
                parked = true;
                park_word = word;
                park_count = flags >> 3;
#if RT_MESH_PREFETCH
                // the triangles are tested only once enough lanes are parked: start their trip
                // from HBM now (48-byte records; a quad face is 96 bytes = two 64-byte pieces)
                {
                    const char* rec = reinterpret_cast<const char*>(sc.tris + (size_t)word * 3);
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(rec));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(rec + 64));
                }
#endif
                break;
            }
            if (!ANY)
            {
                if (t0 >= best)
                    continue;
                if (t1 > best)
                    t1 = best;
            }
#if RT_MESH_PLAIN_SLABS
            if (!(warp_plain ? box_test_plain(nd.q0, nd.q1, r1.o, r1.inv, t0, t1)
                             : box_test(nd.q0, nd.q1, r1.o, r1.inv, t0, t1)))
                continue;
#else
            if (!box_test(nd.q0, nd.q1, r1.o, r1.inv, t0, t1))
                continue;
#endif
            uint32_t axis = flags & RT_NODE_AXIS;
            bool neg = (r1.neg >> axis) & 1u;
            uint32_t near_id = neg ? word : word + 1;
            uint32_t far_id = neg ? word + 1 : word;
#if RT_MESH_TOPCACHE
            if (top_valid)
            {
                RT_MESH_STK_PUT(sp - 1, top_node, top_t0, top_t1);
            }
            top_node = far_id; top_t0 = t0; top_t1 = t1;
            top_valid = true;
#else
            RT_MESH_STK_PUT(sp, far_id, t0, t1);
#endif
            ++sp;
            cur_node = near_id; cur_t0 = t0; cur_t1 = t1;
            have_cur = true;
        }

        const uint32_t m_tri = __ballot_sync(0xffffffffu, parked);
        const uint32_t m_adv = __ballot_sync(0xffffffffu, active && !parked && (have_cur || sp > 0));
        const bool do_tri = m_tri != 0 && (m_adv == 0 || exhausted || __popc(m_tri) >= RT_MESH_SERVICE_MIN);

        if (do_tri && parked)
        {
            for (uint32_t k = 0; k < park_count; ++k)
            {
                V3 p0, p1, p2;
                uint32_t w0, w1, w2;
                load_tri(sc, park_word + k, p0, p1, p2, w0, w1, w2);
                if (COUNT) wc.tri_tests++;
                float t, beta, gamma;
                if (tri_closest(r1.o, r1.d, p0, p1, p2, ANY ? tmax : best, t, beta, gamma))
                {
                    if (ANY) { any_hit = true; sp = 0; have_cur = false; top_valid = false; break; }
                    best = t;
                    best_rec = (int32_t)(park_word + k);
                    if (sc.stage6) break;        // S6 RMesh.h:204-209
                }
            }
            parked = false;
        }

        // retire: hand the slot back to a top-level resume pass (or finish a shadow ray)
        const bool done = active && !parked && !have_cur && sp == 0;
        if (__ballot_sync(0xffffffffu, done))
        {
            bool resume = false;
            if (done)
            {
                if (ANY && any_hit)
                {
                    WaveResult r;
                    r.t = tmax; r.shape = -1; r.tri_rec = -1; r.any_hit = true;
                    io.store(tag, r);
                }
                else
                {
                    if (best_rec >= 0)
                    {
                        // (t, shape, triangle record); the fourth word (mesh shape | sp) stays as the top pass wrote it
                        float* h = reinterpret_cast<float*>(split_rec(sb, tag) + 2);
                        *reinterpret_cast<float2*>(h) = make_float2(best, __int_as_float((int32_t)mesh_shape));
                        h[2] = __int_as_float(best_rec);
                    }
                    resume = true;
                }
                active = false;
            }
            warp_queue_push(ps.out_queue, ps.out_count, resume, tag);
        }
    }
}

// ---------------------------------------------------------------------------
// Mesh pass, sibling-pair form (default).
//
// trace_mesh above does one dependent 32-byte fetch per node popped: pop, fetch the node, wait,
// test its box, descend.  The by-line profile of config C5 (profiles/README.md, r01 v11) put 45 %
// of the stall samples on the first use of the node just fetched: the pass is a chain of memory
// round trips, ~24 per ray (one per pop, plus one for a leaf's triangles).
//
// The reference numbers the two children of a node b and b+1 (RAccel.h:366-371) and the upload
// puts every such pair on one aligned 64-byte piece (rt_scene.cuh).  So one fetch of the PAIR when
// a node is expanded gives both children's boxes, kinds (interior / leaf) and payloads (own child
// pair / triangle range) at once, and everything the reference decides when it later pops either
// child can be decided from data already in registers:
//   * the near child is popped right after the push (RAccel.h:540-556): its cull check and slab
//     test run at once; only if it passes is its own pair fetched -- the one dependent fetch;
//   * the far child's slab values do not depend on when it is popped, only the clip against the
//     current m_t does: A = max(slab_near, t0) and B = min(slab_far, t1) are computed at push time
//     with the reference's operations, and the pop evaluates `t0 >= m_t` and A <= min(B, m_t) --
//     the same numbers (min is associative on numbers; a NaN slab value makes both forms fail);
//     a far child that fails already at push time (A <= B false, or t0 >= m_t: m_t only shrinks)
//     fails at pop time too, so it is not even pushed;
//   * a leaf is never box-tested (RAccel.h:506-511): its triangle range comes from the pair, so a
//     leaf is popped without any fetch of its node.
// Every lane still pops exactly the reference's nodes in the reference's order (the work counters,
// which count a far child's pop when the reference would pop it, stay equal to the oracle's) and
// tests the same triangles against the same m_t, so hits are bit-identical; what changes is that a
// ray now waits for memory once per node EXPANDED (~11 per ray on C5) instead of once per node
// popped (~24).
// Stack entry (16 bytes, shared memory then local): interior (pair | axis << 29, t0, A, B);
// leaf (LEAF | first triangle record, triangle count, -, -).
// ---------------------------------------------------------------------------
#ifndef RT_MESH_PAIR
#define RT_MESH_PAIR 1
#endif
#ifndef RT_PAIR_SMEM_STACK
#define RT_PAIR_SMEM_STACK 8
#endif
#ifndef RT_PAIR_SMEM_STACK_DEEP
#define RT_PAIR_SMEM_STACK_DEEP 10
#endif
#ifndef RT_PAIR_ADVANCE_STEPS
#define RT_PAIR_ADVANCE_STEPS 3
#endif
#ifndef RT_PAIR_TWO_LEAF
#define RT_PAIR_TWO_LEAF 0      /* 1: park on both leaves of a two-leaf pair at once (same-session A/B: C5 -2 %, C4 +-0) */
#endif
#ifndef RT_PAIR_PREFETCH_FAR
#define RT_PAIR_PREFETCH_FAR 0  /* 1: L2 prefetch of a far child's own pair when it is pushed */
#endif
#define RT_PAIR_LEAF 0x80000000u
#define RT_PAIR_MAX_NODES (1u << 29)      /* pair index and split axis share a word */

template <bool PLAIN>
__device__ __forceinline__ void pair_slabs(float4 q0, float4 q1, V3 o, V3 inv, float& lo, float& hi)
{
    if (PLAIN)
    {
        V3 a = (mk(q0.x, q0.y, q0.z) - o) * inv;
        V3 b = (mk(q0.w, q1.x, q1.y) - o) * inv;
        lo = fmaxf(fmaxf(fminf(a.x, b.x), fminf(a.y, b.y)), fminf(a.z, b.z));
        hi = fminf(fminf(fmaxf(a.x, b.x), fmaxf(a.y, b.y)), fmaxf(a.z, b.z));
    }
    else
        box_slabs(q0, q1, o, inv, lo, hi);
}

template <int CAP, bool ANY, bool COUNT, class IO>
__device__ __forceinline__ void trace_mesh_pair(const DScene& sc, const IO& io, const SplitBufs& sb, const SplitPass& ps, WorkCount& wc)
{
    constexpr int kSmemStack = CAP > 32 ? RT_PAIR_SMEM_STACK_DEEP : RT_PAIR_SMEM_STACK;
    static_assert(kSmemStack >= 1 && kSmemStack < CAP, "shared-memory stack slots");
    __shared__ float4 sm_stack[kSmemStack * RT_BLOCK];
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1;
    const uint32_t n = *ps.in_count;

    float4 lm_stack[CAP - kSmemStack];

    bool active = false;
    bool exhausted = false;
    uint32_t tag = 0;
    float tmax = 0.0f;
    LocalRay r1;
    r1.o = r1.d = r1.inv = mk(0.0f, 0.0f, 0.0f);
    r1.neg = 0;
    r1.plain = false;
    float best = 0.0f;           // m_t
    int32_t best_rec = -1;       // triangle record accepted in this mesh, if any
    bool any_hit = false;
    uint32_t mesh_shape = 0;
    bool direct = false;         // no top-level work is pending behind this mesh: finish the ray here
    const DNode* mesh_nodes = sc.mesh_nodes;
    int sp = 0;
    bool have_cur = false;       // cur_*: an interior node whose box test passed, waiting to be expanded
    uint32_t cur_ref = 0;        // its child pair | split axis << 29
    float cur_t0 = 0.0f, cur_t1 = 0.0f;
    bool parked = false;
    uint32_t park_word = 0, park_count = 0;
    uint32_t park2_word = 0, park2_count = 0;   // a second leaf to test right after the first (both children leaves)

#define RT_PAIR_PUT(slot, v_)                                                           \
    { if ((slot) < kSmemStack) sm_stack[(slot) * RT_BLOCK + threadIdx.x] = (v_);        \
      else lm_stack[(slot) - kSmemStack] = (v_); }
#define RT_PAIR_GET(slot, v_)                                                           \
    { if ((slot) < kSmemStack) (v_) = sm_stack[(slot) * RT_BLOCK + threadIdx.x];        \
      else (v_) = lm_stack[(slot) - kSmemStack]; }

    for (;;)
    {
        uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= RT_MESH_REFILL_MIN))
        {
            uint32_t need = __popc(idle);
            uint32_t base = 0;
            if (lane == 0)
                base = atomicAdd(ps.cursor, need);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + need >= n)
                exhausted = true;
            uint32_t j = base + __popc(idle & lt_mask);
            if (!active && j < n)
            {
                active = true;
                parked = false;
                have_cur = false;
                park2_count = 0;
                tag = ps.in_queue[j];
                float4 mo = split_rec(sb, tag)[0], md = split_rec(sb, tag)[1];
                tmax = mo.w;            // (any hit)
                best = mo.w;            // (closest hit)
                best_rec = -1;
                any_hit = false;
                mesh_shape = __float_as_uint(md.w) & 0xffffffu;
                direct = (__float_as_uint(md.w) & RT_SPLIT_DIRECT) != 0u;
                DShape sh = load_shape(sc, mesh_shape);
                DMesh m = sc.meshes[sh.geom];
                // the top pass already warped the ray into mesh-local space and found that it enters
                // the root; redo the root's slab test (same inputs, same bits) only to get the clipped
                // range its children inherit
                r1.o = xyz4(mo);
                r1.d = xyz4(md);
                local_ray_finish(r1);
                mesh_nodes = sc.mesh_nodes + m.first_node;
                DNode root = load_node(mesh_nodes, 0);
                uint32_t rflags = __float_as_uint(root.q1.w);
                uint32_t rword = __float_as_uint(root.q1.z);
                sp = 0;
                if (rflags & RT_NODE_LEAF)
                {
                    parked = true;
                    park_word = rword;
                    park_count = rflags >> 3;
                }
                else
                {
                    float t0 = RT_RAY_TMIN, t1 = ANY ? tmax : best;
                    box_test(root.q0, root.q1, r1.o, r1.inv, t0, t1);
                    cur_ref = rword | ((rflags & RT_NODE_AXIS) << 29);
                    cur_t0 = t0;
                    cur_t1 = t1;
                    have_cur = true;
                }
            }
        }
        if (__ballot_sync(0xffffffffu, active) == 0)
        {
            if (exhausted)
                break;
            continue;
        }
        // warp-uniform: all rays in flight are NaN-free (see trace_mesh)
        const bool warp_plain = __all_sync(0xffffffffu, !active || r1.plain);

        #pragma unroll 1
        for (int it = 0; it < RT_PAIR_ADVANCE_STEPS && active && !parked && (have_cur || sp > 0); ++it)
        {
            // ---- find the next node to expand: pops that fail their (cheap, register-only) checks
            // are retired here, so that the lanes of the warp meet again at the fetch below
            while (!have_cur && sp > 0)
            {
                --sp;
                float4 e;
                RT_PAIR_GET(sp, e);
                const uint32_t ref = __float_as_uint(e.x);
                if (COUNT) wc.node_pops++;
                if (ref & RT_PAIR_LEAF)
                {
                    parked = true;
                    park_word = ref & ~RT_PAIR_LEAF;
                    park_count = __float_as_uint(e.y);
                    break;
                }
                float t1 = e.w;
                if (!ANY)
                {
                    if (e.y >= best)            // t0 >= m_t  (RAccel.h:523-526)
                        continue;
                    t1 = warp_plain ? fminf(e.w, best) : ((e.w > best) ? best : e.w);   // min(B, m_t)
                }
                if (!(e.z <= t1))               // A <= min(B, m_t)
                    continue;
                cur_ref = ref;
                cur_t0 = e.z;
                cur_t1 = t1;
                have_cur = true;
            }
            if (!have_cur)
                break;
            have_cur = false;

            // ---- expand: one 64-byte fetch brings both children
            const uint32_t pair = cur_ref & (RT_PAIR_MAX_NODES - 1u);
            const uint32_t axis = cur_ref >> 29;
            const float4* p = reinterpret_cast<const float4*>(mesh_nodes + pair);
            const float4 a0 = __ldg(p), a1 = __ldg(p + 1), b0 = __ldg(p + 2), b1 = __ldg(p + 3);
            float lo0, hi0, lo1, hi1;
            if (warp_plain)
            {
                pair_slabs<true>(a0, a1, r1.o, r1.inv, lo0, hi0);
                pair_slabs<true>(b0, b1, r1.o, r1.inv, lo1, hi1);
            }
            else
            {
                pair_slabs<false>(a0, a1, r1.o, r1.inv, lo0, hi0);
                pair_slabs<false>(b0, b1, r1.o, r1.inv, lo1, hi1);
            }
            const bool neg = (r1.neg >> axis) & 1u;      // near = first child when the direction is negative on the axis
            const float near_lo = neg ? lo0 : lo1, near_hi = neg ? hi0 : hi1;
            const float far_lo = neg ? lo1 : lo0, far_hi = neg ? hi1 : hi0;
            const uint32_t near_word = __float_as_uint(neg ? a1.z : b1.z), near_flags = __float_as_uint(neg ? a1.w : b1.w);
            const uint32_t far_word = __float_as_uint(neg ? b1.z : a1.z), far_flags = __float_as_uint(neg ? b1.w : a1.w);
            const float t0 = cur_t0, t1 = cur_t1;

            // Both children leaves (the bottom of the tree): the reference tests the near leaf's
            // triangles, pops the far leaf next (a leaf is never culled) and tests its triangles.
            // Park on both at once: one triangle phase, no stack traffic.
            if (RT_PAIR_TWO_LEAF && (far_flags & near_flags & RT_NODE_LEAF) != 0u)
            {
                if (COUNT) wc.node_pops++;
                parked = true;
                park_word = near_word;
                park_count = near_flags >> 3;
                park2_word = far_word;
                park2_count = far_flags >> 3;
                break;
            }
            // far child: pushed first (popped after everything below the near child)
            if (far_flags & RT_NODE_LEAF)
            {
                RT_PAIR_PUT(sp, make_float4(__uint_as_float(RT_PAIR_LEAF | far_word), __uint_as_float(far_flags >> 3), 0.0f, 0.0f));
                ++sp;
            }
            else
            {
                // what its pop will compute that does not depend on m_t: A = max(slab near, t0), B = min(slab far, t1)
                const float A = warp_plain ? fmaxf(far_lo, t0) : std_max(far_lo, t0);
                const float B = warp_plain ? fminf(far_hi, t1) : std_min(far_hi, t1);
                bool keep = A <= B;
                if (!ANY && t0 >= best)
                    keep = false;
                // (when counting work every far child is pushed, so that its pop is counted exactly when the
                // reference pops it -- an any-hit ray may stop before; the pop's own checks then drop it)
                if (keep || COUNT)
                {
                    RT_PAIR_PUT(sp, make_float4(__uint_as_float(far_word | ((far_flags & RT_NODE_AXIS) << 29)), t0, A, B));
                    ++sp;
#if RT_PAIR_PREFETCH_FAR
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(mesh_nodes + far_word));
#endif
                }
            }

            // near child: popped at once
            if (COUNT) wc.node_pops++;
            if (near_flags & RT_NODE_LEAF)
            {
                parked = true;
                park_word = near_word;
                park_count = near_flags >> 3;
                break;
            }
            if (!ANY && t0 >= best)
                continue;
            // (t1 <= m_t already: it is the parent's clipped range and m_t has not changed since)
            const float A = warp_plain ? fmaxf(near_lo, t0) : std_max(near_lo, t0);
            const float B = warp_plain ? fminf(near_hi, t1) : std_min(near_hi, t1);
            if (A <= B)
            {
                cur_ref = near_word | ((near_flags & RT_NODE_AXIS) << 29);
                cur_t0 = A;
                cur_t1 = B;
                have_cur = true;
            }
        }

        const uint32_t m_tri = __ballot_sync(0xffffffffu, parked);
        const uint32_t m_adv = __ballot_sync(0xffffffffu, active && !parked && (have_cur || sp > 0));
        const bool do_tri = m_tri != 0 && (m_adv == 0 || exhausted || __popc(m_tri) >= RT_MESH_SERVICE_MIN);

        if (do_tri && parked)
        {
            #pragma unroll 1
            for (int leaf = 0; leaf < 2; ++leaf)
            {
                for (uint32_t k = 0; k < park_count; ++k)
                {
                    V3 p0, p1, p2;
                    uint32_t w0, w1, w2;
                    load_tri(sc, park_word + k, p0, p1, p2, w0, w1, w2);
                    if (COUNT) wc.tri_tests++;
                    float t, beta, gamma;
                    if (tri_closest(r1.o, r1.d, p0, p1, p2, ANY ? tmax : best, t, beta, gamma))
                    {
                        if (ANY) { any_hit = true; sp = 0; have_cur = false; park2_count = 0; break; }
                        best = t;
                        best_rec = (int32_t)(park_word + k);
                        if (sc.stage6) break;        // S6 RMesh.h:204-209
                    }
                }
                if (park2_count == 0)
                    break;
                // the far leaf of a two-leaf pair: popped now
                if (COUNT) wc.node_pops++;
                park_word = park2_word;
                park_count = park2_count;
                park2_count = 0;
            }
            parked = false;
        }

        // retire: hand the slot back to a top-level resume pass (or finish a shadow ray)
        const bool done = active && !parked && !have_cur && sp == 0;
        if (__ballot_sync(0xffffffffu, done))
        {
            bool resume = false;
            if (done)
            {
                if (ANY && any_hit)
                {
                    WaveResult r;
                    r.t = tmax; r.shape = -1; r.tri_rec = -1; r.any_hit = true;
                    io.store(tag, r);
                }
                else if (direct)
                {
                    // this mesh was the last thing the top-level walk had to look at: the ray is finished
                    // (a shadow ray that found nothing has nothing to report)
                    if (!ANY)
                    {
                        WaveResult r;
                        r.any_hit = false;
                        if (best_rec >= 0)
                        {
                            r.t = best; r.shape = (int32_t)mesh_shape; r.tri_rec = best_rec;
                        }
                        else
                        {
                            const float4 h = split_rec(sb, tag)[2];
                            r.t = h.x; r.shape = __float_as_int(h.y); r.tri_rec = __float_as_int(h.z);
                        }
                        io.store(tag, r);
                    }
                }
                else
                {
                    if (best_rec >= 0)
                    {
                        // (t, shape, triangle record); the fourth word (mesh shape | sp) stays as the top pass wrote it
                        float* h = reinterpret_cast<float*>(split_rec(sb, tag) + 2);
                        *reinterpret_cast<float2*>(h) = make_float2(best, __int_as_float((int32_t)mesh_shape));
                        h[2] = __int_as_float(best_rec);
                    }
                    resume = true;
                }
                active = false;
            }
            warp_queue_push(ps.out_queue, ps.out_count, resume, tag);
        }
    }
#undef RT_PAIR_PUT
#undef RT_PAIR_GET
}

#endif // RAYITO_B200_RT_SPLIT_CUH
