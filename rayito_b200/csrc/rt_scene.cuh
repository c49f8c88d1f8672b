// RtScene: the device-resident scene behind the C ABI, and its upload path.
//
// rt_scene_create() validates an RtSceneDesc (what ShapeSet::prepare() leaves
// behind in the reference, RScene.h:186-205 / RMesh.h:89-129), expands polygon
// faces into fan-triangle records, re-encodes mesh leaves to point at them, and
// copies everything into ONE HBM arena with a single host->device transfer.
#ifndef RAYITO_B200_RT_SCENE_CUH
#define RAYITO_B200_RT_SCENE_CUH

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

#include "rt_device.cuh"
#include "rt_pool.cuh"

extern thread_local std::string g_rt_error;

int rt_fail(int code, const std::string& what);
int rt_cuda_fail(cudaError_t e, const char* where);

#define RT_CUDA(call)                                               \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return rt_cuda_fail(e__, #call);    \
    } while (0)

#define RT_SCENE_CURSORS 16

struct RtScene
{
    int device;
    void* arena;                // one allocation holding every array below
    size_t arena_bytes;
    size_t arena_alloc, work_alloc, cursor_alloc;   // real sizes of the pool blocks behind arena / d_work / d_cursor
    DScene d;                   // device pointers into the arena
    int stack_cap;              // traversal stack entries needed (both levels)
    int top_stack_need;         // ... by the top level alone
    int mesh_stack_need;        // ... by the deepest mesh alone
    uint32_t num_mesh_shapes;   // finite shapes that are meshes (max suspensions per ray)
    uint32_t num_tris;
    uint32_t num_faces;
    std::vector<uint32_t> light_shapes;
    std::vector<uint32_t> light_types;  // RT_SHAPE_* of every light, findLights() order
    std::vector<uint32_t> shape_types;  // RT_SHAPE_* of every shape (finite, then infinite)
    std::vector<uint32_t> shape_brdfs;  // RT_BRDF_* of every shape's material
    bool has_lambert, has_glossy;       // BRDF kinds among the scene's materials
    float upload_ms;
    float bvh_build_ms;                      // device time of the face-BVH build (RT_SCENE_BUILD_MESH_BVH), else 0
    int mesh_depth;                          // deepest leaf of any face BVH (-1: no mesh)
    std::vector<uint32_t> mesh_first_node;   // device node slot of every mesh's root
    std::vector<uint32_t> mesh_first_face, mesh_faces;
    // scratch for the host-buffer entry points (grown on demand)
    void* scratch_in;
    void* scratch_out;
    size_t scratch_in_bytes, scratch_out_bytes;
    uint64_t* d_work;           // RT_WORK_COUNTERS counters
    uint32_t* d_cursor;         // RT_SCENE_CURSORS queue cursors of the ray-batch entry points, one per call in flight
    std::atomic<uint32_t> cursor_next;
    struct RenderBuffers* render;   // wavefront state (rt_render.cuh), lazily created
    bool dynamic_top;           // this render: use the per-lane top-level pass even if a walk table exists
};

namespace rt_detail
{

// Deepest leaf (root = 0) of a reference-format BVH; -1 if malformed.
inline int bvh_depth_walk(const RtBvhNode* nodes, uint32_t count, uint32_t num_prims, std::string& why)
{
    if (count == 0)
        return 0;
    std::vector<std::pair<uint32_t, int> > todo;
    todo.push_back(std::make_pair(0u, 0));
    int deepest = 0;
    size_t visited = 0;
    while (!todo.empty())
    {
        std::pair<uint32_t, int> cur = todo.back();
        todo.pop_back();
        if (++visited > count)
        {
            why = "BVH has a cycle or shared children";
            return -1;
        }
        const RtBvhNode& n = nodes[cur.first];
        if (cur.second > deepest) deepest = cur.second;
        if (n.flags & RT_NODE_LEAF)
        {
            if (n.first_child_or_prim >= num_prims)
            {
                why = "BVH leaf names a primitive that does not exist";
                return -1;
            }
            continue;
        }
        if ((n.flags & RT_NODE_AXIS) == 3u)
        {
            why = "BVH node with split axis 3";
            return -1;
        }
        if ((uint64_t)n.first_child_or_prim + 1 >= count)
        {
            why = "BVH child index out of range";
            return -1;
        }
        todo.push_back(std::make_pair(n.first_child_or_prim + 1, cur.second + 1));
        todo.push_back(std::make_pair(n.first_child_or_prim, cur.second + 1));
    }
    return deepest;
}

// Host worker threads for the upload path (same switch as the host library:
// RAYITO_B200_HOST_THREADS, else the hardware concurrency, at most 32)
inline unsigned host_threads()
{
    const char* env = std::getenv("RAYITO_B200_HOST_THREADS");
    if (env != NULL)
    {
        long n = std::strtol(env, NULL, 10);
        if (n >= 1)
            return n > 256 ? 256u : (unsigned)n;
    }
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    return n > 32 ? 32u : n;
}

// body(begin, end) over contiguous pieces of [0, n); pieces of at least `grain` items
template <typename Body>
inline void parallel_ranges(size_t n, size_t grain, Body body)
{
    size_t pieces = grain ? (n + grain - 1) / grain : 1;
    unsigned threads = host_threads();
    if (pieces > threads) pieces = threads;
    if (pieces <= 1)
    {
        body((size_t)0, n);
        return;
    }
    std::vector<std::thread> workers;
    workers.reserve(pieces - 1);
    size_t started = 1;             // pieces [1, started) run on workers; the rest falls to this thread
    for (; started < pieces; ++started)
    {
        try
        {
            workers.push_back(std::thread(body, n * started / pieces, n * (started + 1) / pieces));
        }
        catch (const std::system_error&)
        {
            break;                  // no more threads to be had: the caller does the remaining pieces itself
        }
    }
    body((size_t)0, n / pieces);
    if (started < pieces)
        body(n * started / pieces, n);
    for (size_t i = 0; i < workers.size(); ++i)
        workers[i].join();
}

// Lays the scene arrays out in one block (256-byte aligned each).  Arrays are first
// declared -- put() for data that exists, reserve() for records the caller builds in
// place -- then finish() allocates the block once and copies the declared data, the
// large arrays piecewise on the worker threads.  No array is staged twice and nothing
// is zero-filled first: at 10 M triangles the block is 660 MB.
struct ArenaBuilder
{
    struct Segment { size_t offset; const void* src; size_t bytes; };
    std::vector<Segment> segments;
    size_t total;
    unsigned char* block;

    ArenaBuilder() : total(0), block(NULL) { }

    size_t reserve(size_t n)
    {
        size_t off = (total + 255) & ~(size_t)255;
        Segment seg = { off, NULL, n ? n : 16 };
        segments.push_back(seg);
        total = off + seg.bytes;
        return off;
    }
    // `src` must stay valid until finish()
    size_t put(const void* src, size_t n)
    {
        size_t off = reserve(n);
        if (n) segments.back().src = src;
        return off;
    }
    // `storage`: at least `total` bytes, owned by the caller
    bool finish(void* storage)
    {
        block = static_cast<unsigned char*>(storage);
        if (block == NULL)
            return false;
        size_t end = 0;
        for (size_t i = 0; i < segments.size(); ++i)
        {
            const Segment& seg = segments[i];
            std::memset(block + end, 0, seg.offset - end);      // alignment gap
            end = seg.offset + seg.bytes;
            if (seg.src == NULL)
            {
                if (seg.bytes <= 16) std::memset(block + seg.offset, 0, seg.bytes);
                continue;
            }
            unsigned char* dst = block + seg.offset;
            const unsigned char* src = static_cast<const unsigned char*>(seg.src);
            parallel_ranges(seg.bytes, (size_t)8 << 20, [dst, src](size_t b, size_t e) { std::memcpy(dst + b, src + b, e - b); });
        }
        return true;
    }
    template <typename V> V* at(size_t offset) { return reinterpret_cast<V*>(block + offset); }
};

struct HostScratch
{
    void* data;
    size_t bytes;
    HostScratch() : data(NULL), bytes(0) { }
    ~HostScratch() { std::free(data); }
    void* get(size_t n)
    {
        if (n > bytes)
        {
            std::free(data);
            data = std::malloc(n);
            bytes = data ? n : 0;
        }
        return data;
    }
    void release() { std::free(data); data = NULL; bytes = 0; }
};
inline HostScratch& validate_scratch()
{
    static thread_local HostScratch scratch;
    return scratch;
}

// bvh_depth for large trees.  The reference numbers children after their parent
// (RAccel.h:366-371), which makes the checks data-parallel.  Pass A, on the worker threads:
// every node is examined on its own (leaf primitive, split axis, child range, children
// numbered after the parent) and leaves its child index in a compact table and its own index
// in its children's "who points at me" slots.  Pass B: a node whose children do not point
// back at it shares them with another node.  With all edges pointing forward and one parent
// per node, one ascending sweep over the compact table gives every depth.  Trees that are
// not numbered that way (another builder) take the serial walk above.
inline int bvh_depth(const RtBvhNode* nodes, uint32_t count, uint32_t num_prims, std::string& why)
{
    if (count < (1u << 16))
        return bvh_depth_walk(nodes, count, num_prims, why);
    const uint32_t kNone = 0xffffffffu;
    // Scratch kept per calling thread between calls (10 bytes per node: fresh pages would cost
    // more in page faults than the checks themselves): [0, count) first child (kNone for a leaf),
    // [count, 2 count) the node pointing at me, then one depth and one "reached" byte per node.
    static_assert(sizeof(std::atomic<uint32_t>) == sizeof(uint32_t), "atomic<uint32_t> must be a plain word");
    void* scratch = validate_scratch().get((size_t)count * 10);
    if (scratch == NULL)
    {
        why = "out of host memory validating a BVH";
        return -1;
    }
    std::atomic<uint32_t>* child = static_cast<std::atomic<uint32_t>*>(scratch);
    std::atomic<uint32_t>* pointer = child + count;
    unsigned char* depth = reinterpret_cast<unsigned char*>(child + 2 * (size_t)count);
    unsigned char* reached = depth + count;
    const std::memory_order relaxed = std::memory_order_relaxed;
    parallel_ranges(count, 1u << 16, [=](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) pointer[i].store(kNone, relaxed);
    });
    // 0 ok, 1 not forward-numbered (fall back), 2.. malformed
    std::atomic<int> verdict(0);
    auto report = [&verdict](int mine) {
        int seen = verdict.load();
        while (seen < mine && !verdict.compare_exchange_weak(seen, mine)) { }
    };
    parallel_ranges(count, 1u << 16, [=, &report](size_t b, size_t e) {
        int mine = 0;
        for (size_t i = b; i < e; ++i)
        {
            const RtBvhNode& n = nodes[i];
            uint32_t c = kNone;
            if (n.flags & RT_NODE_LEAF)
            {
                if (n.first_child_or_prim >= num_prims) mine = mine > 2 ? mine : 2;
            }
            else if ((n.flags & RT_NODE_AXIS) == 3u)
                mine = mine > 3 ? mine : 3;
            else if ((uint64_t)n.first_child_or_prim + 1 >= count)
                mine = mine > 4 ? mine : 4;
            else if (n.first_child_or_prim <= i)
                mine = mine > 1 ? mine : 1;
            else
            {
                c = n.first_child_or_prim;
                pointer[c].store((uint32_t)i, relaxed);
                pointer[c + 1].store((uint32_t)i, relaxed);
            }
            child[i].store(c, relaxed);
        }
        if (mine != 0) report(mine);
    });
    if (verdict.load() == 0)
        parallel_ranges(count, 1u << 16, [=, &report](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i)
            {
                const uint32_t c = child[i].load(relaxed);
                if (c != kNone && (pointer[c].load(relaxed) != i || pointer[c + 1].load(relaxed) != i))
                {
                    report(5);
                    return;
                }
            }
        });
    if (verdict.load() == 0 && pointer[0].load(relaxed) != kNone)
        report(5);              // (cannot happen with forward edges; kept for symmetry with the walk)
    switch (verdict.load())
    {
    case 0: break;
    case 1: return bvh_depth_walk(nodes, count, num_prims, why);
    case 2: why = "BVH leaf names a primitive that does not exist"; return -1;
    case 3: why = "BVH node with split axis 3"; return -1;
    case 4: why = "BVH child index out of range"; return -1;
    default: why = "BVH has a cycle or shared children"; return -1;
    }
    // forward edges, one parent each, the root none: every node somebody points at hangs off
    // the root.  Nodes nobody points at are unreachable and do not count.
    std::memset(depth, 0, (size_t)count * 2);
    reached[0] = 1;
    int deepest = 0;
    for (uint32_t i = 0; i < count; ++i)
    {
        if (!reached[i])
            continue;
        const int d = depth[i];
        if (d > deepest) deepest = d;
        const uint32_t c = child[i].load(relaxed);
        if (c == kNone)
            continue;
        if (d >= 200)
            return 201;             // far past the 49 the callers accept; keeps the byte from wrapping
        depth[c] = depth[c + 1] = (unsigned char)(d + 1);
        reached[c] = reached[c + 1] = 1;
    }
    return deepest;
}

inline float vlen(const float* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

} // namespace rt_detail

namespace rt_build
{
// rt_build.cuh: the face BVH of one mesh built on the device, node for node the reference's tree
inline int build_mesh(int device, DNode* nodes, const float4* tris, const uint32_t* fft, uint32_t first_face, uint32_t num_faces,
                      int* depth, float* build_ms);
}

// flags: RT_SCENE_* of include/rayito_b200.h.  RT_SCENE_BUILD_MESH_BVH: the face BVH of every mesh that comes
// without nodes (RtMesh.num_nodes == 0) is built on the device after the upload (rt_build.cuh); meshes that
// bring their nodes keep them.
inline int rt_scene_build(const RtSceneDesc* desc, int device, uint32_t flags, RtScene** out_scene)
{
    using namespace rt_detail;
    if (desc == NULL || out_scene == NULL)
        return rt_fail(RT_ERR_ARG, "null argument");
    *out_scene = NULL;
    if (flags & ~(uint32_t)RT_SCENE_BUILD_MESH_BVH)
        return rt_fail(RT_ERR_ARG, "unknown scene flags");
    bool dev_build = (flags & RT_SCENE_BUILD_MESH_BVH) != 0;
    if (dev_build && desc->semantics != RT_SEMANTICS_STAGE7)
    {
        // nothing to build is fine; a mesh without nodes is not
        for (uint32_t m = 0; desc->meshes != NULL && m < desc->num_meshes; ++m)
            if (desc->meshes[m].num_nodes == 0 && desc->meshes[m].num_faces > 0)
                return rt_fail(RT_ERR_UNSUPPORTED, "the device BVH build reproduces the Stage 7 builder only (Stage 6 roots its face BVH in the all-vertex box)");
        dev_build = false;
    }
    if (desc->abi_version != RT_ABI_VERSION)
        return rt_fail(RT_ERR_ARG, "RtSceneDesc.abi_version mismatch");
    const uint32_t num_shapes = desc->num_finite + desc->num_infinite;
    if (num_shapes > 0 && desc->shapes == NULL)
        return rt_fail(RT_ERR_ARG, "shapes is null");
    if (desc->set_xform >= desc->num_xforms)
        return rt_fail(RT_ERR_ARG, "set_xform out of range");
    if (desc->num_finite > 2 && desc->num_top_nodes == 0)
        return rt_fail(RT_ERR_ARG, "more than two finite shapes need a top-level BVH (RScene.h:135)");
    if (desc->num_finite <= 2 && desc->num_top_nodes != 0)
        return rt_fail(RT_ERR_ARG, "a top-level BVH is only used with more than two finite shapes");

    if (desc->semantics != RT_SEMANTICS_STAGE7 && desc->semantics != RT_SEMANTICS_STAGE6)
        return rt_fail(RT_ERR_ARG, "unknown RtSceneDesc.semantics");
    // every array with a non-zero count must be there
    if ((desc->num_xforms && desc->xforms == NULL) ||
        (desc->num_keys && (desc->key_time == NULL || desc->key_scale == NULL || desc->key_rotation == NULL || desc->key_translation == NULL)) ||
        (desc->num_top_nodes && desc->top_nodes == NULL) || (desc->num_mesh_nodes && desc->mesh_nodes == NULL) ||
        (desc->num_planes && desc->planes == NULL) || (desc->num_spheres && desc->spheres == NULL) ||
        (desc->num_rects && desc->rects == NULL) || (desc->num_meshes && desc->meshes == NULL) ||
        (desc->num_vertices && desc->vertices == NULL) || (desc->num_normals && desc->normals == NULL) ||
        (desc->num_faces && (desc->face_start == NULL || desc->face_has_normals == NULL)) ||
        (desc->num_indices && (desc->vertex_index == NULL || desc->normal_index == NULL)) ||
        (desc->num_cdf && desc->face_area_cdf == NULL) || (desc->num_materials && desc->materials == NULL) ||
        (desc->num_lights && desc->lights == NULL))
        return rt_fail(RT_ERR_ARG, "RtSceneDesc: an array with a non-zero count is null");
    for (uint32_t i = 0; i < desc->num_xforms; ++i)
    {
        const RtXform& x = desc->xforms[i];
        if (desc->semantics == RT_SEMANTICS_STAGE6 && x.num_keys != 0)
            return rt_fail(RT_ERR_ARG, "Stage 6 semantics has no transforms: every xform must be keyless");
        if ((uint64_t)x.first_key + x.num_keys > desc->num_keys)
            return rt_fail(RT_ERR_ARG, "transform key range out of bounds");
        for (uint32_t k = 1; k < x.num_keys; ++k)
            if (!(desc->key_time[x.first_key + k - 1] < desc->key_time[x.first_key + k]))
                return rt_fail(RT_ERR_ARG, "transform key times must be strictly increasing");
    }

    // Shapes
    // Classify transforms (rt_device.cuh RT_XF_*): exact comparisons on the key values
    std::vector<uint32_t> xf_kind(desc->num_xforms, RT_XF_GENERAL);
    for (uint32_t i = 0; i < desc->num_xforms; ++i)
    {
        const RtXform& x = desc->xforms[i];
        bool translate_only = true, unit_scale = true;
        for (uint32_t k = x.first_key; k < x.first_key + x.num_keys; ++k)
        {
            const float* r = desc->key_rotation + 4 * (size_t)k;
            const float* sc3 = desc->key_scale + 3 * (size_t)k;
            // +0 only: a -0 component would not behave like the keyless identity
            uint32_t bits[3];
            std::memcpy(bits, r + 1, 12);
            const bool unit = sc3[0] == 1.0f && sc3[1] == 1.0f && sc3[2] == 1.0f;
            if (!unit)
                unit_scale = false;
            if (!(r[0] == 1.0f && bits[0] == 0 && bits[1] == 0 && bits[2] == 0 && unit))
                translate_only = false;
        }
        if (translate_only)
            xf_kind[i] = x.num_keys <= 1 ? RT_XF_STATIC : RT_XF_TRANSLATE;
        else if (unit_scale)
            xf_kind[i] = RT_XF_RIGID;
    }
    // Rows of the per-sample transform cache (rt_device.cuh): every transform with two or more keys,
    // two- and three-float4 entries on 32-byte boundaries so that an entry is one or two DRAM sectors
    std::vector<uint4> anim;
    std::vector<uint32_t> xf_cache_slot(desc->num_xforms, 0u);      // row offset + 1
    uint32_t anim_stride = 0;
    // Which transforms get a cache entry, and who reads it (same-session A/Bs, profiles/README.md round 2):
    //   * interpolated ROTATIONS only -- an interpolated translation is a key search and three lerps, cheaper to
    //     redo than to fetch (C4: rotations +2.5 %, everything +1.6 %);
    //   * only when the scene has at most four of them: every entry is evaluated for every sample whether a ray of
    //     that sample meets the shape or not (scene 2, ten tumbling boxes: -9 %; everything cached -20 %);
    //   * the shading kernels read every entry (normal of the hit, light pdfs); the TRAVERSAL kernels read only the
    //     entries of shapes that are small in the scene (top-level box under a quarter of the root box's area).  A
    //     shape that fills the scene is entered by most rays of a warp, so its in-place evaluation already runs
    //     with the warp converged and a fetch only adds memory traffic (the 10 M-triangle mesh of C5: traversal
    //     -1.5 % with the fetch, frame +2 % from the shading side).
    // RAYITO_B200_XFORM_CACHE = none | auto (default) | rotations | all overrides the choice (A/B runs).
    const char* cache_env = std::getenv("RAYITO_B200_XFORM_CACHE");
    const int cache_mode = cache_env == NULL ? 3 : cache_env[0] == 'n' ? 0 : cache_env[0] == 'r' ? 1 : cache_env[0] == 'a' && cache_env[1] == 'l' ? 2 : 3;
    std::vector<float> xf_area_ratio(desc->num_xforms, 2.0f);      // largest top-level box of a shape using the transform / root box
    if (desc->num_top_nodes > 0)
    {
        auto half_area = [](const RtBvhNode& n) {
            float dx = n.bbox_max[0] - n.bbox_min[0], dy = n.bbox_max[1] - n.bbox_min[1], dz = n.bbox_max[2] - n.bbox_min[2];
            return dx * dy + dy * dz + dz * dx;
        };
        const float root = half_area(desc->top_nodes[0]);
        std::vector<float> seen(desc->num_xforms, -1.0f);
        for (uint32_t i = 0; i < desc->num_top_nodes; ++i)
        {
            const RtBvhNode& n = desc->top_nodes[i];
            if (!(n.flags & RT_NODE_LEAF) || n.first_child_or_prim >= desc->num_finite)
                continue;
            const uint32_t xf = desc->shapes[n.first_child_or_prim].xform;
            if (xf >= desc->num_xforms)
                continue;           // (reported below)
            const float ratio = root > 0.0f ? half_area(n) / root : 1.0f;
            if (ratio > seen[xf]) seen[xf] = ratio;
        }
        for (uint32_t i = 0; i < desc->num_xforms; ++i)
            if (seen[i] >= 0.0f) xf_area_ratio[i] = seen[i];
    }
    uint32_t num_rotating = 0;
    for (uint32_t i = 0; i < desc->num_xforms; ++i)
        if (desc->xforms[i].num_keys >= 2 && i != desc->set_xform && xf_kind[i] != RT_XF_TRANSLATE)
            ++num_rotating;
    std::vector<uint32_t> xf_trav_cached(desc->num_xforms, 0u);     // traversal kernels may read the entry
    uint32_t cached_entries = 0;
    (void)cached_entries;
    if (desc->semantics == RT_SEMANTICS_STAGE7 && cache_mode != 0 && !(cache_mode == 3 && num_rotating > 4))
    {
        for (int pass = 0; pass < 2; ++pass)        // wide entries first (aligned), then the one-float4 translations
            for (uint32_t i = 0; i < desc->num_xforms; ++i)
            {
                if (desc->xforms[i].num_keys < 2 || i == desc->set_xform)
                    continue;
                const bool narrow = xf_kind[i] == RT_XF_TRANSLATE;
                if (narrow != (pass == 1) || (narrow && cache_mode != 2))
                    continue;
                xf_trav_cached[i] = (cache_mode != 3 || xf_area_ratio[i] < 0.25f) ? 1u : 0u;
                ++cached_entries;
                uint32_t width = narrow ? 1u : xf_kind[i] == RT_XF_RIGID ? 2u : 3u;
                if (!narrow && (anim_stride & 1u))
                    ++anim_stride;
                anim.push_back(make_uint4(i, anim_stride, xf_kind[i], 0u));
                xf_cache_slot[i] = anim_stride + 1u;
                anim_stride += width;
            }
        anim_stride = (anim_stride + 1u) & ~1u;
        if (anim_stride > RT_XF_CACHE_MAX_STRIDE)
        {
            anim.clear();
            std::fill(xf_cache_slot.begin(), xf_cache_slot.end(), 0u);
            std::fill(xf_trav_cached.begin(), xf_trav_cached.end(), 0u);
            anim_stride = 0;
        }
    }

    std::vector<DShapeMem> shapes(num_shapes);
    for (uint32_t i = 0; i < num_shapes; ++i)
    {
        const RtShape& s = desc->shapes[i];
        uint32_t limit = s.type == RT_SHAPE_PLANE ? desc->num_planes :
                         s.type == RT_SHAPE_SPHERE ? desc->num_spheres :
                         s.type == RT_SHAPE_RECT ? desc->num_rects :
                         s.type == RT_SHAPE_MESH ? desc->num_meshes : 0;
        if (s.geom >= limit || s.xform >= desc->num_xforms || s.material >= desc->num_materials)
            return rt_fail(RT_ERR_ARG, "shape refers to missing geometry, transform or material");
        if ((i < desc->num_finite) == (s.type == RT_SHAPE_PLANE))
            return rt_fail(RT_ERR_ARG, "planes must be the infinite shapes and only they");
        if (s.light >= (int32_t)desc->num_lights)
            return rt_fail(RT_ERR_ARG, "shape light index out of range");
        DShapeMem d;
        std::memset(&d, 0, sizeof(d));
        d.type_kind = s.type | (xf_kind[s.xform] << 8) | (xf_trav_cached[s.xform] << 10) | (xf_cache_slot[s.xform] << 16);
        d.geom = s.geom; d.xform = s.xform; d.material = s.material; d.light = s.light;
        if (xf_kind[s.xform] == RT_XF_STATIC && desc->xforms[s.xform].num_keys == 1)
        {
            const float* t = desc->key_translation + 3 * (size_t)desc->xforms[s.xform].first_key;
            d.tx = t[0]; d.ty = t[1]; d.tz = t[2];
        }
        shapes[i] = d;
    }
    for (uint32_t l = 0; l < desc->num_lights; ++l)
        if (desc->lights[l] >= num_shapes || desc->shapes[desc->lights[l]].light != (int32_t)l)
            return rt_fail(RT_ERR_ARG, "lights[] and RtShape.light disagree");

    // RAYITO_B200_TIMING=1: host-clock phases of the upload path on stderr
    const bool host_timing = std::getenv("RAYITO_B200_TIMING") != NULL;
    std::chrono::steady_clock::time_point hc[6];
    hc[0] = std::chrono::steady_clock::now();

    // BVH depths (the reference's fixed 50-entry stack, RAccel.h:379)
    std::string why;
    int top_depth = bvh_depth(desc->top_nodes, desc->num_top_nodes, desc->num_finite, why);
    if (top_depth < 0)
        return rt_fail(RT_ERR_ARG, "top-level " + why);
    if (top_depth > 49)
        return rt_fail(RT_ERR_DEPTH, "top-level BVH deeper than 49");
    int mesh_depth = -1;
    std::vector<uint32_t> mesh_num_nodes(desc->num_meshes, 0);     // nodes of every mesh's face BVH (given, or to be built)
    std::vector<char> mesh_on_device(desc->num_meshes, 0);          // ... to be built on the device
    for (uint32_t m = 0; m < desc->num_meshes; ++m)
    {
        const RtMesh& mesh = desc->meshes[m];
        mesh_on_device[m] = dev_build && mesh.num_nodes == 0 && mesh.num_faces > 0;
        mesh_num_nodes[m] = mesh_on_device[m] ? 2 * mesh.num_faces - 1 : mesh.num_nodes;
        if ((uint64_t)mesh.first_node + mesh.num_nodes > desc->num_mesh_nodes ||
            (uint64_t)mesh.first_face + mesh.num_faces > desc->num_faces ||
            (uint64_t)mesh.first_vertex + mesh.num_vertices > desc->num_vertices ||
            (uint64_t)mesh.first_normal + mesh.num_normals > desc->num_normals ||
            (uint64_t)mesh.first_cdf + mesh.num_faces + 1 > desc->num_cdf)
            return rt_fail(RT_ERR_ARG, "mesh ranges out of bounds");
        if (mesh_on_device[m])
            continue;           // depth: known once the device has built the tree
        if (mesh.num_faces != 0 && mesh.num_nodes == 0)
            return rt_fail(RT_ERR_ARG, "mesh has faces but no BVH nodes: prepare() it on the host, or create the scene with "
                                       "rt_scene_create_ex(..., RT_SCENE_BUILD_MESH_BVH, ...) to have its tree built on the device");
        if (mesh.num_nodes != 0 && mesh.num_nodes != 2 * mesh.num_faces - 1)
            return rt_fail(RT_ERR_ARG, "mesh BVH must have 2*faces-1 nodes");
        int d = bvh_depth(desc->mesh_nodes + mesh.first_node, mesh.num_nodes, mesh.num_faces, why);
        if (d < 0)
            return rt_fail(RT_ERR_ARG, "mesh " + why);
        if (d > 49)
            return rt_fail(RT_ERR_DEPTH, "mesh BVH deeper than 49: the reference's 50-entry traversal stack would overflow");
        if (d > mesh_depth) mesh_depth = d;
    }
    int stack_cap = (desc->num_top_nodes ? top_depth + 1 : (int)desc->num_finite) + (mesh_depth >= 0 ? mesh_depth + 1 : 0);

    hc[1] = std::chrono::steady_clock::now();
    // Fan-expand faces into triangle records
    std::vector<uint32_t> face_first_tri(desc->num_faces + 1, 0);
    for (uint32_t f = 0; f < desc->num_faces; ++f)
    {
        uint32_t n = desc->face_start[f + 1] - desc->face_start[f];
        if (desc->face_start[f + 1] < desc->face_start[f] || n < 3 || desc->face_start[f + 1] > desc->num_indices)
            return rt_fail(RT_ERR_ARG, "face with fewer than 3 vertices or bad face_start");
        if (n - 2 >= (1u << 28))
            return rt_fail(RT_ERR_ARG, "face with too many vertices");
        face_first_tri[f + 1] = face_first_tri[f] + (n - 2);
    }
    const uint32_t num_tris = face_first_tri[desc->num_faces];
    // Device node slots.  The reference numbers a node's two children b and b+1 with b odd
    // (the root is node 0, RAccel.h:366-371), and both are fetched whenever the parent's box
    // is entered.  L2 fills from HBM in aligned 64-byte pieces, so every mesh's nodes are
    // stored one slot up (node i in slot first_node + i with first_node odd): each sibling
    // pair is then exactly one aligned 64-byte piece, and the fetch of the near child brings
    // the far child along.  RAYITO_B200_NODE_ALIGN=0 keeps the pairs straddling (A/B runs).
    const char* align_env = std::getenv("RAYITO_B200_NODE_ALIGN");
    const bool align_pairs = !(align_env != NULL && align_env[0] == '0');
    // Trees that come from the host first (they are part of the copy), trees the device builds behind them
    std::vector<uint32_t> dev_first_node(desc->num_meshes, 0);
    uint64_t dev_nodes = 0, host_tree_nodes = 0;
    for (int pass = 0; pass < 2; ++pass)
    {
        for (uint32_t m = 0; m < desc->num_meshes; ++m)
        {
            if ((mesh_on_device[m] != 0) != (pass == 1))
                continue;
            if (align_pairs && (dev_nodes & 1u) == 0)
                ++dev_nodes;
            dev_first_node[m] = (uint32_t)dev_nodes;
            dev_nodes += mesh_num_nodes[m];
        }
        if (pass == 0)
            host_tree_nodes = dev_nodes;
    }
    // the face-BVH pass packs (child pair index, split axis) and (leaf flag, first triangle record) into one word each
    if (dev_nodes >= (1ull << 29) || num_tris >= (1ull << 31))
        return rt_fail(RT_ERR_UNSUPPORTED, "more than 2^29 mesh BVH nodes or 2^31 fan triangles");
    std::vector<DMesh> meshes(desc->num_meshes);
    for (uint32_t m = 0; m < desc->num_meshes; ++m)
    {
        const RtMesh& mesh = desc->meshes[m];
        DMesh dm;
        dm.first_node = dev_first_node[m];
        dm.num_nodes = mesh_num_nodes[m];
        dm.first_tri = face_first_tri[mesh.first_face];
        dm.first_face = mesh.first_face;
        dm.num_faces = mesh.num_faces;
        dm.first_cdf = mesh.first_cdf;
        dm.total_area = mesh.total_area;
        dm.pad = 0;
        meshes[m] = dm;
    }
    std::vector<DNode> top_nodes(desc->num_top_nodes);
    for (uint32_t i = 0; i < desc->num_top_nodes; ++i)
    {
        const RtBvhNode& n = desc->top_nodes[i];
        float wf, ff;
        std::memcpy(&wf, &n.first_child_or_prim, 4);
        std::memcpy(&ff, &n.flags, 4);
        DNode dn;
        dn.q0 = make_float4(n.bbox_min[0], n.bbox_min[1], n.bbox_min[2], n.bbox_max[0]);
        dn.q1 = make_float4(n.bbox_max[1], n.bbox_max[2], wf, ff);
        top_nodes[i] = dn;
    }

    // Depth-first pop order of the top level per direction octant (DTopStep)
    std::vector<DTopStep> top_walk;
    uint32_t top_walk_steps = 0;
    {
        bool ok = true;
        uint32_t steps = desc->num_top_nodes ? desc->num_top_nodes : desc->num_finite;
        if (steps == 0 || steps > RT_WALK_MAX_STEPS)
            ok = false;
        for (uint32_t oct = 0; oct < 8 && ok; ++oct)
        {
            std::vector<std::pair<uint32_t, uint32_t> > stack;       // (node, depth)
            if (desc->num_top_nodes)
                stack.push_back(std::make_pair(0u, 0u));
            else
                for (uint32_t k = desc->num_finite; k > 0; --k)        // linear list: shapes in order, all "depth 0"
                    stack.push_back(std::make_pair(RT_TOKEN_SHAPE | (k - 1), 0u));
            while (!stack.empty() && ok)
            {
                std::pair<uint32_t, uint32_t> cur = stack.back();
                stack.pop_back();
                DTopStep st;
                std::memset(&st, 0, sizeof(st));
                uint32_t nflags;
                if (cur.first & RT_TOKEN_SHAPE)
                {
                    st.node = RT_WALK_TOKEN;
                    st.word = cur.first & ~RT_TOKEN_SHAPE;
                    nflags = RT_NODE_LEAF;
                }
                else
                {
                    const RtBvhNode& n = desc->top_nodes[cur.first];
                    st.node = cur.first;
                    st.word = n.first_child_or_prim;
                    nflags = n.flags & (RT_NODE_LEAF | RT_NODE_AXIS);
                }
                if (cur.second > RT_WALK_MAX_DEPTH || stack.size() > RT_WALK_MAX_DEPTH)
                {
                    ok = false;
                    break;
                }
                st.flags = nflags | (cur.second << 8) | ((uint32_t)stack.size() << 16);
                for (size_t k = 0; k < stack.size(); ++k)
                {
                    st.pending_node[k] = stack[k].first;
                    st.pending_depth |= stack[k].second << (4 * k);
                }
                top_walk.push_back(st);
                if (!(nflags & RT_NODE_LEAF))
                {
                    if (cur.second + 1 > RT_WALK_MAX_DEPTH)
                    {
                        ok = false;
                        break;
                    }
                    bool neg = (oct >> (nflags & RT_NODE_AXIS)) & 1u;
                    uint32_t near_id = neg ? st.word : st.word + 1, far_id = neg ? st.word + 1 : st.word;
                    stack.push_back(std::make_pair(far_id, cur.second + 1));
                    stack.push_back(std::make_pair(near_id, cur.second + 1));
                }
            }
            if (ok && top_walk.size() != (size_t)(oct + 1) * steps)
                ok = false;
        }
        if (ok)
            top_walk_steps = steps;
        else
            top_walk.clear();
    }

    // Analytic shapes with their per-call constants folded in
    std::vector<DPlane> planes(desc->num_planes);
    for (uint32_t i = 0; i < desc->num_planes; ++i)
    {
        const RtPlane& p = desc->planes[i];
        DPlane d;
        d.px = p.position[0]; d.py = p.position[1]; d.pz = p.position[2];
        d.nx = p.normal[0]; d.ny = p.normal[1]; d.nz = p.normal[2];
        d.pos_dot_n = p.position[0] * p.normal[0] + p.position[1] * p.normal[1] + p.position[2] * p.normal[2];
        d.bullseye = p.bullseye;
        planes[i] = d;
    }
    std::vector<DSphere> spheres(desc->num_spheres);
    for (uint32_t i = 0; i < desc->num_spheres; ++i)
    {
        const RtSphere& s = desc->spheres[i];
        DSphere d = { s.position[0], s.position[1], s.position[2], s.radius };
        spheres[i] = d;
    }
    std::vector<DRect> rects(desc->num_rects);
    for (uint32_t i = 0; i < desc->num_rects; ++i)
    {
        const RtRect& r = desc->rects[i];
        DRect d;
        std::memset(&d, 0, sizeof(d));
        d.px = r.position[0]; d.py = r.position[1]; d.pz = r.position[2];
        // cross(side1, side2).normalized()  (RLight.h:66)
        float n[3] = { r.side1[1] * r.side2[2] - r.side1[2] * r.side2[1],
                       r.side1[2] * r.side2[0] - r.side1[0] * r.side2[2],
                       r.side1[0] * r.side2[1] - r.side1[1] * r.side2[0] };
        float nl = vlen(n);
        if (nl > 0) { n[0] /= nl; n[1] /= nl; n[2] /= nl; }
        d.nx = n[0]; d.ny = n[1]; d.nz = n[2];
        float l1 = vlen(r.side1), l2 = vlen(r.side2);
        d.len1 = l1; d.len2 = l2;
        d.s1x = r.side1[0]; d.s1y = r.side1[1]; d.s1z = r.side1[2];
        d.s2x = r.side2[0]; d.s2y = r.side2[1]; d.s2z = r.side2[2];
        if (l1 > 0) { d.s1x /= l1; d.s1y /= l1; d.s1z /= l1; }
        if (l2 > 0) { d.s2x /= l2; d.s2y /= l2; d.s2z /= l2; }
        d.pos_dot_n = r.position[0] * n[0] + r.position[1] * n[1] + r.position[2] * n[2];
        d.r1x = r.side1[0]; d.r1y = r.side1[1]; d.r1z = r.side1[2];
        d.r2x = r.side2[0]; d.r2y = r.side2[1]; d.r2z = r.side2[2];
        rects[i] = d;
    }
    std::vector<DXform> xforms(desc->num_xforms);
    for (uint32_t i = 0; i < desc->num_xforms; ++i)
    {
        xforms[i].first_key = desc->xforms[i].first_key;
        xforms[i].num_keys = desc->xforms[i].num_keys;
        xforms[i].kind = xf_kind[i];
        xforms[i].pad = 0;
    }

    // One arena, one copy to the device
    ArenaBuilder ab;
    size_t o_shapes = ab.put(shapes.data(), shapes.size() * sizeof(DShapeMem));
    size_t o_top = ab.put(top_nodes.data(), top_nodes.size() * sizeof(DNode));
    size_t o_tris = ab.reserve((size_t)num_tris * 3 * sizeof(float4));
    size_t o_trin = ab.reserve((size_t)num_tris * sizeof(uint4));
    size_t o_fft = ab.put(face_first_tri.data(), face_first_tri.size() * sizeof(uint32_t));
    size_t o_normals = ab.put(desc->normals, (size_t)desc->num_normals * 12);
    size_t o_xforms = ab.put(xforms.data(), xforms.size() * sizeof(DXform));
    size_t o_ktime = ab.put(desc->key_time, (size_t)desc->num_keys * 4);
    size_t o_kscale = ab.put(desc->key_scale, (size_t)desc->num_keys * 12);
    size_t o_krot = ab.put(desc->key_rotation, (size_t)desc->num_keys * 16);
    size_t o_ktrans = ab.put(desc->key_translation, (size_t)desc->num_keys * 12);
    size_t o_planes = ab.put(planes.data(), planes.size() * sizeof(DPlane));
    size_t o_spheres = ab.put(spheres.data(), spheres.size() * sizeof(DSphere));
    size_t o_rects = ab.put(rects.data(), rects.size() * sizeof(DRect));
    size_t o_meshes = ab.put(meshes.data(), meshes.size() * sizeof(DMesh));
    size_t o_cdf = ab.put(desc->face_area_cdf, (size_t)desc->num_cdf * 4);
    size_t o_mats = ab.put(desc->materials, (size_t)desc->num_materials * sizeof(RtMaterial));
    size_t o_lights = ab.put(desc->lights, (size_t)desc->num_lights * 4);
    size_t o_walk = ab.put(top_walk.data(), top_walk.size() * sizeof(DTopStep));
    size_t o_anim = ab.put(anim.data(), anim.size() * sizeof(uint4));
    // The node slots go last: the host's trees are staged in place below and copied with the rest, the slots of
    // the trees the device builds lie behind them and are not part of the copy (nothing of those crosses PCIe)
    size_t o_mnodes = ab.reserve((size_t)host_tree_nodes * sizeof(DNode));
    const size_t copy_bytes = ab.total;
    const size_t arena_total = std::max(ab.total, o_mnodes + (size_t)(dev_nodes ? dev_nodes : 1) * sizeof(DNode));

    hc[2] = std::chrono::steady_clock::now();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        if (host_timing)
            std::fprintf(stderr, "[rayito_b200] rt_scene_create host clock (no device): validate BVHs %.1f ms, small records + layout %.1f\n",
                         std::chrono::duration<double, std::milli>(hc[1] - hc[0]).count(),
                         std::chrono::duration<double, std::milli>(hc[2] - hc[1]).count());
        return rt_fail(RT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev)
        return rt_fail(RT_ERR_ARG, "device ordinal out of range");
    RT_CUDA(cudaSetDevice(device));
    if (const char* fetch_env = std::getenv("RAYITO_B200_L2_FETCH"))    // A/B runs: 32 / 64 / 128 bytes
    {
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::strtol(fetch_env, NULL, 10));
        cudaGetLastError();
    }

    // The staging block is shared by all callers: held until the copy has left it
    hc[3] = std::chrono::steady_clock::now();
    std::lock_guard<std::mutex> stage_guard(stage_lock());
    if (!ab.finish(stage_acquire_locked(ab.total ? ab.total : 16)))
        return rt_fail(RT_ERR_ARG, "out of host memory staging the scene");

    // Fan-triangle records and device nodes of every mesh, written straight into the
    // staging block by the worker threads (faces and nodes are independent of each other)
    {
        float4* tris = ab.at<float4>(o_tris);
        uint4* tri_normals = ab.at<uint4>(o_trin);
        DNode* mesh_nodes = ab.at<DNode>(o_mnodes);
        const uint32_t* fft = face_first_tri.data();
        std::atomic<int> bad(0);
        for (uint32_t m = 0; m < desc->num_meshes; ++m)
        {
            const RtMesh mesh = desc->meshes[m];
            parallel_ranges(mesh.num_faces, 1u << 14, [=, &bad](size_t fb, size_t fe) {
                for (size_t f = fb; f < fe; ++f)
                {
                    uint32_t gf = mesh.first_face + (uint32_t)f;
                    uint32_t start = desc->face_start[gf];
                    uint32_t n = desc->face_start[gf + 1] - start;
                    bool has_n = desc->face_has_normals[gf] != 0;
                    for (uint32_t k = 0; k + 2 < n; ++k)
                    {
                        uint32_t vi[3] = { desc->vertex_index[start], desc->vertex_index[start + k + 1], desc->vertex_index[start + k + 2] };
                        uint32_t rec = fft[gf] + k;
                        uint32_t words[3] = { (uint32_t)f, k, has_n ? 1u : 0u };
                        for (int c = 0; c < 3; ++c)
                        {
                            if (vi[c] >= mesh.num_vertices)
                            {
                                bad.store(1);
                                return;
                            }
                            const float* v = desc->vertices + 3 * (size_t)(mesh.first_vertex + vi[c]);
                            float w;
                            std::memcpy(&w, &words[c], 4);
                            tris[(size_t)rec * 3 + c] = make_float4(v[0], v[1], v[2], w);
                        }
                        uint4 ni = make_uint4(0, 0, 0, 0);
                        if (has_n)
                        {
                            uint32_t a = desc->normal_index[start], b = desc->normal_index[start + k + 1], c = desc->normal_index[start + k + 2];
                            if (a >= mesh.num_normals || b >= mesh.num_normals || c >= mesh.num_normals)
                            {
                                bad.store(2);
                                return;
                            }
                            ni = make_uint4(mesh.first_normal + a, mesh.first_normal + b, mesh.first_normal + c, 0);
                        }
                        tri_normals[rec] = ni;
                    }
                }
            });
            if (mesh_on_device[m])
                continue;
            const uint32_t dev_first = dev_first_node[m];
            if (dev_first != 0)
                std::memset(&mesh_nodes[dev_first - 1], 0, sizeof(DNode));      // the padding slot
            parallel_ranges(mesh.num_nodes, 1u << 15, [=, &bad](size_t nb, size_t ne) {
                for (size_t i = nb; i < ne; ++i)
                {
                    const RtBvhNode& n = desc->mesh_nodes[mesh.first_node + i];
                    uint32_t word = n.first_child_or_prim, flags = n.flags;
                    if (flags & RT_NODE_LEAF)
                    {
                        // (also nodes the walk from the root never reaches: they are converted like the rest)
                        if (n.first_child_or_prim >= mesh.num_faces)
                        {
                            bad.store(3);
                            return;
                        }
                        uint32_t gf = mesh.first_face + n.first_child_or_prim;
                        word = fft[gf];
                        flags = RT_NODE_LEAF | ((fft[gf + 1] - fft[gf]) << 3);
                    }
                    float wf, ff;
                    std::memcpy(&wf, &word, 4);
                    std::memcpy(&ff, &flags, 4);
                    DNode dn;
                    dn.q0 = make_float4(n.bbox_min[0], n.bbox_min[1], n.bbox_min[2], n.bbox_max[0]);
                    dn.q1 = make_float4(n.bbox_max[1], n.bbox_max[2], wf, ff);
                    mesh_nodes[dev_first + i] = dn;
                }
            });
        }
        if (bad.load() != 0)
            return rt_fail(RT_ERR_ARG, bad.load() == 1 ? "vertex index out of range" :
                                       bad.load() == 2 ? "normal index out of range" :
                                                         "mesh BVH leaf (not reachable from the root) names a face that does not exist");
    }

    RtScene* sc = new RtScene();
    sc->device = device;
    sc->arena = NULL;
    sc->arena_bytes = arena_total;
    sc->stack_cap = stack_cap;
    sc->top_stack_need = desc->num_top_nodes ? top_depth + 1 : (int)desc->num_finite;
    sc->mesh_stack_need = mesh_depth >= 0 ? mesh_depth + 1 : 0;
    sc->num_mesh_shapes = 0;
    for (uint32_t i = 0; i < desc->num_finite; ++i)
        if (desc->shapes[i].type == RT_SHAPE_MESH) sc->num_mesh_shapes++;
    sc->num_tris = num_tris;
    sc->num_faces = desc->num_faces;
    sc->light_shapes.assign(desc->lights, desc->lights + desc->num_lights);
    sc->light_types.clear();
    for (uint32_t l = 0; l < desc->num_lights; ++l)
        sc->light_types.push_back(desc->shapes[desc->lights[l]].type);
    sc->shape_types.clear();
    sc->shape_brdfs.clear();
    for (uint32_t i = 0; i < num_shapes; ++i)
    {
        sc->shape_types.push_back(desc->shapes[i].type);
        sc->shape_brdfs.push_back(desc->materials[desc->shapes[i].material].brdf);
    }
    sc->has_lambert = sc->has_glossy = false;
    for (uint32_t m = 0; m < desc->num_materials; ++m)
    {
        if (desc->materials[m].brdf == RT_BRDF_LAMBERT) sc->has_lambert = true;
        if (desc->materials[m].brdf == RT_BRDF_GLOSSY) sc->has_glossy = true;
    }
    sc->scratch_in = sc->scratch_out = NULL;
    sc->scratch_in_bytes = sc->scratch_out_bytes = 0;
    sc->d_work = NULL;
    sc->d_cursor = NULL;
    sc->cursor_next.store(0);
    sc->arena_alloc = sc->work_alloc = sc->cursor_alloc = 0;
    sc->render = NULL;
    sc->dynamic_top = false;

    hc[4] = std::chrono::steady_clock::now();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaError_t err = pool_alloc(device, &sc->arena, sc->arena_bytes, &sc->arena_alloc);
    if (err == cudaSuccess) err = pool_alloc(device, (void**)&sc->d_work, 8 * sizeof(uint64_t), &sc->work_alloc);
    if (err == cudaSuccess) err = cudaMemset(sc->d_work, 0, 8 * sizeof(uint64_t));
    if (err == cudaSuccess) err = pool_alloc(device, (void**)&sc->d_cursor, 64, &sc->cursor_alloc);
    cudaEventRecord(e0, 0);
    if (err == cudaSuccess) err = cudaMemcpy(sc->arena, ab.block, copy_bytes, cudaMemcpyHostToDevice);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    sc->upload_ms = 0.0f;
    cudaEventElapsedTime(&sc->upload_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    hc[5] = std::chrono::steady_clock::now();
    if (host_timing)
    {
        double ms[5];
        for (int i = 0; i < 5; ++i)
            ms[i] = std::chrono::duration<double, std::milli>(hc[i + 1] - hc[i]).count();
        std::fprintf(stderr, "[rayito_b200] rt_scene_create host clock: validate BVHs %.1f ms, small records + layout %.1f, "
                             "device select %.1f, stage triangles/nodes %.1f, alloc + copy %.1f (copy alone %.1f, %.1f MB)\n",
                     ms[0], ms[1], ms[2], ms[3], ms[4], sc->upload_ms, copy_bytes / 1e6);
    }
    if (err != cudaSuccess)
    {
        pool_free(device, sc->arena, sc->arena_alloc);
        pool_free(device, sc->d_work, sc->work_alloc);
        pool_free(device, sc->d_cursor, sc->cursor_alloc);
        delete sc;
        return rt_cuda_fail(err, "scene upload");
    }

    const char* base = static_cast<const char*>(sc->arena);
    DScene& d = sc->d;
    d.set_xform = desc->set_xform;
    d.num_finite = desc->num_finite;
    d.num_infinite = desc->num_infinite;
    d.num_top_nodes = desc->num_top_nodes;
    d.num_lights = desc->num_lights;
    d.stage6 = desc->semantics == RT_SEMANTICS_STAGE6 ? 1u : 0u;
    d.shapes = reinterpret_cast<const DShapeMem*>(base + o_shapes);
    d.top_nodes = reinterpret_cast<const DNode*>(base + o_top);
    d.mesh_nodes = reinterpret_cast<const DNode*>(base + o_mnodes);
    d.tris = reinterpret_cast<const float4*>(base + o_tris);
    d.tri_normals = reinterpret_cast<const uint4*>(base + o_trin);
    d.face_first_tri = reinterpret_cast<const uint32_t*>(base + o_fft);
    d.normals = reinterpret_cast<const float*>(base + o_normals);
    d.xforms = reinterpret_cast<const DXform*>(base + o_xforms);
    d.key_time = reinterpret_cast<const float*>(base + o_ktime);
    d.key_scale = reinterpret_cast<const float*>(base + o_kscale);
    d.key_rot = reinterpret_cast<const float*>(base + o_krot);
    d.key_trans = reinterpret_cast<const float*>(base + o_ktrans);
    d.planes = reinterpret_cast<const DPlane*>(base + o_planes);
    d.spheres = reinterpret_cast<const DSphere*>(base + o_spheres);
    d.rects = reinterpret_cast<const DRect*>(base + o_rects);
    d.meshes = reinterpret_cast<const DMesh*>(base + o_meshes);
    d.face_area_cdf = reinterpret_cast<const float*>(base + o_cdf);
    d.materials = reinterpret_cast<const RtMaterial*>(base + o_mats);
    d.lights = reinterpret_cast<const uint32_t*>(base + o_lights);
    d.top_walk = top_walk_steps ? reinterpret_cast<const DTopStep*>(base + o_walk) : NULL;
    d.top_walk_steps = top_walk_steps;
    d.top_walk_levels = (uint32_t)(desc->num_top_nodes ? top_depth : 0) + 2u;
    d.anim = reinterpret_cast<const uint4*>(base + o_anim);
    d.num_anim = (uint32_t)anim.size();
    d.anim_stride = anim_stride;

    sc->mesh_first_node = dev_first_node;
    sc->mesh_first_face.resize(desc->num_meshes);
    sc->mesh_faces.resize(desc->num_meshes);
    for (uint32_t m = 0; m < desc->num_meshes; ++m)
    {
        sc->mesh_first_face[m] = desc->meshes[m].first_face;
        sc->mesh_faces[m] = desc->meshes[m].num_faces;
    }
    sc->bvh_build_ms = 0.0f;
    sc->mesh_depth = mesh_depth;
    if (dev_build)
    {
        // Bvh<Mesh>::build for every mesh, on the device, out of the triangle records just uploaded
        DNode* nodes = reinterpret_cast<DNode*>(static_cast<char*>(sc->arena) + o_mnodes);
        int rc = RT_OK;
        for (uint32_t m = 0; m < desc->num_meshes && rc == RT_OK; ++m)
        {
            if (!mesh_on_device[m])
                continue;
            if (dev_first_node[m] != 0)
                cudaMemsetAsync(nodes + dev_first_node[m] - 1, 0, sizeof(DNode), 0);      // the padding slot
            int depth = 0;
            float ms = 0.0f;
            rc = rt_build::build_mesh(device, nodes + dev_first_node[m], d.tris, d.face_first_tri, desc->meshes[m].first_face,
                                      desc->meshes[m].num_faces, &depth, &ms);
            sc->bvh_build_ms += ms;
            if (rc == RT_OK && depth > 49)
                rc = rt_fail(RT_ERR_DEPTH, "mesh BVH deeper than 49: the reference's 50-entry traversal stack would overflow");
            if (depth > mesh_depth) mesh_depth = depth;
        }
        if (rc != RT_OK)
        {
            pool_free(device, sc->arena, sc->arena_alloc);
            pool_free(device, sc->d_work, sc->work_alloc);
            pool_free(device, sc->d_cursor, sc->cursor_alloc);
            delete sc;
            return rc;
        }
        sc->mesh_depth = mesh_depth;
        sc->stack_cap = (desc->num_top_nodes ? top_depth + 1 : (int)desc->num_finite) + (mesh_depth >= 0 ? mesh_depth + 1 : 0);
        sc->mesh_stack_need = mesh_depth >= 0 ? mesh_depth + 1 : 0;
        if (host_timing)
            std::fprintf(stderr, "[rayito_b200] rt_scene_create: face BVHs built on the device in %.2f ms (deepest leaf %d)\n",
                         sc->bvh_build_ms, mesh_depth);
    }

    *out_scene = sc;
    return RT_OK;
}

#endif // RAYITO_B200_RT_SCENE_CUH
