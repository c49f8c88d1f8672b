// Face BVH built ON THE DEVICE, node for node the tree the reference builds.
//
// Bvh<T>::build / buildRange (Rayito_Stage7_QT/RAccel.h:262-374) is a serial recursion: a node splits the
// longest axis of its box at the midpoint, std::partition moves the elements whose box centre lies above
// the cut to the front, an empty side cuts the range in half instead, the children get the unions of their
// elements' boxes, slots are handed out in recursion order.  Hit records depend on that exact tree (the slab
// test is not watertight and ties in t keep the first face found), so a device build has to reproduce
//   (1) the ELEMENT ORDER std::partition leaves behind -- it decides what "cut in half" means further down,
//   (2) the slot numbering of the recursion, and
//   (3) the boxes, down to the sign of a zero (std::min / std::max keep their first argument on a tie).
// All three have closed forms, which is what makes the build data-parallel:
//   (1) libstdc++'s partition for bidirectional iterators walks one cursor up past elements that satisfy the
//       predicate and one down past elements that do not, swaps the pair it stops at, and repeats until the
//       cursors meet.  With m elements satisfying the predicate they meet at m: exactly the "false" elements
//       in [0, m) and the "true" ones in [m, n) move, and the k-th false from the left changes places with
//       the k-th true from the right.  One prefix sum of the predicate gives every mover its k.
//   (2) One element per leaf: a subtree over n elements has 2n-1 nodes.  A node whose descendants start at
//       slot `base` and whose left side holds nL elements puts its children at base, base+1, the left
//       child's descendants from base+2 and the right child's from base+2*nL.
//   (3) An in-order min with "keep the first on a tie" is the minimum of the keys (value with -0 == +0,
//       position): one 64-bit atomicMin; the sign of a winning zero rides in the key's lowest bit.
// Two phases.  Ranges of more than RT_BUILD_SMALL elements are split LEVEL BY LEVEL, every level a handful of
// element-parallel kernels over the whole item array (predicate, one scan, mover lists, swaps, child boxes,
// child ranges); a range of at most RT_BUILD_SMALL elements is finished by ONE thread that runs the
// reference's recursion literally (explicit stack, the serial partition, in-order box unions).  Nodes are
// written straight into the scene arena in the traversal kernels' layout (leaves re-encoded to their
// triangle records), so nothing of the tree ever crosses PCIe.
// 5 M quads (config C5): see profiles/README.md for the measured time against the 16-core host build.
#ifndef RAYITO_B200_RT_BUILD_CUH
#define RAYITO_B200_RT_BUILD_CUH

#include <cub/device/device_scan.cuh>

#include "rt_scene.cuh"

#ifndef RT_BUILD_SMALL
#define RT_BUILD_SMALL 32u
#endif
#define RT_BUILD_NONE 0xffffffffu

namespace rt_build
{

// One build element: the reference's BuildElement (RAccel.h:216-220), in the two-float4 shape of a node
//   a = (min x, min y, min z, max x)   b = (max y, max z, prim, -)
struct Item
{
    float4 a;
    float4 b;
};

// A range of the item array that still has to become a subtree
struct Seg
{
    uint32_t begin, end;       // element range
    uint32_t node;             // slot of this subtree's root (mesh-local, reference numbering)
    uint32_t base;             // first slot of its descendants
    uint32_t depth;
    uint32_t axis;
    float where;
    uint32_t mid;              // first element of the right child after the partition
    uint32_t pairs;            // number of swaps the partition performs
    uint32_t child[2];         // next level's ids of the two children (NONE: handed to the small list)
    float lo[3], hi[3];        // node box
};

struct Small
{
    uint32_t begin, end, node, base, depth;
    float lo[3], hi[3];
};

__device__ __forceinline__ float comp(float x, float y, float z, uint32_t axis) { return axis == 0 ? x : (axis == 1 ? y : z); }

// BuildElementPredicate (RAccel.h:226-240): splitAxis < (max + min) * 0.5f on the split axis
__device__ __forceinline__ bool above_split(const Item& it, uint32_t axis, float where)
{
    float mn = comp(it.a.x, it.a.y, it.a.z, axis);
    float mx = comp(it.a.w, it.b.x, it.b.y, axis);
    return where < (mx + mn) * 0.5f;
}

// buildRange's choice of axis and cut (RAccel.h:305-326)
__device__ __forceinline__ void plan_split(const float* lo, const float* hi, uint32_t& axis, float& where)
{
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    if (ex > ey)
        axis = ex > ez ? 0u : 2u;
    else
        axis = ey > ez ? 1u : 2u;
    where = (hi[axis] + lo[axis]) * 0.5f;
}

// ---- ordered keys: in-order std::min / std::max with their first-argument-wins ties --------------------
__device__ __forceinline__ uint32_t ordered_bits(float f)
{
    if (f == 0.0f) f = 0.0f;        // -0 and +0 compare equal in std::min / std::max
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_value(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ uint32_t neg_zero(float f) { return __float_as_uint(f) == 0x80000000u ? 1u : 0u; }
// smallest value, earliest position on ties
__device__ __forceinline__ unsigned long long min_key(float f, uint32_t pos)
{
    return ((unsigned long long)ordered_bits(f) << 32) | ((unsigned long long)pos << 1) | neg_zero(f);
}
// largest value, earliest position on ties
__device__ __forceinline__ unsigned long long max_key(float f, uint32_t pos)
{
    return ((unsigned long long)ordered_bits(f) << 32) | ((unsigned long long)(0x7fffffffu - pos) << 1) | neg_zero(f);
}
__device__ __forceinline__ float key_value(unsigned long long key)
{
    float v = ordered_value((uint32_t)(key >> 32));
    return (v == 0.0f && (key & 1ull)) ? -0.0f : v;
}
#define RT_BUILD_MIN_IDENTITY 0xffffffffffffffffull
#define RT_BUILD_MAX_IDENTITY 0ull

// Six keys of one item's box, reduced over the warp when every lane feeds the same destination
__device__ __forceinline__ void box_keys_reduce(const Item& it, uint32_t pos, bool valid, uint32_t dest, unsigned long long* keys /* [dest][6] */)
{
    unsigned long long k[6];
    k[0] = valid ? min_key(it.a.x, pos) : RT_BUILD_MIN_IDENTITY;
    k[1] = valid ? min_key(it.a.y, pos) : RT_BUILD_MIN_IDENTITY;
    k[2] = valid ? min_key(it.a.z, pos) : RT_BUILD_MIN_IDENTITY;
    k[3] = valid ? max_key(it.a.w, pos) : RT_BUILD_MAX_IDENTITY;
    k[4] = valid ? max_key(it.b.x, pos) : RT_BUILD_MAX_IDENTITY;
    k[5] = valid ? max_key(it.b.y, pos) : RT_BUILD_MAX_IDENTITY;
    const uint32_t live = __ballot_sync(0xffffffffu, valid);
    if (live == 0)
        return;
    const uint32_t leader = __ffs(live) - 1;
    const uint32_t dest0 = __shfl_sync(0xffffffffu, dest, leader);
    const bool uniform = __all_sync(0xffffffffu, !valid || dest == dest0);
    if (uniform)
    {
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1)
        {
            #pragma unroll
            for (int c = 0; c < 6; ++c)
            {
                unsigned long long o = __shfl_xor_sync(0xffffffffu, k[c], off);
                k[c] = c < 3 ? (o < k[c] ? o : k[c]) : (o > k[c] ? o : k[c]);
            }
        }
        if ((threadIdx.x & 31u) == 0)
        {
            #pragma unroll
            for (int c = 0; c < 3; ++c) atomicMin(keys + (size_t)dest0 * 6 + c, k[c]);
            #pragma unroll
            for (int c = 3; c < 6; ++c) atomicMax(keys + (size_t)dest0 * 6 + c, k[c]);
        }
    }
    else if (valid)
    {
        #pragma unroll
        for (int c = 0; c < 3; ++c) atomicMin(keys + (size_t)dest * 6 + c, k[c]);
        #pragma unroll
        for (int c = 3; c < 6; ++c) atomicMax(keys + (size_t)dest * 6 + c, k[c]);
    }
}

struct Ctx
{
    Item* items;               // [n]
    uint32_t* seg_of;          // [n] id of the range an element is in (this level), NONE once it left the level phase
    uint32_t* pred;            // [n + 1] predicate (0 / 1), last entry 0
    uint32_t* scan;            // [n + 1] exclusive prefix sum of pred
    uint32_t* left;            // [n] mover lists, a range's entries inside the range's own span
    uint32_t* right;           // [n]
    Seg* cur;                  // this level's ranges
    Seg* next;                 // next level's
    unsigned long long* keys;  // [2 * max ranges][6] child-box keys (entry 0 doubles as the root's)
    Small* small;              // ranges for the one-thread phase
    uint32_t* counters;        // [0] next level's count, [1] small count, [2] deepest leaf, [3] small-list overflow
    uint32_t small_cap;
    uint32_t n;
    // where the tree goes
    DNode* nodes;              // the mesh's first device node slot
    const float4* tris;        // fan-triangle records (xyz + id word), three per triangle
    const uint32_t* fft;       // global face -> first triangle record
    uint32_t first_face;       // mesh's first global face
};

// Leaf in the traversal kernels' encoding (rt_scene.cuh): word = first triangle record, flags = LEAF | count << 3
__device__ __forceinline__ void write_leaf(const Ctx& c, uint32_t node, const float* lo, const float* hi, uint32_t prim)
{
    const uint32_t gf = c.first_face + prim;
    const uint32_t first = c.fft[gf], count = c.fft[gf + 1] - first;
    DNode dn;
    dn.q0 = make_float4(lo[0], lo[1], lo[2], hi[0]);
    dn.q1 = make_float4(hi[1], hi[2], __uint_as_float(first), __uint_as_float(RT_NODE_LEAF | (count << 3)));
    c.nodes[node] = dn;
}
__device__ __forceinline__ void write_interior(const Ctx& c, uint32_t node, const float* lo, const float* hi, uint32_t axis, uint32_t first_child)
{
    DNode dn;
    dn.q0 = make_float4(lo[0], lo[1], lo[2], hi[0]);
    dn.q1 = make_float4(hi[1], hi[2], __uint_as_float(first_child), __uint_as_float(axis));
    c.nodes[node] = dn;
}

// Element boxes (Mesh::elementBBox, RMesh.h:226-236: expand over the face's vertices in order) and the keys of
// their union in element order (Bvh::build, RAccel.h:273-279)
__global__ void __launch_bounds__(256) k_items(const Ctx c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < c.n;
    Item it;
    it.a = it.b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (valid)
    {
        const uint32_t gf = c.first_face + i;
        const uint32_t first = c.fft[gf], count = c.fft[gf + 1] - first;
        float lo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
        float hi[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
        // fan triangle k of a face is (v0, v[k+1], v[k+2]): the face's vertices in order are the three of
        // triangle 0 and the last one of every further triangle
        for (uint32_t k = 0; k < count; ++k)
            for (uint32_t v = (k == 0 ? 0u : 2u); v < 3u; ++v)
            {
                const float4 p = c.tris[((size_t)first + k) * 3 + v];
                // BBox::expand: m_min = min(m_min, p), std::min keeps its first argument on ties
                lo[0] = (p.x < lo[0]) ? p.x : lo[0];  lo[1] = (p.y < lo[1]) ? p.y : lo[1];  lo[2] = (p.z < lo[2]) ? p.z : lo[2];
                hi[0] = (hi[0] < p.x) ? p.x : hi[0];  hi[1] = (hi[1] < p.y) ? p.y : hi[1];  hi[2] = (hi[2] < p.z) ? p.z : hi[2];
            }
        it.a = make_float4(lo[0], lo[1], lo[2], hi[0]);
        it.b = make_float4(hi[1], hi[2], __uint_as_float(i), 0.0f);
        c.items[i] = it;
        c.seg_of[i] = c.n > RT_BUILD_SMALL ? 0u : RT_BUILD_NONE;
    }
    box_keys_reduce(it, i, valid, 0u, c.keys);
}

__global__ void k_root(const Ctx c)
{
    float lo[3], hi[3];
    for (int k = 0; k < 3; ++k)
    {
        lo[k] = key_value(c.keys[k]);
        hi[k] = key_value(c.keys[3 + k]);
    }
    if (c.n > RT_BUILD_SMALL)
    {
        Seg s;
        s.begin = 0; s.end = c.n; s.node = 0; s.base = 1; s.depth = 0;
        s.axis = 0; s.where = 0.0f; s.mid = 0; s.pairs = 0; s.child[0] = s.child[1] = RT_BUILD_NONE;
        for (int k = 0; k < 3; ++k) { s.lo[k] = lo[k]; s.hi[k] = hi[k]; }
        c.cur[0] = s;
        c.counters[1] = 0;
    }
    else
    {
        Small s;
        s.begin = 0; s.end = c.n; s.node = 0; s.base = 1; s.depth = 0;
        for (int k = 0; k < 3; ++k) { s.lo[k] = lo[k]; s.hi[k] = hi[k]; }
        c.small[0] = s;
        c.counters[1] = 1;
    }
    c.counters[0] = 0;
    c.counters[2] = 0;
    c.counters[3] = 0;
    c.pred[c.n] = 0;
}

// ---- one level ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_plan(const Ctx c, uint32_t nseg)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg)
        return;
    Seg& g = c.cur[s];
    uint32_t axis;
    float where;
    plan_split(g.lo, g.hi, axis, where);
    g.axis = axis;
    g.where = where;
    write_interior(c, g.node, g.lo, g.hi, axis, g.base);
    for (int k = 0; k < 12; ++k)
        c.keys[(size_t)s * 12 + k] = (k % 6) < 3 ? RT_BUILD_MIN_IDENTITY : RT_BUILD_MAX_IDENTITY;
}

__global__ void __launch_bounds__(256) k_pred(const Ctx c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n)
        return;
    const uint32_t s = c.seg_of[i];
    uint32_t p = 0;
    if (s != RT_BUILD_NONE)
    {
        const Seg& g = c.cur[s];
        p = above_split(c.items[i], g.axis, g.where) ? 1u : 0u;
    }
    c.pred[i] = p;
}

__global__ void __launch_bounds__(128) k_mid(const Ctx c, uint32_t nseg)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg)
        return;
    Seg& g = c.cur[s];
    const uint32_t n = g.end - g.begin;
    const uint32_t trues = c.scan[g.end] - c.scan[g.begin];
    if (trues == 0 || trues == n)
    {
        // one side empty: std::partition moved nothing; cut the range in half (RAccel.h:345-352; n >= 2)
        g.mid = g.begin + n / 2;
        g.pairs = 0;
    }
    else
    {
        g.mid = g.begin + trues;
        g.pairs = trues - (c.scan[g.mid] - c.scan[g.begin]);      // falses among the first `trues` elements
    }
}

__global__ void __launch_bounds__(256) k_lists(const Ctx c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n)
        return;
    const uint32_t s = c.seg_of[i];
    if (s == RT_BUILD_NONE)
        return;
    const Seg& g = c.cur[s];
    if (g.pairs == 0)
        return;
    const bool p = c.pred[i] != 0;
    if (i < g.mid)
    {
        if (!p)     // the k-th false from the left, k = falses before it
            c.left[g.begin + ((i - g.begin) - (c.scan[i] - c.scan[g.begin]))] = i;
    }
    else if (p)     // the k-th true from the right, k = trues behind it
        c.right[g.begin + (c.scan[g.end] - c.scan[i + 1])] = i;
}

__global__ void __launch_bounds__(256) k_swap(const Ctx c)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= c.n)
        return;
    const uint32_t s = c.seg_of[p];
    if (s == RT_BUILD_NONE)
        return;
    const Seg& g = c.cur[s];
    if (p - g.begin >= g.pairs)
        return;
    const uint32_t i = c.left[p], j = c.right[p];
    const Item x = c.items[i], y = c.items[j];
    c.items[i] = y;
    c.items[j] = x;
}

__global__ void __launch_bounds__(256) k_boxes(const Ctx c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < c.n;
    uint32_t s = valid ? c.seg_of[i] : RT_BUILD_NONE;
    valid = valid && s != RT_BUILD_NONE;
    Item it;
    it.a = it.b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    uint32_t dest = 0;
    if (valid)
    {
        it = c.items[i];
        dest = 2 * s + (i >= c.cur[s].mid ? 1u : 0u);
    }
    box_keys_reduce(it, i, valid, dest, c.keys);
}

__global__ void __launch_bounds__(128) k_children(const Ctx c, uint32_t nseg)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg)
        return;
    Seg& g = c.cur[s];
    for (uint32_t side = 0; side < 2; ++side)
    {
        const uint32_t b = side == 0 ? g.begin : g.mid, e = side == 0 ? g.mid : g.end;
        const unsigned long long* k = c.keys + ((size_t)2 * s + side) * 6;
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a)
        {
            lo[a] = key_value(k[a]);
            hi[a] = key_value(k[3 + a]);
        }
        const uint32_t node = g.base + side;
        const uint32_t base = side == 0 ? g.base + 2 : g.base + 2 * (g.mid - g.begin);
        if (e - b > RT_BUILD_SMALL)
        {
            const uint32_t id = atomicAdd(c.counters + 0, 1u);
            Seg n;
            n.begin = b; n.end = e; n.node = node; n.base = base; n.depth = g.depth + 1;
            n.axis = 0; n.where = 0.0f; n.mid = 0; n.pairs = 0; n.child[0] = n.child[1] = RT_BUILD_NONE;
            for (int a = 0; a < 3; ++a) { n.lo[a] = lo[a]; n.hi[a] = hi[a]; }
            c.next[id] = n;
            g.child[side] = id;
        }
        else
        {
            const uint32_t id = atomicAdd(c.counters + 1, 1u);
            if (id < c.small_cap)
            {
                Small n;
                n.begin = b; n.end = e; n.node = node; n.base = base; n.depth = g.depth + 1;
                for (int a = 0; a < 3; ++a) { n.lo[a] = lo[a]; n.hi[a] = hi[a]; }
                c.small[id] = n;
            }
            else
                c.counters[3] = 1;
            g.child[side] = RT_BUILD_NONE;
        }
    }
}

__global__ void __launch_bounds__(256) k_relabel(const Ctx c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n)
        return;
    const uint32_t s = c.seg_of[i];
    if (s == RT_BUILD_NONE)
        return;
    const Seg& g = c.cur[s];
    c.seg_of[i] = g.child[i >= g.mid ? 1 : 0];
}

// ---- small ranges: the reference's recursion, literally, one thread per range ---------------------------
__global__ void __launch_bounds__(64) k_small(const Ctx c, uint32_t count)
{
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count)
        return;
    Small stack[RT_BUILD_SMALL + 2];
    int sp = 0;
    stack[sp++] = c.small[id];
    uint32_t deepest = 0;
    while (sp > 0)
    {
        const Small job = stack[--sp];
        if (job.depth > deepest)
            deepest = job.depth;
        if (job.end - job.begin <= 1)
        {
            write_leaf(c, job.node, job.lo, job.hi, __float_as_uint(c.items[job.begin].b.z));
            continue;
        }
        uint32_t axis;
        float where;
        plan_split(job.lo, job.hi, axis, where);
        write_interior(c, job.node, job.lo, job.hi, axis, job.base);
        // std::partition, bidirectional form (libstdc++ stl_algo.h __partition)
        uint32_t first = job.begin, last = job.end;
        for (;;)
        {
            bool done = false;
            for (;;)
            {
                if (first == last) { done = true; break; }
                if (above_split(c.items[first], axis, where)) ++first; else break;
            }
            if (done) break;
            --last;
            for (;;)
            {
                if (first == last) { done = true; break; }
                if (!above_split(c.items[last], axis, where)) --last; else break;
            }
            if (done) break;
            const Item x = c.items[first], y = c.items[last];
            c.items[first] = y;
            c.items[last] = x;
            ++first;
        }
        uint32_t mid = first;
        if (mid <= job.begin || mid >= job.end)
        {
            mid = job.begin + (job.end - job.begin) / 2;
            if (mid < job.begin + 1) mid = job.begin + 1;
            else if (mid > job.end - 1) mid = job.end - 1;
        }
        Small l, r;
        for (int side = 0; side < 2; ++side)
        {
            Small& ch = side == 0 ? l : r;
            const uint32_t b = side == 0 ? job.begin : mid, e = side == 0 ? mid : job.end;
            float lo[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
            float hi[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
            for (uint32_t i = b; i < e; ++i)
            {
                const Item it = c.items[i];
                // BBox::combined: min(a.m_min, b.m_min), max(a.m_max, b.m_max) with std::min / std::max
                lo[0] = (it.a.x < lo[0]) ? it.a.x : lo[0];  lo[1] = (it.a.y < lo[1]) ? it.a.y : lo[1];  lo[2] = (it.a.z < lo[2]) ? it.a.z : lo[2];
                hi[0] = (hi[0] < it.a.w) ? it.a.w : hi[0];  hi[1] = (hi[1] < it.b.x) ? it.b.x : hi[1];  hi[2] = (hi[2] < it.b.y) ? it.b.y : hi[2];
            }
            ch.begin = b; ch.end = e; ch.node = job.base + side; ch.depth = job.depth + 1;
            ch.base = side == 0 ? job.base + 2 : job.base + 2 * (mid - job.begin);
            for (int a = 0; a < 3; ++a) { ch.lo[a] = lo[a]; ch.hi[a] = hi[a]; }
        }
        stack[sp++] = r;
        stack[sp++] = l;
    }
    atomicMax(c.counters + 2, deepest);
}

struct Buffers
{
    void* block;
    size_t bytes;
    Buffers() : block(NULL), bytes(0) { }
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Builds the face BVH of one mesh into `nodes` (2 * num_faces - 1 slots).  *depth: depth of the deepest leaf.
// *build_ms: device time.  Stream 0; returns after the device is done.
inline int build_mesh(int device, DNode* nodes, const float4* tris, const uint32_t* fft, uint32_t first_face, uint32_t num_faces,
                      int* depth, float* build_ms)
{
    *depth = 0;
    if (build_ms) *build_ms = 0.0f;
    if (num_faces == 0)
        return RT_OK;
    const uint32_t n = num_faces;
    const uint32_t max_segs = n / (RT_BUILD_SMALL + 1) + 2;
    const uint32_t small_cap = n / 4 + 1024;
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(NULL, scan_bytes, (const uint32_t*)NULL, (uint32_t*)NULL, (int)(n + 1));
    size_t off = 0;
    const size_t o_items = off;   off = align256(off + (size_t)n * sizeof(Item));
    const size_t o_segof = off;   off = align256(off + (size_t)n * 4);
    const size_t o_pred = off;    off = align256(off + (size_t)(n + 1) * 4);
    const size_t o_scan = off;    off = align256(off + (size_t)(n + 1) * 4);
    const size_t o_left = off;    off = align256(off + (size_t)n * 4);
    const size_t o_right = off;   off = align256(off + (size_t)n * 4);
    const size_t o_cur = off;     off = align256(off + (size_t)max_segs * sizeof(Seg));
    const size_t o_next = off;    off = align256(off + (size_t)max_segs * sizeof(Seg));
    const size_t o_keys = off;    off = align256(off + (size_t)max_segs * 12 * sizeof(unsigned long long));
    const size_t o_small = off;   off = align256(off + (size_t)small_cap * sizeof(Small));
    const size_t o_count = off;   off = align256(off + 16 * 4);
    const size_t o_temp = off;    off = align256(off + scan_bytes);
    void* block = NULL;
    size_t got = 0;
    RT_CUDA(rt_detail::pool_alloc(device, &block, off, &got));
    char* base = static_cast<char*>(block);
    Ctx c;
    c.items = reinterpret_cast<Item*>(base + o_items);
    c.seg_of = reinterpret_cast<uint32_t*>(base + o_segof);
    c.pred = reinterpret_cast<uint32_t*>(base + o_pred);
    c.scan = reinterpret_cast<uint32_t*>(base + o_scan);
    c.left = reinterpret_cast<uint32_t*>(base + o_left);
    c.right = reinterpret_cast<uint32_t*>(base + o_right);
    c.cur = reinterpret_cast<Seg*>(base + o_cur);
    c.next = reinterpret_cast<Seg*>(base + o_next);
    c.keys = reinterpret_cast<unsigned long long*>(base + o_keys);
    c.small = reinterpret_cast<Small*>(base + o_small);
    c.counters = reinterpret_cast<uint32_t*>(base + o_count);
    c.small_cap = small_cap;
    c.n = n;
    c.nodes = nodes;
    c.tris = tris;
    c.fft = fft;
    c.first_face = first_face;
    void* scan_temp = base + o_temp;

    int rc = RT_OK;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    const unsigned eb = (n + 255) / 256;
    // root keys live in keys[0..5]
    {
        unsigned long long init[6] = { RT_BUILD_MIN_IDENTITY, RT_BUILD_MIN_IDENTITY, RT_BUILD_MIN_IDENTITY,
                                       RT_BUILD_MAX_IDENTITY, RT_BUILD_MAX_IDENTITY, RT_BUILD_MAX_IDENTITY };
        cudaMemcpyAsync(c.keys, init, sizeof(init), cudaMemcpyHostToDevice, 0);
    }
    k_items<<<eb, 256>>>(c);
    k_root<<<1, 1>>>(c);
    uint32_t nseg = n > RT_BUILD_SMALL ? 1u : 0u;
    uint32_t host_counters[4] = { 0, 0, 0, 0 };
    cudaError_t err = cudaGetLastError();
    int levels = 0;
    while (err == cudaSuccess && nseg > 0)
    {
        if (nseg > max_segs || ++levels > 4096)
        {
            rc = rt_fail(RT_ERR_UNSUPPORTED, "device BVH build: range list overflow");
            break;
        }
        const unsigned sb = (nseg + 127) / 128;
        k_plan<<<sb, 128>>>(c, nseg);
        k_pred<<<eb, 256>>>(c);
        cub::DeviceScan::ExclusiveSum(scan_temp, scan_bytes, c.pred, c.scan, (int)(n + 1), 0);
        k_mid<<<sb, 128>>>(c, nseg);
        k_lists<<<eb, 256>>>(c);
        k_swap<<<eb, 256>>>(c);
        k_boxes<<<eb, 256>>>(c);
        k_children<<<sb, 128>>>(c, nseg);
        k_relabel<<<eb, 256>>>(c);
        err = cudaMemcpy(host_counters, c.counters, sizeof(host_counters), cudaMemcpyDeviceToHost);
        if (err != cudaSuccess)
            break;
        nseg = host_counters[0];
        cudaMemsetAsync(c.counters, 0, 4, 0);
        Seg* t = c.cur; c.cur = c.next; c.next = t;
    }
    if (err == cudaSuccess && rc == RT_OK)
    {
        err = cudaMemcpy(host_counters, c.counters, sizeof(host_counters), cudaMemcpyDeviceToHost);
        if (err == cudaSuccess && host_counters[3] != 0)
            rc = rt_fail(RT_ERR_UNSUPPORTED, "device BVH build: too many small ranges (degenerate mesh); build on the host");
        if (err == cudaSuccess && rc == RT_OK && host_counters[1] > 0)
        {
            k_small<<<(host_counters[1] + 63) / 64, 64>>>(c, host_counters[1]);
            err = cudaGetLastError();
        }
        if (err == cudaSuccess && rc == RT_OK)
            err = cudaMemcpy(host_counters, c.counters, sizeof(host_counters), cudaMemcpyDeviceToHost);
    }
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    if (build_ms) cudaEventElapsedTime(build_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    rt_detail::pool_free(device, block, got);
    if (err != cudaSuccess)
        return rt_cuda_fail(err, "device BVH build");
    if (rc != RT_OK)
        return rc;
    *depth = (int)host_counters[2];
    return RT_OK;
}

} // namespace rt_build

#endif // RAYITO_B200_RT_BUILD_CUH
