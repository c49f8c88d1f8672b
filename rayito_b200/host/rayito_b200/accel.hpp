// Ray / Intersection records, bounding boxes and the BVH *builder* of the host API.
//
// Traversal lives on the GPU (rayito_b200/csrc/rt_trace.cu).  The builder here
// must produce exactly the tree the reference builds -- same split rule, same
// std::partition from the same libstdc++, same child numbering -- because the
// reference's non-watertight slab test makes hit results depend on the tree
// (SURVEY.md section 7 "hard parts").  Reference: Rayito_Stage7_QT/RAccel.h.
#ifndef RAYITO_B200_ACCEL_HPP
#define RAYITO_B200_ACCEL_HPP

#include <algorithm>
#include <cstdlib>
#include <limits>
#include <new>
#include <mutex>
#include <vector>

#include "math.hpp"
#include "parallel.hpp"

namespace Rayito
{

// RRay.h:23,28
const float kRayTMin = 0.0001f;
const float kRayTMax = 1.0e30f;

// RRay.h:31-88; 32 bytes, identical to RtRay of the C ABI
struct Ray
{
    Point m_origin;
    Vector m_direction;
    float m_tMax;
    float m_time;

    Ray() : m_origin(), m_direction(0.0f, 0.0f, 1.0f), m_tMax(kRayTMax), m_time(0.0f) { }
    Ray(const Point& origin, const Vector& direction, float tMax = kRayTMax, float time = 0.0f)
        : m_origin(origin), m_direction(direction), m_tMax(tMax), m_time(time) { }

    Point calculate(float t) const { return m_origin + t * m_direction; }

    Ray transformToLocal(const Transform& x) const
    {
        return Ray(x.toLocalPoint(m_time, m_origin), x.toLocalVector(m_time, m_direction), m_tMax, m_time);
    }
    Ray transformFromLocal(const Transform& x) const
    {
        return Ray(x.fromLocalPoint(m_time, m_origin), x.fromLocalVector(m_time, m_direction), m_tMax, m_time);
    }
};

class Shape;
class Material;

// RRay.h:98-155
struct Intersection
{
    Ray m_ray;
    float m_t;
    Shape* m_pShape;
    Material* m_pMaterial;
    Color m_colorModifier;
    Vector m_normal;

    Intersection() : m_ray(), m_t(kRayTMax), m_pShape(NULL), m_pMaterial(NULL), m_colorModifier(1.0f, 1.0f, 1.0f), m_normal() { }
    Intersection(const Ray& ray)
        : m_ray(ray), m_t(ray.m_tMax), m_pShape(NULL), m_pMaterial(NULL), m_colorModifier(1.0f, 1.0f, 1.0f), m_normal() { }

    bool intersected() const { return m_pShape != NULL; }
    Point position() const { return m_ray.calculate(m_t); }
};


// RAccel.h:22-115 (build-time subset; the slab test is device code)
struct BBox
{
    Point m_min, m_max;

    BBox() : m_min(std::numeric_limits<float>::max()), m_max(-std::numeric_limits<float>::max()) { }
    BBox(const Point& lo, const Point& hi) : m_min(lo), m_max(hi) { }

    bool valid() const { return m_min.m_x < m_max.m_x && m_min.m_y < m_max.m_y && m_min.m_z < m_max.m_z; }
    bool empty() const { return !valid(); }

    BBox combined(const BBox& b) const { return BBox(min(m_min, b.m_min), max(m_max, b.m_max)); }
    void expand(const Point& p) { m_min = min(m_min, p); m_max = max(m_max, p); }
    BBox intersection(const BBox& b) const { return BBox(max(m_min, b.m_min), min(m_max, b.m_max)); }
    bool overlaps(const BBox& b) const { return intersection(b).valid(); }
    bool contains(const Point& p) const
    {
        return m_min.m_x <= p.m_x && m_max.m_x >= p.m_x &&
               m_min.m_y <= p.m_y && m_max.m_y >= p.m_y &&
               m_min.m_z <= p.m_z && m_max.m_z >= p.m_z;
    }

    // Box around the eight transformed corners, visited in the reference's corner
    // order (RAccel.h:93-114)
    BBox transformFromLocal(float time, const Transform& x) const
    {
        BBox out;
        for (int c = 0; c < 8; ++c)
        {
            Point corner((c & 4) ? m_max.m_x : m_min.m_x,
                         (c & 2) ? m_max.m_y : m_min.m_y,
                         (c & 1) ? m_max.m_z : m_min.m_z);
            out.expand(x.fromLocalPoint(time, corner));
        }
        return out;
    }
};

typedef unsigned int BvhNodeFlags;
const BvhNodeFlags kSplitX = 0;
const BvhNodeFlags kSplitY = 1;
const BvhNodeFlags kSplitZ = 2;
const BvhNodeFlags kSplitFlags = 0x3;
const BvhNodeFlags kLeafNode = 0x4;

// 32 bytes, bit-compatible with RtBvhNode (RAccel.h:136-145)
struct BvhNode
{
    BBox m_bbox;
    union
    {
        unsigned int m_firstChild;
        unsigned int m_prim;
    };
    BvhNodeFlags m_flags;

    bool leafNode() const { return (m_flags & kLeafNode) != 0; }
    bool interiorNode() const { return (m_flags & kLeafNode) == 0; }
    BvhNodeFlags split() const { return m_flags & kSplitFlags; }
    unsigned int leftChildIndex() const { return m_firstChild; }
    unsigned int rightChildIndex() const { return m_firstChild + 1; }
    unsigned int prim() const { return m_prim; }
};

// The reference's traversal keeps a 50-entry stack and silently stops when it
// would overflow (RAccel.h:379,414,502).  The GPU path refuses such trees.
const unsigned int kMaxTraversalSteps = 50;

} // namespace Rayito

namespace rayito_b200
{
// Which face BVH Mesh::prepare() builds (per calling thread, like stageSemantics()).
//   kTreeReference  the reference's tree, node for node (RAccel.h:262-374): hit records bit-equal.
//   kTreeSah        PERF MODE: same node format, numbering, leaf size and traversal, but every node is
//                   split where a binned surface-area heuristic puts the cut instead of at the midpoint
//                   of its longest axis.  The reference's slab test is not watertight, so a different
//                   tree can decide a grazing ray differently: parity is MEASURED, not bit-exact
//                   (tests/test_gpu_perf_tree.py states the bars).  Never the default.
//   kTreeDevice     the reference's tree again, node for node, but built ON THE GPU inside raytrace()'s scene
//                   upload (rt_scene_create_ex with RT_SCENE_BUILD_MESH_BVH, rayito_b200/csrc/rt_build.cuh):
//                   prepare() leaves the face BVH unbuilt and no node crosses PCIe.  Stage 7 rules only;
//                   with Stage 6 rules the host builds as usual.
//   kTreeAuto       DEFAULT: per mesh, kTreeDevice from kDeviceBuildFaces faces up (where the host build starts
//                   to cost: 5 M quads take 117 ms on 16 host cores and 10.6 ms on the B200, and 320 MB of
//                   nodes stay off PCIe), kTreeReference below.  Either way the tree is the reference's, node
//                   for node.
enum TreeMode { kTreeReference = 0, kTreeSah = 1, kTreeDevice = 2, kTreeAuto = 3 };
const unsigned kDeviceBuildFaces = 1u << 16;
unsigned& treeMode();
}

namespace Rayito
{

// BVH over the elements of T (T provides numElements() and elementBBox(i)).
// One element per leaf; interior nodes split the longest axis of the node box at
// its midpoint; elements whose box centre lies ABOVE the split go first ("left"
// child = high side); if that leaves a side empty the range is cut in half.
template <typename T>
class Bvh
{
public:
    explicit Bvh(T& object) : m_object(object), m_maxDepth(0) { }

    // rootBox: box of the root node when it is not the union of the element boxes
    // (Stage 6 passes m_object.bbox(), S6 RAccel.h:259; Stage 7 the union, RAccel.h:284)
    //
    // Node numbering follows the reference's recursion (RAccel.h:366-371): a node takes
    // the next two free slots for its children, then its left subtree is built completely
    // before its right subtree.  With one element per leaf a subtree over n elements
    // always has 2n-1 nodes, so the slots are known before anything below is built:
    // if a node's first free slot is `base` and its left side gets nL elements, the
    // children sit at base and base+1, the left child's descendants start at base+2
    // and the right child's at base+2nL.  Subtrees therefore build independently -- on
    // worker threads for large meshes -- and land exactly where the serial recursion
    // puts them; std::partition runs on disjoint ranges, so element order is untouched.
    bool build(const BBox* rootBox = NULL, unsigned mode = rayito_b200::kTreeReference)
    {
        m_maxDepth = 0;
        unsigned int count = m_object.numElements();
        if (count == 0)
        {
            m_nodes.release();
            return true;
        }

        // Uninitialised per-thread scratch, kept between builds (rayito_b200::buildScratch)
        const unsigned int threads = count >= kParallelElements ? rayito_b200::hostThreads() : 1u;
        Item* items = static_cast<Item*>(rayito_b200::buildScratch().get((size_t)count * sizeof(Item)));
        if (items == NULL)
            throw std::bad_alloc();
        BBox whole;
        {
            // Element boxes; the union keeps the serial left-to-right association
            // (std::min/max keep their first argument on ties, e.g. -0 against +0)
            unsigned int chunks = threads > 1 ? rayito_b200::chunkCount(count, kParallelElements / 4) : 1u;
            std::vector<BBox> partial(chunks);
            T& object = m_object;
            Item* out = items;
            BBox* part = &partial[0];
            rayito_b200::parallelChunks(count, chunks, [&object, out, part](unsigned c, size_t b, size_t e) {
                BBox acc;
                for (size_t i = b; i < e; ++i)
                {
                    out[i].prim = (unsigned int)i;
                    out[i].box = object.elementBBox((unsigned int)i);
                    acc = acc.combined(out[i].box);
                }
                part[c] = acc;
            });
            for (unsigned int c = 0; c < chunks; ++c)
                whole = whole.combined(partial[c]);
        }
        m_nodes.allocate((size_t)count * 2 - 1);

        // The few splits at the top of a large tree run on one worker each while the others wait
        // for subtrees to exist (5 M quads on the 16-core GPU host: 0.12 s in all; running those
        // splits on every worker with an exact parallel std::partition was measured no faster).
        Job root = { 0, count, 0, 1, 0, rootBox ? *rootBox : whole };
        rayito_b200::JobBag<Job> bag;
        bag.add(root);
        std::mutex depthMutex;
        Builder builder = { items, &m_nodes[0], threads > 1 ? kSpawnElements : 0u, &m_maxDepth, &depthMutex, mode };
        bag.drain(threads, builder);
        return true;
    }

    // No tree on the host (rayito_b200::kTreeDevice: the GPU builds it during the upload)
    void clear()
    {
        m_nodes.release();
        m_maxDepth = 0;
    }

    const BvhNode* nodes() const { return m_nodes.size() == 0 ? NULL : &m_nodes[0]; }
    unsigned int numNodes() const { return (unsigned int)m_nodes.size(); }
    // Depth of the deepest leaf (root = 0); the traversal stack needs depth + 1
    unsigned int maxDepth() const { return m_maxDepth; }

private:
    // malloc'ed array whose elements are written before they are read (every slot of a
    // finished build is); no constructor pass over hundreds of megabytes
    template <typename V>
    class RawArray
    {
    public:
        RawArray() : m_data(NULL), m_size(0) { }
        explicit RawArray(size_t n) : m_data(NULL), m_size(0) { allocate(n); }
        ~RawArray() { release(); }
        void allocate(size_t n)
        {
            if (n == m_size && m_data != NULL)
                return;             // a re-prepare() of the same mesh rewrites every slot
            release();
            m_data = n ? static_cast<V*>(std::malloc(n * sizeof(V))) : NULL;
            if (n && m_data == NULL)
                throw std::bad_alloc();
            m_size = n;
        }
        void release() { std::free(m_data); m_data = NULL; m_size = 0; }
        size_t size() const { return m_size; }
        V& operator[](size_t i) { return m_data[i]; }
        const V& operator[](size_t i) const { return m_data[i]; }
    private:
        RawArray(const RawArray&);
        RawArray& operator=(const RawArray&);
        V* m_data;
        size_t m_size;
    };

    struct Item
    {
        unsigned int prim;
        BBox box;
    };
    struct Job
    {
        unsigned int begin, end;    // element range
        unsigned int node;          // slot of this subtree's root
        unsigned int base;          // first slot of its descendants
        unsigned int depth;
        BBox box;
    };
    // Below this many elements a build stays on the calling thread
    static const unsigned int kParallelElements = 1u << 16;
    // Subtrees at least this large are handed to the job bag instead of the local stack
    static const unsigned int kSpawnElements = 1u << 13;
    // One step of buildRange (RAccel.h:290-374) for the node of `job`: leaf, or split axis,
    // partition, child boxes and the two child jobs.
    static bool splitNode(Item* items, BvhNode* nodes, const Job& job, Job& left, Job& right, unsigned mode)
    {
        BvhNode& node = nodes[job.node];
        node.m_bbox = job.box;
        if (job.end - job.begin <= 1)
        {
            node.m_flags = kLeafNode;
            node.m_prim = items[job.begin].prim;
            return false;
        }
        if (mode == rayito_b200::kTreeSah && splitNodeSah(items, node, job, left, right))
            return true;

        Vector extent = job.box.m_max - job.box.m_min;
        BvhNodeFlags axis;
        if (extent.m_x > extent.m_y)
            axis = extent.m_x > extent.m_z ? kSplitX : kSplitZ;
        else
            axis = extent.m_y > extent.m_z ? kSplitY : kSplitZ;
        float where = (component(job.box.m_max, axis) + component(job.box.m_min, axis)) * 0.5f;
        node.m_flags = axis;

        Item* cut = std::partition(items + job.begin, items + job.end, AboveSplit(where, axis));
        unsigned int mid = (unsigned int)(cut - items);
        if (mid <= job.begin || mid >= job.end)
        {
            mid = job.begin + (job.end - job.begin) / 2;
            if (mid < job.begin + 1) mid = job.begin + 1;
            else if (mid > job.end - 1) mid = job.end - 1;
        }

        BBox leftBox = unionOf(items, job.begin, mid);
        BBox rightBox = unionOf(items, mid, job.end);

        node.m_firstChild = job.base;
        Job l = { job.begin, mid, job.base, job.base + 2, job.depth + 1, leftBox };
        Job r = { mid, job.end, job.base + 1, job.base + 2 * (mid - job.begin), job.depth + 1, rightBox };
        left = l;
        right = r;
        return true;
    }

    // Union of the boxes of items [begin, end) in the serial left-to-right association
    static BBox unionOf(const Item* items, unsigned int begin, unsigned int end)
    {
        BBox all;
        for (unsigned int i = begin; i < end; ++i)
            all = all.combined(items[i].box);
        return all;
    }

    // PERF MODE split (rayito_b200::kTreeSah): binned surface-area heuristic over the element
    // centres, 16 bins on each axis.  Keeps every convention the traversal relies on -- the first
    // ("left") child holds the HIGH side of the cut on the node's split axis, children are numbered
    // base, base + 1 and one element ends up in each leaf -- so the device kernels and the upload do
    // not know which builder made the tree.  false: no cut separates the centres (all equal); the
    // caller then cuts the range in half as the reference does.
    static const int kSahBins = 16;
    static float halfArea(const BBox& b)
    {
        Vector e = b.m_max - b.m_min;
        return e.m_x * e.m_y + e.m_y * e.m_z + e.m_z * e.m_x;
    }
    struct BinAbove
    {
        BvhNodeFlags axis;
        float lo, scale;
        int cut;
        int bin(const Item& it) const
        {
            float c = (component(it.box.m_max, axis) + component(it.box.m_min, axis)) * 0.5f;
            int k = (int)((c - lo) * scale);
            return k < 0 ? 0 : (k > kSahBins - 1 ? kSahBins - 1 : k);
        }
        bool operator()(const Item& it) const { return bin(it) >= cut; }
    };
    static bool splitNodeSah(Item* items, BvhNode& node, const Job& job, Job& left, Job& right)
    {
        BBox centres;
        for (unsigned int i = job.begin; i < job.end; ++i)
            centres.expand((items[i].box.m_max + items[i].box.m_min) * 0.5f);
        float bestCost = std::numeric_limits<float>::max();
        BinAbove best = { kSplitX, 0.0f, 0.0f, 0 };
        BBox bestLow, bestHigh;
        unsigned int bestHighCount = 0;
        for (BvhNodeFlags axis = kSplitX; axis <= kSplitZ; ++axis)
        {
            const float lo = component(centres.m_min, axis), hi = component(centres.m_max, axis);
            if (!(hi > lo))
                continue;
            BinAbove f = { axis, lo, (float)kSahBins / (hi - lo), 0 };
            BBox box[kSahBins];
            unsigned int count[kSahBins] = { 0 };
            for (unsigned int i = job.begin; i < job.end; ++i)
            {
                int k = f.bin(items[i]);
                box[k] = box[k].combined(items[i].box);
                ++count[k];
            }
            // suffix boxes (high side of every cut), then one sweep up from the low side
            BBox highBox[kSahBins];
            unsigned int highCount[kSahBins];
            BBox acc;
            unsigned int n = 0;
            for (int k = kSahBins - 1; k >= 0; --k)
            {
                acc = acc.combined(box[k]);
                n += count[k];
                highBox[k] = acc;
                highCount[k] = n;
            }
            BBox low;
            unsigned int lowCount = 0;
            for (int cut = 1; cut < kSahBins; ++cut)
            {
                low = low.combined(box[cut - 1]);
                lowCount += count[cut - 1];
                if (lowCount == 0 || highCount[cut] == 0)
                    continue;
                float cost = halfArea(low) * (float)lowCount + halfArea(highBox[cut]) * (float)highCount[cut];
                if (cost < bestCost)
                {
                    bestCost = cost;
                    best = f;
                    best.cut = cut;
                    bestLow = low;
                    bestHigh = highBox[cut];
                    bestHighCount = highCount[cut];
                }
            }
        }
        if (bestHighCount == 0)
            return false;
        node.m_flags = best.axis;
        std::partition(items + job.begin, items + job.end, best);
        const unsigned int mid = job.begin + bestHighCount;
        node.m_firstChild = job.base;
        Job l = { job.begin, mid, job.base, job.base + 2, job.depth + 1, bestHigh };
        Job r = { mid, job.end, job.base + 1, job.base + 2 * (mid - job.begin), job.depth + 1, bestLow };
        left = l;
        right = r;
        return true;
    }
    // Builds one subtree depth-first with an explicit work list instead of recursion
    // (degenerate inputs can be ~N deep).
    struct Builder
    {
        Item* items;
        BvhNode* nodes;
        unsigned int spawnElements;     // 0: never hand subtrees to other workers
        unsigned int* maxDepth;
        std::mutex* depthMutex;
        unsigned mode;

        void operator()(const Job& start, rayito_b200::JobBag<Job>& bag) const
        {
            unsigned int deepest = 0;
            std::vector<Job> jobs;
            jobs.push_back(start);
            while (!jobs.empty())
            {
                Job job = jobs.back();
                jobs.pop_back();
                if (job.depth > deepest)
                    deepest = job.depth;
                Job left, right;
                if (!splitNode(items, nodes, job, left, right, mode))
                    continue;
                if (spawnElements != 0 && right.end - right.begin >= spawnElements)
                    bag.add(right);
                else
                    jobs.push_back(right);
                jobs.push_back(left);
            }
            std::lock_guard<std::mutex> lock(*depthMutex);
            if (deepest > *maxDepth)
                *maxDepth = deepest;
        }
    };
    // RAccel.h:226-240
    struct AboveSplit
    {
        float where;
        BvhNodeFlags axis;
        AboveSplit(float w, BvhNodeFlags a) : where(w), axis(a) { }
        bool operator()(const Item& it) const
        {
            return where < (component(it.box.m_max, axis) + component(it.box.m_min, axis)) * 0.5f;
        }
    };

    static float component(const Vector& v, BvhNodeFlags axis)
    {
        return axis == kSplitX ? v.m_x : (axis == kSplitY ? v.m_y : v.m_z);
    }

    T& m_object;
    RawArray<BvhNode> m_nodes;
    unsigned int m_maxDepth;
};

} // namespace Rayito

#endif // RAYITO_B200_ACCEL_HPP
