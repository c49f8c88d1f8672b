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
#include <limits>
#include <vector>

#include "math.hpp"

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
    bool build(const BBox* rootBox = NULL)
    {
        m_nodes.clear();
        m_maxDepth = 0;
        unsigned int count = m_object.numElements();
        if (count == 0)
            return true;

        std::vector<Item> items(count);
        BBox whole;
        for (unsigned int i = 0; i < count; ++i)
        {
            items[i].prim = i;
            items[i].box = m_object.elementBBox(i);
            whole = whole.combined(items[i].box);
        }
        m_nodes.resize((size_t)count * 2 - 1);
        m_used = 1;
        // Depth-first with an explicit work list instead of recursion (degenerate
        // inputs can be ~N deep); children are still numbered in the reference's
        // pre-order: a node reserves both child slots, then its left subtree is
        // built completely before its right subtree (RAccel.h:366-371).
        std::vector<Job> jobs;
        Job root = { 0, count, 0, 0, rootBox ? *rootBox : whole };
        jobs.push_back(root);
        while (!jobs.empty())
        {
            Job job = jobs.back();
            jobs.pop_back();
            if (job.depth > m_maxDepth)
                m_maxDepth = job.depth;
            BvhNode& node = m_nodes[job.node];
            node.m_bbox = job.box;
            if (job.end - job.begin <= 1)
            {
                node.m_flags = kLeafNode;
                node.m_prim = items[job.begin].prim;
                continue;
            }

            Vector extent = job.box.m_max - job.box.m_min;
            BvhNodeFlags axis;
            if (extent.m_x > extent.m_y)
                axis = extent.m_x > extent.m_z ? kSplitX : kSplitZ;
            else
                axis = extent.m_y > extent.m_z ? kSplitY : kSplitZ;
            float where = (component(job.box.m_max, axis) + component(job.box.m_min, axis)) * 0.5f;
            node.m_flags = axis;

            Item* first = &items[0];
            Item* cut = std::partition(first + job.begin, first + job.end, AboveSplit(where, axis));
            unsigned int mid = (unsigned int)(cut - first);
            if (mid <= job.begin || mid >= job.end)
            {
                mid = job.begin + (job.end - job.begin) / 2;
                if (mid < job.begin + 1) mid = job.begin + 1;
                else if (mid > job.end - 1) mid = job.end - 1;
            }

            BBox leftBox, rightBox;
            for (unsigned int i = job.begin; i < mid; ++i) leftBox = leftBox.combined(items[i].box);
            for (unsigned int i = mid; i < job.end; ++i) rightBox = rightBox.combined(items[i].box);

            // In the reference the right child's subtree gets its node numbers only
            // after the whole left subtree; numbering therefore cannot be assigned
            // when the job is queued.  Instead the right job is queued first (so it
            // runs after the left subtree) and takes its children's slots then.
            // Both children of THIS node, however, are reserved right now.
            node.m_firstChild = m_used;
            m_used += 2;
            Job right = { mid, job.end, node.m_firstChild + 1, job.depth + 1, rightBox };
            Job left = { job.begin, mid, node.m_firstChild, job.depth + 1, leftBox };
            jobs.push_back(right);
            jobs.push_back(left);
        }
        return true;
    }

    const BvhNode* nodes() const { return m_nodes.empty() ? NULL : &m_nodes[0]; }
    unsigned int numNodes() const { return (unsigned int)m_nodes.size(); }
    // Depth of the deepest leaf (root = 0); the traversal stack needs depth + 1
    unsigned int maxDepth() const { return m_maxDepth; }

private:
    struct Item
    {
        unsigned int prim;
        BBox box;
    };
    struct Job
    {
        unsigned int begin, end, node, depth;
        BBox box;
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
    std::vector<BvhNode> m_nodes;
    unsigned int m_used;
    unsigned int m_maxDepth;
};

} // namespace Rayito

#endif // RAYITO_B200_ACCEL_HPP
