// ORACLE (test infrastructure, never shipped, never on the product path).
//
// C-ABI driver around the UNMODIFIED reference Stage 7 render core.  It is compiled
// by oracle/Makefile together with /root/reference/Rayito_Stage7_QT/RaytraceMain.cpp
// and OBJMesh.cpp *where they lie* (nothing from the reference is copied into this
// repo) into oracle/_ref/libref_s7.so.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg may load that library.
//
// What it exposes:
//   * scene construction through fixtures/scene_recipes.h (the same file
//     the product compiles against its own headers),
//   * closest-hit / any-hit on caller supplied ray batches, with the winning
//     (shape, face, triangle) recovered by a Mesh subclass (no reference edits),
//   * the reference's own raytrace() (16 worker threads, RaytraceMain.cpp:485-579),
//     timed, with every scene.intersect / scene.doesIntersect call counted (the
//     definition of a "ray" in BASELINE.md) and optionally recorded,
//   * small probes of Rng / CorrelatedMultiJitterSampler / Transform / Bvh internals
//     used to pin the product's host code and sample stream.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <list>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

// Open up the reference classes for inspection.  Access specifiers do not change
// layout or name mangling under the Itanium ABI, so this TU stays link-compatible
// with the reference .cpp files compiled without the defines.
#define protected public
#define private public
#include "rayito.h"
#include "RMesh.h"
#undef protected
#undef private

namespace
{

struct ProbeSlot
{
    uint64_t closestCalls;
    uint64_t anyCalls;
    bool record;
    std::vector<float> closestRays;   // 8 floats per ray: o, d, tMax, time
    std::vector<float> anyRays;
    char pad[64];
};

const int kMaxSlots = 256;
ProbeSlot g_slots[kMaxSlots];
std::atomic<int> g_nextSlot(0);
std::atomic<bool> g_recordRays(false);

thread_local ProbeSlot* tl_slot = NULL;
thread_local const void* tl_lastMesh = NULL;
thread_local unsigned tl_lastFace = 0xffffffffu;
thread_local unsigned tl_lastTri = 0xffffffffu;

ProbeSlot& slot()
{
    if (tl_slot == NULL)
    {
        int s = g_nextSlot.fetch_add(1) % kMaxSlots;
        tl_slot = &g_slots[s];
    }
    return *tl_slot;
}

void resetSlots()
{
    for (int i = 0; i < kMaxSlots; ++i)
    {
        g_slots[i].closestCalls = 0;
        g_slots[i].anyCalls = 0;
        g_slots[i].closestRays.clear();
        g_slots[i].anyRays.clear();
    }
}

void pushRay(std::vector<float>& dst, const Rayito::Ray& r)
{
    const float v[8] = { r.m_origin.m_x, r.m_origin.m_y, r.m_origin.m_z,
                         r.m_direction.m_x, r.m_direction.m_y, r.m_direction.m_z,
                         r.m_tMax, r.m_time };
    dst.insert(dst.end(), v, v + 8);
}

// Mesh that remembers which face / fan triangle last accepted a hit.  It runs the
// reference's own protected Mesh::intersectTri in the reference's own loop order
// (RMesh.h:226-238), so results are the reference's.
class ProbeMesh : public Rayito::Mesh
{
public:
    explicit ProbeMesh(const Rayito::Mesh& m)
        : Rayito::Mesh(m.m_vertices, m.m_normals, m.m_faces, m.m_pMaterial)
    {
        m_transform = m.m_transform;
    }

    virtual bool intersect(Rayito::Intersection& isect) { return Rayito::Mesh::intersect(isect); }
    virtual bool doesIntersect(const Rayito::Ray& ray) { return Rayito::Mesh::doesIntersect(ray); }
    virtual bool doesIntersect(const Rayito::Ray& ray, unsigned int index) { return Rayito::Mesh::doesIntersect(ray, index); }

    virtual bool intersect(Rayito::Intersection& isect, unsigned int index)
    {
        bool any = false;
        size_t numTris = m_faces[index].m_vertexIndices.size() - 2;
        for (size_t i = 0; i < numTris; ++i)
        {
            if (intersectTri(index, (unsigned int)i, isect))
            {
                any = true;
                tl_lastMesh = this;
                tl_lastFace = index;
                tl_lastTri = (unsigned int)i;
            }
        }
        return any;
    }
};

Rayito::Mesh* wrapMesh(Rayito::Mesh* plain)
{
    ProbeMesh* probe = new ProbeMesh(*plain);
    delete plain;
    return probe;
}

// ShapeSet whose top-level entry points count (and optionally record) each call:
// exactly the calls pathTrace makes at RaytraceMain.cpp:294, :395 and :423.
class CountingSet : public Rayito::ShapeSet
{
public:
    virtual bool intersect(Rayito::Intersection& isect)
    {
        ProbeSlot& s = slot();
        s.closestCalls++;
        if (g_recordRays.load(std::memory_order_relaxed))
            pushRay(s.closestRays, isect.m_ray);
        return Rayito::ShapeSet::intersect(isect);
    }
    virtual bool doesIntersect(const Rayito::Ray& ray)
    {
        ProbeSlot& s = slot();
        s.anyCalls++;
        if (g_recordRays.load(std::memory_order_relaxed))
            pushRay(s.anyRays, ray);
        return Rayito::ShapeSet::doesIntersect(ray);
    }
    virtual bool intersect(Rayito::Intersection& isect, unsigned int index) { return Rayito::ShapeSet::intersect(isect, index); }
    virtual bool doesIntersect(const Rayito::Ray& ray, unsigned int index) { return Rayito::ShapeSet::doesIntersect(ray, index); }
};

} // namespace

#define RAYITO_RECIPE_WRAP_MESH(meshPtr) wrapMesh(meshPtr)
#include "scene_recipes.h"

namespace
{

struct RefScene
{
    int sceneId;
    std::string objPath;
    unsigned gridU, gridV;
    CountingSet* set;
    rayito_recipes::SceneStore* store;
    bool prepared;
    double prepareSeconds;

    RefScene() : sceneId(0), gridU(0), gridV(0), set(NULL), store(NULL), prepared(false), prepareSeconds(0) { }
    ~RefScene() { delete set; delete store; }
};

bool buildInto(RefScene& rs)
{
    rs.set = new CountingSet();
    rs.store = new rayito_recipes::SceneStore();
    switch (rs.sceneId)
    {
    case 1: return rayito_recipes::buildStage7Scene1(*rs.set, *rs.store, rs.objPath.c_str());
    case 3: return rayito_recipes::buildStage7Scene1(*rs.set, *rs.store, rs.objPath.c_str(), true);
    case 2: return rayito_recipes::buildStage7Scene2(*rs.set, *rs.store);
    case 5: return rayito_recipes::buildSyntheticMeshScene(*rs.set, *rs.store, rs.gridU, rs.gridV);
    case 7: case 8: case 9: return rayito_recipes::buildEdgeScene(*rs.set, *rs.store, rs.sceneId - 7);
    case 10: return rayito_recipes::buildDeepScene(*rs.set, *rs.store, rs.gridU, rs.gridV, 0);
    case 11: return rayito_recipes::buildDeepScene(*rs.set, *rs.store, rs.gridU, rs.gridV, 18);
    default: return false;
    }
}

void prepareOnce(RefScene& rs)
{
    if (rs.prepared)
        return;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    rs.set->prepare();
    rs.prepareSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    rs.prepared = true;
}

int shapeIndexOf(const RefScene& rs, const Rayito::Shape* p)
{
    if (p == NULL)
        return -1;
    const std::vector<Rayito::Shape*>& fin = rs.set->m_shapes;
    for (size_t i = 0; i < fin.size(); ++i)
        if (fin[i] == p) return (int)i;
    const std::vector<Rayito::Shape*>& inf = rs.set->m_infiniteShapes;
    for (size_t i = 0; i < inf.size(); ++i)
        if (inf[i] == p) return (int)(fin.size() + i);
    return -2;
}

// Resolve the geometric shape behind a set member (a ShapeLight forwards to one)
Rayito::Shape* geometryOf(Rayito::Shape* p)
{
    Rayito::ShapeLight* sl = dynamic_cast<Rayito::ShapeLight*>(p);
    return sl ? sl->m_pShape : p;
}

} // namespace

extern "C"
{

struct RefHit
{
    float t;          // kRayTMax-initialised m_t if nothing was hit
    int shape;        // index in finite-then-infinite order, -1 = miss
    int face;         // mesh face index, -1 for analytic shapes
    int tri;          // fan triangle inside the face, -1 for analytic shapes
    float normal[3];
    float colorModifier[3];
};

struct RefRenderStats
{
    double renderSeconds;     // steady_clock around reference raytrace() (includes prepare())
    double prepareSeconds;    // scene.prepare() alone, measured on a twin scene
    uint64_t closestCalls;
    uint64_t anyCalls;
    unsigned threads;         // worker threads the reference created
};

void* ref_scene_create(int sceneId, const char* objPath, unsigned gridU, unsigned gridV)
{
    RefScene* rs = new RefScene();
    rs->sceneId = sceneId;
    rs->objPath = objPath ? objPath : "";
    rs->gridU = gridU;
    rs->gridV = gridV;
    if (!buildInto(*rs))
    {
        delete rs;
        return NULL;
    }
    prepareOnce(*rs);
    return rs;
}

void ref_scene_destroy(void* h) { delete static_cast<RefScene*>(h); }

int ref_scene_num_finite(void* h) { return (int)static_cast<RefScene*>(h)->set->m_shapes.size(); }
int ref_scene_num_infinite(void* h) { return (int)static_cast<RefScene*>(h)->set->m_infiniteShapes.size(); }
double ref_scene_prepare_seconds(void* h) { return static_cast<RefScene*>(h)->prepareSeconds; }

// rays: n x 8 floats (origin, direction, tMax, time)
void ref_trace_closest(void* h, const float* rays, size_t n, RefHit* out)
{
    RefScene& rs = *static_cast<RefScene*>(h);
    for (size_t i = 0; i < n; ++i)
    {
        const float* r = rays + i * 8;
        Rayito::Ray ray(Rayito::Point(r[0], r[1], r[2]), Rayito::Vector(r[3], r[4], r[5]), r[6], r[7]);
        Rayito::Intersection isect(ray);
        tl_lastMesh = NULL;
        tl_lastFace = tl_lastTri = 0xffffffffu;
        bool hit = rs.set->Rayito::ShapeSet::intersect(isect);
        RefHit& o = out[i];
        o.t = isect.m_t;
        o.shape = hit ? shapeIndexOf(rs, isect.m_pShape) : -1;
        o.face = o.tri = -1;
        if (hit && o.shape >= 0 && o.shape < (int)rs.set->m_shapes.size())
        {
            Rayito::Shape* geom = geometryOf(rs.set->m_shapes[o.shape]);
            if (geom == tl_lastMesh)
            {
                o.face = (int)tl_lastFace;
                o.tri = (int)tl_lastTri;
            }
        }
        o.normal[0] = isect.m_normal.m_x; o.normal[1] = isect.m_normal.m_y; o.normal[2] = isect.m_normal.m_z;
        o.colorModifier[0] = isect.m_colorModifier.m_r;
        o.colorModifier[1] = isect.m_colorModifier.m_g;
        o.colorModifier[2] = isect.m_colorModifier.m_b;
    }
}

void ref_trace_any(void* h, const float* rays, size_t n, uint8_t* out)
{
    RefScene& rs = *static_cast<RefScene*>(h);
    for (size_t i = 0; i < n; ++i)
    {
        const float* r = rays + i * 8;
        Rayito::Ray ray(Rayito::Point(r[0], r[1], r[2]), Rayito::Vector(r[3], r[4], r[5]), r[6], r[7]);
        out[i] = rs.set->Rayito::ShapeSet::doesIntersect(ray) ? 1 : 0;
    }
}

// Run the reference raytrace() on a FRESH copy of the scene (so prepare() runs
// exactly once, as in the GUI).  cam: fov, origin[3], target[3], up[3], focal,
// lens, shutterOpen, shutterClose (14 floats).  rgb: W*H*3 floats, row-major.
// If recordRays != 0 the rays are kept for ref_recorded_rays().
int ref_render(void* h, const float* cam, size_t width, size_t height,
               unsigned ps, unsigned ls, unsigned depth,
               float* rgb, RefRenderStats* stats, int recordRays)
{
    RefScene& proto = *static_cast<RefScene*>(h);
    RefScene fresh;
    fresh.sceneId = proto.sceneId;
    fresh.objPath = proto.objPath;
    fresh.gridU = proto.gridU;
    fresh.gridV = proto.gridV;
    if (!buildInto(fresh))
        return 1;
    Rayito::PerspectiveCamera camera(cam[0],
                                     Rayito::Point(cam[1], cam[2], cam[3]),
                                     Rayito::Point(cam[4], cam[5], cam[6]),
                                     Rayito::Point(cam[7], cam[8], cam[9]),
                                     cam[10], cam[11], cam[12], cam[13]);
    resetSlots();
    g_recordRays.store(recordRays != 0);
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    Rayito::Image* image = Rayito::raytrace(*fresh.set, camera, width, height, ps, ls, depth);
    double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    g_recordRays.store(false);
    if (rgb != NULL)
    {
        for (size_t y = 0; y < height; ++y)
            for (size_t x = 0; x < width; ++x)
            {
                const Rayito::Color& c = image->pixel(x, y);
                float* px = rgb + (y * width + x) * 3;
                px[0] = c.m_r; px[1] = c.m_g; px[2] = c.m_b;
            }
    }
    delete image;
    if (stats != NULL)
    {
        stats->renderSeconds = seconds;
        stats->prepareSeconds = proto.prepareSeconds;
        stats->closestCalls = stats->anyCalls = 0;
        for (int i = 0; i < kMaxSlots; ++i)
        {
            stats->closestCalls += g_slots[i].closestCalls;
            stats->anyCalls += g_slots[i].anyCalls;
        }
        size_t cw = width >= 4 ? width / 4 : 1, chh = height >= 4 ? height / 4 : 1;
        size_t xc = width > 4 ? width / cw : 1, yc = height > 4 ? height / chh : 1;
        if (xc * cw < width) xc++;
        if (yc * chh < height) yc++;
        stats->threads = (unsigned)(xc * yc);
    }
    return 0;
}

// kind 0 = closest-hit rays, 1 = any-hit rays.  Returns the number of recorded
// rays; copies up to cap of them (8 floats each) into out when out != NULL.
size_t ref_recorded_rays(int kind, float* out, size_t cap)
{
    size_t total = 0;
    for (int i = 0; i < kMaxSlots; ++i)
    {
        const std::vector<float>& v = kind == 0 ? g_slots[i].closestRays : g_slots[i].anyRays;
        size_t n = v.size() / 8;
        if (out != NULL)
        {
            size_t room = cap > total ? cap - total : 0;
            size_t take = std::min(room, n);
            if (take) std::memcpy(out + total * 8, v.data(), take * 8 * sizeof(float));
        }
        total += n;
    }
    return total;
}

// ---- probes of reference internals (pin the product's host code) -----------

// BVH nodes of the top-level set (shape < 0) or of a mesh member.  Each node is
// written as 8 words: min xyz, max xyz (float bits), firstChild/prim, flags.
// Returns the node count (0 when the set uses its linear list).
unsigned ref_bvh_nodes(void* h, int shape, uint32_t* out, unsigned capNodes)
{
    RefScene& rs = *static_cast<RefScene*>(h);
    const Rayito::BvhNode* nodes = NULL;
    unsigned count = 0;
    if (shape < 0)
    {
        nodes = rs.set->m_bvh.m_nodes;
        count = rs.set->m_bvh.m_numNodes;
    }
    else
    {
        Rayito::Mesh* mesh = dynamic_cast<Rayito::Mesh*>(geometryOf(rs.set->m_shapes[shape]));
        if (mesh == NULL) return 0;
        nodes = mesh->m_bvh.m_nodes;
        count = mesh->m_bvh.m_numNodes;
    }
    if (out != NULL)
    {
        for (unsigned i = 0; i < count && i < capNodes; ++i)
            std::memcpy(out + (size_t)i * 8, &nodes[i], 32);
    }
    return count;
}

// Mesh geometry as the reference's OBJ reader produced it.  Any out pointer may
// be NULL.  faceSizes gets one entry per face; indices are concatenated.
int ref_mesh_counts(void* h, int shape, unsigned* nVerts, unsigned* nNormals, unsigned* nFaces, unsigned* nIndices)
{
    RefScene& rs = *static_cast<RefScene*>(h);
    Rayito::Mesh* mesh = dynamic_cast<Rayito::Mesh*>(geometryOf(rs.set->m_shapes[shape]));
    if (mesh == NULL) return 1;
    *nVerts = (unsigned)mesh->m_vertices.size();
    *nNormals = (unsigned)mesh->m_normals.size();
    *nFaces = (unsigned)mesh->m_faces.size();
    unsigned total = 0;
    for (size_t f = 0; f < mesh->m_faces.size(); ++f) total += (unsigned)mesh->m_faces[f].m_vertexIndices.size();
    *nIndices = total;
    return 0;
}

int ref_mesh_data(void* h, int shape, float* verts, float* normals, unsigned* faceSizes,
                  unsigned* vertexIndices, unsigned* normalIndices, float* areaCdf, float* bbox6)
{
    RefScene& rs = *static_cast<RefScene*>(h);
    Rayito::Mesh* mesh = dynamic_cast<Rayito::Mesh*>(geometryOf(rs.set->m_shapes[shape]));
    if (mesh == NULL) return 1;
    if (verts) std::memcpy(verts, mesh->m_vertices.data(), mesh->m_vertices.size() * 12);
    if (normals && !mesh->m_normals.empty()) std::memcpy(normals, mesh->m_normals.data(), mesh->m_normals.size() * 12);
    size_t k = 0;
    for (size_t f = 0; f < mesh->m_faces.size(); ++f)
    {
        const Rayito::Face& face = mesh->m_faces[f];
        if (faceSizes) faceSizes[f] = (unsigned)face.m_vertexIndices.size();
        for (size_t i = 0; i < face.m_vertexIndices.size(); ++i, ++k)
        {
            if (vertexIndices) vertexIndices[k] = face.m_vertexIndices[i];
            if (normalIndices) normalIndices[k] = i < face.m_normalIndices.size() ? face.m_normalIndices[i] : 0xffffffffu;
        }
    }
    if (areaCdf) std::memcpy(areaCdf, mesh->m_faceAreaCDF.data(), mesh->m_faceAreaCDF.size() * 4);
    if (bbox6) std::memcpy(bbox6, &mesh->m_bbox, 24);
    return 0;
}

// Transform keys of a set member's geometry after prepare(): per key 11 floats
// (time, scale xyz, rotation wxyz, translation xyz).  Returns the key count
// (0 = keyless identity transform).
unsigned ref_shape_keys(void* h, int shape, float* out, unsigned capKeys)
{
    RefScene& rs = *static_cast<RefScene*>(h);
    int nFinite = (int)rs.set->m_shapes.size();
    Rayito::Shape* s = shape < nFinite ? geometryOf(rs.set->m_shapes[shape]) : rs.set->m_infiniteShapes[shape - nFinite];
    const Rayito::Transform& t = s->m_transform;
    unsigned n = (unsigned)t.m_time.size();
    for (unsigned i = 0; out != NULL && i < n && i < capKeys; ++i)
    {
        float* o = out + (size_t)i * 11;
        o[0] = t.m_time[i];
        o[1] = t.m_scale[i].m_x; o[2] = t.m_scale[i].m_y; o[3] = t.m_scale[i].m_z;
        o[4] = t.m_rotate[i].m_w; o[5] = t.m_rotate[i].m_v.m_x; o[6] = t.m_rotate[i].m_v.m_y; o[7] = t.m_rotate[i].m_v.m_z;
        o[8] = t.m_translate[i].m_x; o[9] = t.m_translate[i].m_y; o[10] = t.m_translate[i].m_z;
    }
    return n;
}

// The MWC generator, stepped literally (RSampling.h:52-57)
void ref_rng_sequence(unsigned z, unsigned w, size_t count, unsigned* out)
{
    Rayito::Rng rng(z, w);
    for (size_t i = 0; i < count; ++i) out[i] = rng.nextUInt32();
}

// CorrelatedMultiJitterSampler::sample1D / sample2D for index 0..count-1
void ref_cmj_1d(unsigned samples, unsigned permutation, unsigned count, float* out)
{
    Rayito::Rng rng;
    Rayito::CorrelatedMultiJitterSampler s(samples, rng, permutation);
    for (unsigned i = 0; i < count; ++i) out[i] = s.sample1D(i);
}

void ref_cmj_2d(unsigned xSamples, unsigned ySamples, unsigned permutation, unsigned count, float* out)
{
    Rayito::Rng rng;
    Rayito::CorrelatedMultiJitterSampler s(xSamples, ySamples, rng, permutation);
    for (unsigned i = 0; i < count; ++i) s.sample2D(i, out[2 * i], out[2 * i + 1]);
}

// PerspectiveCamera::makeRay for a list of (xScreen, yScreen, lensU, lensV, timeU)
void ref_camera_rays(const float* cam, const float* args5, size_t n, float* rays8)
{
    Rayito::PerspectiveCamera camera(cam[0],
                                     Rayito::Point(cam[1], cam[2], cam[3]),
                                     Rayito::Point(cam[4], cam[5], cam[6]),
                                     Rayito::Point(cam[7], cam[8], cam[9]),
                                     cam[10], cam[11], cam[12], cam[13]);
    for (size_t i = 0; i < n; ++i)
    {
        const float* a = args5 + i * 5;
        Rayito::Ray r = camera.makeRay(a[0], a[1], a[2], a[3], a[4]);
        float* o = rays8 + i * 8;
        o[0] = r.m_origin.m_x; o[1] = r.m_origin.m_y; o[2] = r.m_origin.m_z;
        o[3] = r.m_direction.m_x; o[4] = r.m_direction.m_y; o[5] = r.m_direction.m_z;
        o[6] = r.m_tMax; o[7] = r.m_time;
    }
}

const char* ref_build_info()
{
    return "reference: Rayito_Stage7_QT (unmodified RaytraceMain.cpp + OBJMesh.cpp), g++ " __VERSION__
           ", -O3, no -march, no -ffast-math";
}

} // extern "C"
