// FlatScene: the SoA form of a prepared scene that crosses the C ABI.
//
// ShapeSet::flatten() (scene.hpp) fills one of these after prepare(); desc()
// then yields the RtSceneDesc view that rt_scene_create() uploads.  The arrays
// are exactly the state the reference holds after scene.prepare()
// (Rayito_Stage7_QT/RScene.h:186-205, RMesh.h:89-129): transform keys with
// normalised rotations, reference-format BVH nodes, polygon faces, area CDFs.
#ifndef RAYITO_B200_FLAT_SCENE_HPP
#define RAYITO_B200_FLAT_SCENE_HPP

#include <map>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "rayito_b200.h"
#include "math.hpp"

namespace rayito_b200
{

// Which tutorial stage's rules raytrace() and prepare() follow: RT_SEMANTICS_STAGE7
// (default) or RT_SEMANTICS_STAGE6 (see RtSceneDesc.semantics).  Process-wide; an
// application built for the Stage 6 API sets it once before building its scene
// (or compiles with -DRAYITO_B200_STAGE=6, see rayito.h).
unsigned& stageSemantics();

// std::vector whose resize() leaves new elements uninitialised: the big face and node tables
// are sized first and then written by the worker threads; zero-filling hundreds of megabytes
// on one thread beforehand would cost as much as the fill itself.
template <typename T>
struct DefaultInitAllocator : std::allocator<T>
{
    template <typename U> struct rebind { typedef DefaultInitAllocator<U> other; };
    DefaultInitAllocator() { }
    template <typename U> DefaultInitAllocator(const DefaultInitAllocator<U>&) { }
    template <typename U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <typename U, typename A0, typename... Args> void construct(U* p, A0&& a0, Args&&... args)
    {
        ::new (static_cast<void*>(p)) U(std::forward<A0>(a0), std::forward<Args>(args)...);
    }
};
template <typename T> struct RawVector { typedef std::vector<T, DefaultInitAllocator<T> > type; };

struct FlatScene
{
    unsigned setXform;
    unsigned numFinite, numInfinite;
    std::vector<RtShape> shapes;
    std::vector<RtBvhNode> topNodes;
    unsigned topDepth;

    std::vector<RtXform> xforms;
    std::vector<float> keyTime, keyScale, keyRotation, keyTranslation;

    std::vector<RtPlane> planes;
    std::vector<RtSphere> spheres;
    std::vector<RtRect> rects;
    std::vector<RtMesh> meshes;
    std::vector<unsigned> meshDepth;

    std::vector<float> vertices, normals;
    RawVector<uint32_t>::type faceStart, faceHasNormals, vertexIndex, normalIndex;
    RawVector<RtBvhNode>::type meshNodes;
    std::vector<float> faceAreaCdf;

    std::vector<RtMaterial> materials;
    std::vector<uint32_t> lights;

    unsigned semantics;          // RT_SEMANTICS_*
    std::string error;

    FlatScene() : setXform(0), numFinite(0), numInfinite(0), topDepth(0), semantics(RT_SEMANTICS_STAGE7) { }

    // Empty the scene but keep the arrays' storage: raytrace() flattens into one
    // per-thread FlatScene so that a large scene's ~GB of tables is not handed back to
    // the OS and page-faulted in again on every call.
    void reset()
    {
        setXform = 0; numFinite = 0; numInfinite = 0; topDepth = 0;
        shapes.clear(); topNodes.clear();
        xforms.clear(); keyTime.clear(); keyScale.clear(); keyRotation.clear(); keyTranslation.clear();
        planes.clear(); spheres.clear(); rects.clear(); meshes.clear(); meshDepth.clear();
        vertices.clear(); normals.clear();
        faceStart.clear(); faceHasNormals.clear(); vertexIndex.clear(); normalIndex.clear();
        meshNodes.clear(); faceAreaCdf.clear();
        materials.clear(); lights.clear();
        semantics = RT_SEMANTICS_STAGE7;
        error.clear();
        m_materialIndex.clear();
    }
    // Give the storage back (rayito_b200::releaseHostCaches)
    void shrink() { FlatScene empty; std::swap(*this, empty); }

    unsigned addXform(const Rayito::Transform& t)
    {
        RtXform x;
        x.first_key = (uint32_t)keyTime.size();
        x.num_keys = (uint32_t)t.storedKeys();
        for (size_t k = 0; k < t.storedKeys(); ++k)
        {
            const Rayito::Vector& s = t.scaleKeys()[k];
            const Rayito::Quaternion& r = t.rotationKeys()[k];
            const Rayito::Vector& tr = t.translationKeys()[k];
            keyTime.push_back(t.keyTimes()[k]);
            keyScale.push_back(s.m_x); keyScale.push_back(s.m_y); keyScale.push_back(s.m_z);
            keyRotation.push_back(r.m_w); keyRotation.push_back(r.m_v.m_x);
            keyRotation.push_back(r.m_v.m_y); keyRotation.push_back(r.m_v.m_z);
            keyTranslation.push_back(tr.m_x); keyTranslation.push_back(tr.m_y); keyTranslation.push_back(tr.m_z);
        }
        xforms.push_back(x);
        return (unsigned)xforms.size() - 1;
    }

    // One RtMaterial per distinct host Material object
    unsigned addMaterial(const void* identity, const RtMaterial& m)
    {
        std::map<const void*, unsigned>::const_iterator it = m_materialIndex.find(identity);
        if (it != m_materialIndex.end())
            return it->second;
        materials.push_back(m);
        m_materialIndex[identity] = (unsigned)materials.size() - 1;
        return (unsigned)materials.size() - 1;
    }

    RtSceneDesc desc() const
    {
        RtSceneDesc d;
        d.abi_version = RT_ABI_VERSION;
        d.set_xform = setXform;
        d.num_finite = numFinite;
        d.num_infinite = numInfinite;
        d.shapes = ptr(shapes);
        d.num_top_nodes = (uint32_t)topNodes.size();
        d.top_nodes = ptr(topNodes);
        d.num_xforms = (uint32_t)xforms.size();
        d.xforms = ptr(xforms);
        d.num_keys = (uint32_t)keyTime.size();
        d.key_time = ptr(keyTime);
        d.key_scale = ptr(keyScale);
        d.key_rotation = ptr(keyRotation);
        d.key_translation = ptr(keyTranslation);
        d.num_planes = (uint32_t)planes.size();   d.planes = ptr(planes);
        d.num_spheres = (uint32_t)spheres.size(); d.spheres = ptr(spheres);
        d.num_rects = (uint32_t)rects.size();     d.rects = ptr(rects);
        d.num_meshes = (uint32_t)meshes.size();   d.meshes = ptr(meshes);
        d.num_vertices = (uint32_t)(vertices.size() / 3); d.vertices = ptr(vertices);
        d.num_normals = (uint32_t)(normals.size() / 3);   d.normals = ptr(normals);
        d.num_faces = (uint32_t)faceHasNormals.size();
        d.face_start = ptr(faceStart);
        d.face_has_normals = ptr(faceHasNormals);
        d.num_indices = (uint32_t)vertexIndex.size();
        d.vertex_index = ptr(vertexIndex);
        d.normal_index = ptr(normalIndex);
        d.num_mesh_nodes = (uint32_t)meshNodes.size(); d.mesh_nodes = ptr(meshNodes);
        d.num_cdf = (uint32_t)faceAreaCdf.size();      d.face_area_cdf = ptr(faceAreaCdf);
        d.num_materials = (uint32_t)materials.size();  d.materials = ptr(materials);
        d.num_lights = (uint32_t)lights.size();        d.lights = ptr(lights);
        d.semantics = semantics;
        return d;
    }

private:
    std::map<const void*, unsigned> m_materialIndex;

    template <typename V, typename A>
    static const V* ptr(const std::vector<V, A>& v) { return v.empty() ? NULL : &v[0]; }
};

} // namespace rayito_b200

#endif // RAYITO_B200_FLAT_SCENE_HPP
