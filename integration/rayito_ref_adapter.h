// rayito_ref_adapter.h -- the binding a Rayito maintainer adds to the REFERENCE tree to put
// its own classes on the B200 render core (option B).
//
// Include it in ONE translation unit of the reference application, after nothing else of
// Rayito: it includes the reference's own headers (Rayito_Stage7_QT/rayito.h, RMesh.h) with
// their protected / private sections opened -- the reference keeps the state the device needs
// (ShapeSet::m_shapes, Bvh::m_nodes, Transform keys, Mesh tables; RScene.h:246-267,
// RAccel.h:257-260, RMath.h:845-848, RMesh.h:252-259) non-public and offers no accessors.
// Access specifiers change neither layout nor name mangling under the Itanium C++ ABI, so the
// unit stays link-compatible with the rest of the application.  Nothing of the reference is
// modified or copied: the adapter only READS a scene the reference prepared itself.
//
//   Rayito::ShapeSet scene; ... build it with the reference's API ...
//   Rayito::PerspectiveCamera cam(...);
//   Rayito::Image* img = rayito_b200_adapter::raytrace(scene, cam, W, H, ps, ls, depth);
//
// raytrace() below mirrors Rayito::raytrace() (RaytraceMain.cpp:485-579): findLights, then the
// reference's OWN scene.prepare() (its BVH builds, key normalisation, area CDFs), then
// flattenForDevice() -> rt_scene_create -> rt_render, and the pixels go back into a reference Image.
//
// Link with -lrayito_b200 (include/rayito_b200.h is the only other header needed).
#ifndef RAYITO_REF_ADAPTER_H
#define RAYITO_REF_ADAPTER_H

#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#define protected public
#define private public
#include "rayito.h"
#include "RMesh.h"
#undef protected
#undef private

#include "rayito_b200.h"

namespace rayito_b200_adapter
{

// The flattened scene: owns the arrays an RtSceneDesc points into.
struct FlatRefScene
{
    std::vector<RtShape> shapes;
    std::vector<RtBvhNode> topNodes;
    std::vector<RtXform> xforms;
    std::vector<float> keyTime, keyScale, keyRotation, keyTranslation;
    std::vector<RtPlane> planes;
    std::vector<RtSphere> spheres;
    std::vector<RtRect> rects;
    std::vector<RtMesh> meshes;
    std::vector<float> vertices, normals;
    std::vector<uint32_t> faceStart, faceHasNormals, vertexIndex, normalIndex;
    std::vector<RtBvhNode> meshNodes;
    std::vector<float> faceAreaCdf;
    std::vector<RtMaterial> materials;
    std::vector<uint32_t> lights;
    uint32_t setXform, numFinite, numInfinite;
    std::map<const void*, uint32_t> materialIndex;

    FlatRefScene() : setXform(0), numFinite(0), numInfinite(0) { }

    RtSceneDesc desc() const
    {
        RtSceneDesc d;
        std::memset(&d, 0, sizeof(d));
        d.abi_version = RT_ABI_VERSION;
        d.set_xform = setXform;
        d.num_finite = numFinite;
        d.num_infinite = numInfinite;
        d.shapes = data(shapes);
        d.num_top_nodes = (uint32_t)topNodes.size();    d.top_nodes = data(topNodes);
        d.num_xforms = (uint32_t)xforms.size();         d.xforms = data(xforms);
        d.num_keys = (uint32_t)keyTime.size();
        d.key_time = data(keyTime); d.key_scale = data(keyScale);
        d.key_rotation = data(keyRotation); d.key_translation = data(keyTranslation);
        d.num_planes = (uint32_t)planes.size();         d.planes = data(planes);
        d.num_spheres = (uint32_t)spheres.size();       d.spheres = data(spheres);
        d.num_rects = (uint32_t)rects.size();           d.rects = data(rects);
        d.num_meshes = (uint32_t)meshes.size();         d.meshes = data(meshes);
        d.num_vertices = (uint32_t)(vertices.size() / 3); d.vertices = data(vertices);
        d.num_normals = (uint32_t)(normals.size() / 3);   d.normals = data(normals);
        d.num_faces = (uint32_t)faceHasNormals.size();
        d.face_start = data(faceStart);
        d.face_has_normals = data(faceHasNormals);
        d.num_indices = (uint32_t)vertexIndex.size();
        d.vertex_index = data(vertexIndex);
        d.normal_index = data(normalIndex);
        d.num_mesh_nodes = (uint32_t)meshNodes.size();  d.mesh_nodes = data(meshNodes);
        d.num_cdf = (uint32_t)faceAreaCdf.size();       d.face_area_cdf = data(faceAreaCdf);
        d.num_materials = (uint32_t)materials.size();   d.materials = data(materials);
        d.num_lights = (uint32_t)lights.size();         d.lights = data(lights);
        d.semantics = RT_SEMANTICS_STAGE7;
        return d;
    }

private:
    template <typename V> static const V* data(const std::vector<V>& v) { return v.empty() ? NULL : &v[0]; }
};

namespace detail
{

// Transform keys as Transform::prepare() left them (RMath.h:800-811: rotations normalised)
inline uint32_t addXform(FlatRefScene& out, const Rayito::Transform& t)
{
    RtXform x;
    x.first_key = (uint32_t)out.keyTime.size();
    x.num_keys = (uint32_t)t.m_time.size();
    for (size_t k = 0; k < t.m_time.size(); ++k)
    {
        out.keyTime.push_back(t.m_time[k]);
        out.keyScale.push_back(t.m_scale[k].m_x); out.keyScale.push_back(t.m_scale[k].m_y); out.keyScale.push_back(t.m_scale[k].m_z);
        out.keyRotation.push_back(t.m_rotate[k].m_w); out.keyRotation.push_back(t.m_rotate[k].m_v.m_x);
        out.keyRotation.push_back(t.m_rotate[k].m_v.m_y); out.keyRotation.push_back(t.m_rotate[k].m_v.m_z);
        out.keyTranslation.push_back(t.m_translate[k].m_x); out.keyTranslation.push_back(t.m_translate[k].m_y);
        out.keyTranslation.push_back(t.m_translate[k].m_z);
    }
    out.xforms.push_back(x);
    return (uint32_t)out.xforms.size() - 1;
}

// One RtMaterial per distinct Material object (RMaterial.h:455-554)
inline uint32_t addMaterial(FlatRefScene& out, Rayito::Material* m)
{
    std::map<const void*, uint32_t>::const_iterator it = out.materialIndex.find(m);
    if (it != out.materialIndex.end())
        return it->second;
    RtMaterial rm;
    std::memset(&rm, 0, sizeof(rm));
    rm.brdf = RT_BRDF_NONE;
    if (Rayito::DiffuseMaterial* d = dynamic_cast<Rayito::DiffuseMaterial*>(m))
    {
        rm.color[0] = d->m_color.m_r; rm.color[1] = d->m_color.m_g; rm.color[2] = d->m_color.m_b;
        rm.brdf = RT_BRDF_LAMBERT;
    }
    else if (Rayito::GlossyMaterial* g = dynamic_cast<Rayito::GlossyMaterial*>(m))
    {
        rm.color[0] = g->m_color.m_r; rm.color[1] = g->m_color.m_g; rm.color[2] = g->m_color.m_b;
        rm.exponent = g->m_glossy.m_exponent;
        rm.brdf = RT_BRDF_GLOSSY;
    }
    else if (Rayito::ReflectionMaterial* r = dynamic_cast<Rayito::ReflectionMaterial*>(m))
    {
        rm.color[0] = r->m_color.m_r; rm.color[1] = r->m_color.m_g; rm.color[2] = r->m_color.m_b;
        rm.brdf = RT_BRDF_MIRROR;
    }
    else if (m != NULL)
    {
        Rayito::Color e = m->emittance();       // Emitter: colour * power (RMaterial.h:537)
        rm.emittance[0] = e.m_r; rm.emittance[1] = e.m_g; rm.emittance[2] = e.m_b;
    }
    out.materials.push_back(rm);
    out.materialIndex[m] = (uint32_t)out.materials.size() - 1;
    return (uint32_t)out.materials.size() - 1;
}

inline RtBvhNode convertNode(const Rayito::BvhNode& n)
{
    RtBvhNode r;
    r.bbox_min[0] = n.m_bbox.m_min.m_x; r.bbox_min[1] = n.m_bbox.m_min.m_y; r.bbox_min[2] = n.m_bbox.m_min.m_z;
    r.bbox_max[0] = n.m_bbox.m_max.m_x; r.bbox_max[1] = n.m_bbox.m_max.m_y; r.bbox_max[2] = n.m_bbox.m_max.m_z;
    r.first_child_or_prim = n.m_firstChild;
    r.flags = (uint32_t)n.m_flags;
    return r;
}

// One member of the set.  A ShapeLight is flattened to the geometry and transform of the
// shape it wraps, with the light's Emitter as material (RLight.h:250-332).
inline void addShape(FlatRefScene& out, Rayito::Shape* shape, RtShape& self)
{
    if (Rayito::ShapeLight* sl = dynamic_cast<Rayito::ShapeLight*>(shape))
    {
        addShape(out, sl->m_pShape, self);
        self.material = addMaterial(out, &sl->m_material);
        return;
    }
    self.xform = addXform(out, shape->m_transform);
    if (Rayito::Plane* p = dynamic_cast<Rayito::Plane*>(shape))
    {
        RtPlane r;
        r.position[0] = p->m_position.m_x; r.position[1] = p->m_position.m_y; r.position[2] = p->m_position.m_z;
        r.normal[0] = p->m_normal.m_x; r.normal[1] = p->m_normal.m_y; r.normal[2] = p->m_normal.m_z;
        r.bullseye = p->m_bullseye ? 1u : 0u;
        out.planes.push_back(r);
        self.type = RT_SHAPE_PLANE;
        self.geom = (uint32_t)out.planes.size() - 1;
        self.material = addMaterial(out, p->m_pMaterial);
    }
    else if (Rayito::Sphere* s = dynamic_cast<Rayito::Sphere*>(shape))
    {
        RtSphere r;
        r.position[0] = s->m_position.m_x; r.position[1] = s->m_position.m_y; r.position[2] = s->m_position.m_z;
        r.radius = s->m_radius;
        out.spheres.push_back(r);
        self.type = RT_SHAPE_SPHERE;
        self.geom = (uint32_t)out.spheres.size() - 1;
        self.material = addMaterial(out, s->m_pMaterial);
    }
    else if (Rayito::RectangleLight* rl = dynamic_cast<Rayito::RectangleLight*>(shape))
    {
        RtRect r;
        r.position[0] = rl->m_position.m_x; r.position[1] = rl->m_position.m_y; r.position[2] = rl->m_position.m_z;
        r.side1[0] = rl->m_side1.m_x; r.side1[1] = rl->m_side1.m_y; r.side1[2] = rl->m_side1.m_z;
        r.side2[0] = rl->m_side2.m_x; r.side2[1] = rl->m_side2.m_y; r.side2[2] = rl->m_side2.m_z;
        out.rects.push_back(r);
        self.type = RT_SHAPE_RECT;
        self.geom = (uint32_t)out.rects.size() - 1;
        self.material = addMaterial(out, &rl->m_material);
    }
    else if (Rayito::Mesh* mesh = dynamic_cast<Rayito::Mesh*>(shape))
    {
        RtMesh m;
        m.first_vertex = (uint32_t)(out.vertices.size() / 3);
        m.num_vertices = (uint32_t)mesh->m_vertices.size();
        m.first_normal = (uint32_t)(out.normals.size() / 3);
        m.num_normals = (uint32_t)mesh->m_normals.size();
        m.first_face = (uint32_t)out.faceHasNormals.size();
        m.num_faces = (uint32_t)mesh->m_faces.size();
        m.first_node = (uint32_t)out.meshNodes.size();
        m.num_nodes = mesh->m_bvh.m_numNodes;
        m.first_cdf = (uint32_t)out.faceAreaCdf.size();
        m.total_area = mesh->m_totalArea;
        for (size_t i = 0; i < mesh->m_vertices.size(); ++i)
        {
            out.vertices.push_back(mesh->m_vertices[i].m_x);
            out.vertices.push_back(mesh->m_vertices[i].m_y);
            out.vertices.push_back(mesh->m_vertices[i].m_z);
        }
        for (size_t i = 0; i < mesh->m_normals.size(); ++i)
        {
            out.normals.push_back(mesh->m_normals[i].m_x);
            out.normals.push_back(mesh->m_normals[i].m_y);
            out.normals.push_back(mesh->m_normals[i].m_z);
        }
        if (out.faceStart.empty())
            out.faceStart.push_back(0);
        for (size_t f = 0; f < mesh->m_faces.size(); ++f)
        {
            const Rayito::Face& face = mesh->m_faces[f];
            const bool hasNormals = !face.m_normalIndices.empty();
            if (face.m_vertexIndices.size() < 3 || (hasNormals && face.m_normalIndices.size() != face.m_vertexIndices.size()))
                throw std::runtime_error("rayito_b200 adapter: malformed mesh face");
            for (size_t i = 0; i < face.m_vertexIndices.size(); ++i)
            {
                out.vertexIndex.push_back(face.m_vertexIndices[i]);
                out.normalIndex.push_back(hasNormals ? face.m_normalIndices[i] : 0u);
            }
            out.faceHasNormals.push_back(hasNormals ? 1u : 0u);
            out.faceStart.push_back((uint32_t)out.vertexIndex.size());
        }
        for (unsigned int i = 0; i < mesh->m_bvh.m_numNodes; ++i)
            out.meshNodes.push_back(convertNode(mesh->m_bvh.m_nodes[i]));
        out.faceAreaCdf.insert(out.faceAreaCdf.end(), mesh->m_faceAreaCDF.begin(), mesh->m_faceAreaCDF.end());
        out.meshes.push_back(m);
        self.type = RT_SHAPE_MESH;
        self.geom = (uint32_t)out.meshes.size() - 1;
        self.material = addMaterial(out, mesh->m_pMaterial);
    }
    else
        throw std::runtime_error(std::string("rayito_b200 adapter: shape type has no device representation: ") + typeid(*shape).name());
}

} // namespace detail

// What ShapeSet::prepare() left behind (RScene.h:186-205), as the SoA scene of include/rayito_b200.h.
// `scene` must have been prepared by the reference; `lights` is the list findLights() filled.
inline void flattenForDevice(Rayito::ShapeSet& scene, const std::vector<Rayito::Shape*>& lights, FlatRefScene& out)
{
    out = FlatRefScene();
    out.setXform = detail::addXform(out, scene.m_transform);
    out.numFinite = (uint32_t)scene.m_shapes.size();
    out.numInfinite = (uint32_t)scene.m_infiniteShapes.size();
    std::vector<Rayito::Shape*> all(scene.m_shapes);
    all.insert(all.end(), scene.m_infiniteShapes.begin(), scene.m_infiniteShapes.end());
    for (size_t i = 0; i < all.size(); ++i)
    {
        RtShape s;
        s.type = s.geom = s.xform = s.material = 0;
        s.light = -1;
        detail::addShape(out, all[i], s);
        for (size_t l = 0; l < lights.size(); ++l)
            if (lights[l] == all[i]) s.light = (int32_t)l;
        out.shapes.push_back(s);
    }
    for (size_t l = 0; l < lights.size(); ++l)
    {
        size_t idx = all.size();
        for (size_t i = 0; i < all.size(); ++i)
            if (all[i] == lights[l]) idx = i;
        if (idx == all.size())
            throw std::runtime_error("rayito_b200 adapter: a light is not a member of the scene");
        out.lights.push_back((uint32_t)idx);
    }
    // the set builds its own BVH only for more than two finite shapes (RScene.h:203-204)
    if (scene.m_shapes.size() > 2)
        for (unsigned int i = 0; i < scene.m_bvh.m_numNodes; ++i)
            out.topNodes.push_back(detail::convertNode(scene.m_bvh.m_nodes[i]));
}

// PerspectiveCamera after its constructor ran (RaytraceMain.cpp:205-222)
inline void describeCamera(const Rayito::PerspectiveCamera& cam, RtCamera& out)
{
    out.origin[0] = cam.m_origin.m_x; out.origin[1] = cam.m_origin.m_y; out.origin[2] = cam.m_origin.m_z;
    out.forward[0] = cam.m_forward.m_x; out.forward[1] = cam.m_forward.m_y; out.forward[2] = cam.m_forward.m_z;
    out.right[0] = cam.m_right.m_x; out.right[1] = cam.m_right.m_y; out.right[2] = cam.m_right.m_z;
    out.up[0] = cam.m_up.m_x; out.up[1] = cam.m_up.m_y; out.up[2] = cam.m_up.m_z;
    out.tan_fov = cam.m_tanFov;
    out.focal_distance = cam.m_focalDistance;
    out.lens_radius = cam.m_lensRadius;
    out.shutter_open = cam.m_shutterOpen;
    out.shutter_close = cam.m_shutterClose;
}

// Drop-in for Rayito::raytrace() (rayito.h:138-144) on the reference's own classes.
inline Rayito::Image* raytrace(Rayito::ShapeSet& scene, const Rayito::PerspectiveCamera& cam, size_t width, size_t height,
                               unsigned int pixelSamplesHint, unsigned int lightSamplesHint, unsigned int maxRayDepth,
                               int device = 0, RtRenderStats* statsOut = NULL)
{
    std::vector<Rayito::Shape*> lights;
    scene.findLights(lights);           // same order as the reference: lights first, then prepare (RaytraceMain.cpp:494-497)
    scene.prepare();
    FlatRefScene flat;
    flattenForDevice(scene, lights, flat);
    RtSceneDesc desc = flat.desc();
    RtCamera camera;
    describeCamera(cam, camera);
    RtScene* dev = NULL;
    if (rt_scene_create(&desc, device, &dev) != RT_OK)
        throw std::runtime_error(std::string("rt_scene_create: ") + rt_last_error_string());
    RtRenderParams params;
    std::memset(&params, 0, sizeof(params));
    params.width = (uint32_t)width;
    params.height = (uint32_t)height;
    params.pixel_samples_hint = pixelSamplesHint;
    params.light_samples_hint = lightSamplesHint;
    params.max_ray_depth = maxRayDepth;
    params.world = 1;
    std::vector<float> rgb(width * height * 3);
    RtRenderStats stats;
    int rc = rt_render(dev, &camera, &params, rgb.empty() ? NULL : &rgb[0], &stats);
    std::string err = rc == RT_OK ? "" : rt_last_error_string();
    rt_scene_destroy(dev);
    if (rc != RT_OK)
        throw std::runtime_error("rt_render: " + err);
    if (statsOut) *statsOut = stats;
    Rayito::Image* image = new Rayito::Image(width, height);
    for (size_t y = 0; y < height; ++y)
        for (size_t x = 0; x < width; ++x)
        {
            const float* px = &rgb[(y * width + x) * 3];
            image->pixel(x, y) = Rayito::Color(px[0], px[1], px[2]);
        }
    return image;
}

} // namespace rayito_b200_adapter

#endif // RAYITO_REF_ADAPTER_H
