// Host-side scene graph of the Rayito API surface: materials, Shape, ShapeSet,
// Plane, Sphere, lights, Face/Mesh.
//
// The classes keep the reference's public names and constructor signatures
// (Rayito_Stage7_QT/RMaterial.h, RScene.h, RLight.h, RMesh.h) so scene-building
// code compiles unchanged, but they only DESCRIBE the scene.  prepare() does the
// reference's host-side preparation (key normalisation, bounding boxes, area CDF,
// BVH builds -- bit-identical, see accel.hpp) and flatten() emits the SoA buffers
// the CUDA kernels traverse.  Intersection, sampling and shading are device code.
#ifndef RAYITO_B200_SCENE_HPP
#define RAYITO_B200_SCENE_HPP

#include <cstring>
#include <vector>

#include "accel.hpp"
#include "flat_scene.hpp"

namespace Rayito
{

//
// Materials (RMaterial.h:438-554).  The BRDF each one selects (Lambert, Glossy =
// Ashikhmin-Shirley without anisotropy, PerfectReflection) is evaluated on the GPU.
//
class Material
{
public:
    virtual ~Material() { }
    virtual Color emittance() { return Color(); }
    virtual void describe(RtMaterial& out) const = 0;

protected:
    static void fill(RtMaterial& out, const Color& color, const Color& emit, float exponent, unsigned brdf)
    {
        out.color[0] = color.m_r; out.color[1] = color.m_g; out.color[2] = color.m_b;
        out.emittance[0] = emit.m_r; out.emittance[1] = emit.m_g; out.emittance[2] = emit.m_b;
        out.exponent = exponent;
        out.brdf = brdf;
    }
};

class DiffuseMaterial : public Material
{
public:
    DiffuseMaterial(const Color& color) : m_color(color) { }
    virtual void describe(RtMaterial& out) const { fill(out, m_color, Color(), 0.0f, RT_BRDF_LAMBERT); }
protected:
    Color m_color;
};

class GlossyMaterial : public Material
{
public:
    // exponent = 1 / roughness^2 in float (RMaterial.h:211)
    GlossyMaterial(const Color& color, float roughness)
        : m_color(color), m_exponent(1.0f / (roughness * roughness)) { }
    virtual void describe(RtMaterial& out) const { fill(out, m_color, Color(), m_exponent, RT_BRDF_GLOSSY); }
protected:
    Color m_color;
    float m_exponent;
};

class ReflectionMaterial : public Material
{
public:
    ReflectionMaterial(const Color& color) : m_color(color) { }
    virtual void describe(RtMaterial& out) const { fill(out, m_color, Color(), 0.0f, RT_BRDF_MIRROR); }
protected:
    Color m_color;
};

class Emitter : public Material
{
public:
    Emitter(const Color& color, float power) : m_color(color), m_power(power) { }
    virtual Color emittance() { return m_color * m_power; }
    virtual void describe(RtMaterial& out) const { fill(out, Color(), m_color * m_power, 0.0f, RT_BRDF_NONE); }
protected:
    Color m_color;
    float m_power;
};


//
// Shape base (RScene.h:29-109)
//
class Shape
{
public:
    Shape() : m_transform() { }
    virtual ~Shape() { }

    const Transform& transform() const { return m_transform; }
    Transform& transform() { return m_transform; }

    virtual BBox bbox() = 0;
    virtual bool infiniteExtent() const { return false; }
    virtual void prepare() { m_transform.prepare(); }
    virtual void findLights(std::vector<Shape*>& outLightList) { (void)outLightList; }
    virtual bool isLight() const { return false; }

    // BVH element hooks
    virtual unsigned int numElements() const { return 0; }
    virtual BBox elementBBox(unsigned int) const { return BBox(); }

    // Emit this shape's device description.  `self` arrives zeroed with light = -1;
    // the shape fills type / geom / xform / material.  Returns false (with
    // out.error set) for shapes the GPU core cannot represent.
    virtual bool flatten(rayito_b200::FlatScene& out, RtShape& self)
    {
        (void)self;
        out.error = "shape type has no device representation";
        return false;
    }

protected:
    Transform m_transform;

    static unsigned materialIndex(rayito_b200::FlatScene& out, const Material* m)
    {
        RtMaterial rm;
        if (m != NULL)
            m->describe(rm);
        else
        {
            rm.color[0] = rm.color[1] = rm.color[2] = 0.0f;
            rm.emittance[0] = rm.emittance[1] = rm.emittance[2] = 0.0f;
            rm.exponent = 0.0f;
            rm.brdf = RT_BRDF_NONE;
        }
        return out.addMaterial(m, rm);
    }
};


// One-sided infinite plane with the optional bullseye pattern (RScene.h:273-377)
class Plane : public Shape
{
public:
    Plane(const Point& position, const Vector& normal, Material* pMaterial, bool bullseye = false)
        : Shape(), m_position(position), m_normal(normal.normalized()), m_pMaterial(pMaterial), m_bullseye(bullseye) { }

    virtual BBox bbox() { return BBox(); }
    virtual bool infiniteExtent() const { return true; }

    virtual bool flatten(rayito_b200::FlatScene& out, RtShape& self)
    {
        RtPlane p;
        p.position[0] = m_position.m_x; p.position[1] = m_position.m_y; p.position[2] = m_position.m_z;
        p.normal[0] = m_normal.m_x; p.normal[1] = m_normal.m_y; p.normal[2] = m_normal.m_z;
        p.bullseye = m_bullseye ? 1u : 0u;
        out.planes.push_back(p);
        self.type = RT_SHAPE_PLANE;
        self.geom = (uint32_t)out.planes.size() - 1;
        self.xform = out.addXform(m_transform);
        self.material = materialIndex(out, m_pMaterial);
        return true;
    }

protected:
    Point m_position;
    Vector m_normal;
    Material* m_pMaterial;
    bool m_bullseye;
};


class Sphere : public Shape
{
public:
    Sphere(const Point& position = Point(), float radius = 1.0f, Material* pMaterial = NULL)
        : Shape(), m_position(position), m_radius(radius), m_pMaterial(pMaterial) { }

    void setMaterial(Material* pMaterial) { m_pMaterial = pMaterial; }

    // Union over the key times of the transformed local box (RScene.h:514-524)
    virtual BBox bbox()
    {
        BBox result;
        for (size_t k = 0; k < m_transform.numKeys(); ++k)
        {
            float time = m_transform.keyTime(k);
            result = result.combined(BBox(m_position - Point(m_radius),
                                          m_position + Point(m_radius)).transformFromLocal(time, m_transform));
        }
        return result;
    }

    virtual bool flatten(rayito_b200::FlatScene& out, RtShape& self)
    {
        RtSphere s;
        s.position[0] = m_position.m_x; s.position[1] = m_position.m_y; s.position[2] = m_position.m_z;
        s.radius = m_radius;
        out.spheres.push_back(s);
        self.type = RT_SHAPE_SPHERE;
        self.geom = (uint32_t)out.spheres.size() - 1;
        self.xform = out.addXform(m_transform);
        self.material = materialIndex(out, m_pMaterial);
        return true;
    }

protected:
    Point m_position;
    float m_radius;
    Material* m_pMaterial;
};


//
// Lights (RLight.h)
//
class Light : public Shape
{
public:
    Light(const Color& c, float power) : Shape(), m_color(c), m_power(power), m_material(c, power) { }

    virtual void findLights(std::vector<Shape*>& outLightList) { outLightList.push_back(this); }
    virtual bool isLight() const { return true; }
    virtual Color emitted() const { return m_color * m_power; }

protected:
    Color m_color;
    float m_power;
    Emitter m_material;
};

// Double-sided parallelogram light (RLight.h:43-244)
class RectangleLight : public Light
{
public:
    RectangleLight(const Point& pos, const Vector& side1, const Vector& side2, const Color& color, float power)
        : Light(color, power), m_position(pos), m_side1(side1), m_side2(side2) { }

    // Corners transformed at every key time (RLight.h:165-182)
    virtual BBox bbox()
    {
        Point corners[4] = { m_position, m_position + m_side1, m_position + m_side2, m_position + m_side1 + m_side2 };
        BBox result;
        for (size_t k = 0; k < m_transform.numKeys(); ++k)
        {
            float time = m_transform.keyTime(k);
            for (int i = 0; i < 4; ++i)
                result.expand(m_transform.fromLocalPoint(time, corners[i]));
        }
        return result;
    }

    virtual bool flatten(rayito_b200::FlatScene& out, RtShape& self)
    {
        RtRect r;
        r.position[0] = m_position.m_x; r.position[1] = m_position.m_y; r.position[2] = m_position.m_z;
        r.side1[0] = m_side1.m_x; r.side1[1] = m_side1.m_y; r.side1[2] = m_side1.m_z;
        r.side2[0] = m_side2.m_x; r.side2[1] = m_side2.m_y; r.side2[2] = m_side2.m_z;
        out.rects.push_back(r);
        self.type = RT_SHAPE_RECT;
        self.geom = (uint32_t)out.rects.size() - 1;
        self.xform = out.addXform(m_transform);
        self.material = materialIndex(out, &m_material);
        return true;
    }

protected:
    Point m_position;
    Vector m_side1, m_side2;
};

// Light that borrows the geometry AND the transform of another shape; its own
// transform is never consulted (RLight.h:247-332).
class ShapeLight : public Light
{
public:
    ShapeLight(Shape* pShape, const Color& color, float power) : Light(color, power), m_pShape(pShape) { }

    virtual BBox bbox() { return m_pShape->bbox(); }
    // Note: does NOT prepare its own transform (RLight.h:286-289)
    virtual void prepare() { m_pShape->prepare(); }

    virtual bool flatten(rayito_b200::FlatScene& out, RtShape& self)
    {
        if (!m_pShape->flatten(out, self))
            return false;
        self.material = materialIndex(out, &m_material);
        return true;
    }

protected:
    Shape* m_pShape;
};


//
// Polygon mesh (RMesh.h).  Faces are convex polygons, fan-triangulated from their
// first vertex; the BVH is built over faces, one face per leaf.
//
struct Face
{
    std::vector<unsigned int> m_vertexIndices;
    std::vector<unsigned int> m_normalIndices;   // empty, or one per vertex
};

class Mesh : public Shape
{
public:
    Mesh(const std::vector<Point>& verts, const std::vector<Vector>& normals,
         const std::vector<Face>& faces, Material* pMaterial)
        : m_vertices(verts), m_normals(normals), m_faces(faces), m_pMaterial(pMaterial),
          m_bbox(), m_bvh(*this), m_faceAreaCDF(), m_totalArea(0.0f) { }

    void setMaterial(Material* pMaterial) { m_pMaterial = pMaterial; }

    virtual BBox bbox() { return m_bbox; }

    // RMesh.h:89-129: world box over all vertices at every key time, running face
    // area totals (the CDF used to sample mesh lights), then the face BVH.
    virtual void prepare()
    {
        Shape::prepare();

        m_bbox = BBox();
        for (size_t k = 0; k < m_transform.numKeys(); ++k)
        {
            float time = m_transform.keyTime(k);
            // The interpolated TRS is the same for every vertex at this key time;
            // evaluating it once gives the same bits as the reference's per-vertex
            // re-evaluation because the evaluation is a pure function of time.
            Quaternion rot = m_transform.rotation(time);
            Vector scl = m_transform.scaling(time);
            Vector trn = m_transform.translation(time);
            // Index-ordered chunks, partial boxes folded in chunk order: the same
            // left-to-right min/max association as the reference's single loop
            const unsigned chunks = rayito_b200::chunkCount(m_vertices.size(), 1u << 16);
            std::vector<BBox> partial(chunks);
            BBox* part = &partial[0];
            const Point* verts = m_vertices.empty() ? NULL : &m_vertices[0];
            rayito_b200::parallelChunks(m_vertices.size(), chunks, [=](unsigned c, size_t b, size_t e) {
                BBox acc;
                for (size_t i = b; i < e; ++i)
                    acc.expand(rot * (verts[i] * scl) + trn);
                part[c] = acc;
            });
            for (unsigned c = 0; c < chunks; ++c)
                m_bbox = m_bbox.combined(partial[c]);
        }

        // Face areas on the worker threads (slot f+1 holds the area of face f), then the
        // running total in face order on this thread: float addition is not associative,
        // so the sum itself stays serial (RMesh.h:107-124).
        const size_t numFaces = m_faces.size();
        m_faceAreaCDF.resize(numFaces + 1);
        {
            float* area = &m_faceAreaCDF[0] + 1;
            const Face* faces = numFaces ? &m_faces[0] : NULL;
            const Point* verts = m_vertices.empty() ? NULL : &m_vertices[0];
            rayito_b200::parallelChunks(numFaces, rayito_b200::chunkCount(numFaces, 1u << 15), [=](unsigned, size_t b, size_t e) {
                for (size_t f = b; f < e; ++f)
                {
                    const std::vector<unsigned int>& vi = faces[f].m_vertexIndices;
                    float faceArea = 0.0f;
                    for (size_t tri = 0; tri + 2 < vi.size(); ++tri)
                    {
                        Point p0 = verts[vi[0]];
                        Point p1 = verts[vi[tri + 1]];
                        Point p2 = verts[vi[tri + 2]];
                        faceArea += cross(p1 - p0, p2 - p0).length() * 0.5f;
                    }
                    area[f] = faceArea;
                }
            });
        }
        m_totalArea = 0.0f;
        for (size_t f = 0; f < numFaces; ++f)
        {
            float faceArea = m_faceAreaCDF[f + 1];
            m_faceAreaCDF[f] = m_totalArea;
            m_totalArea += faceArea;
        }
        m_faceAreaCDF[numFaces] = m_totalArea;

        // (rayito_b200::treeMode(): the reference's tree unless the application asked for the perf-mode one;
        // a large mesh's is built by the GPU during the upload)
        const unsigned mode = rayito_b200::treeMode();
        const bool onDevice = rayito_b200::stageSemantics() == RT_SEMANTICS_STAGE7 &&
                              (mode == rayito_b200::kTreeDevice ||
                               (mode == rayito_b200::kTreeAuto && m_faces.size() >= rayito_b200::kDeviceBuildFaces));
        if (onDevice)
            m_bvh.clear();              // built on the GPU out of the uploaded faces (rt_scene_create_ex)
        else if (rayito_b200::stageSemantics() == RT_SEMANTICS_STAGE6)
            m_bvh.build(&m_bbox, mode);       // all vertices, used by a face or not (S6 RMesh.h:82-86)
        else
            m_bvh.build(NULL, mode);
    }

    virtual unsigned int numElements() const { return (unsigned int)m_faces.size(); }
    virtual BBox elementBBox(unsigned int index) const
    {
        BBox box;
        const std::vector<unsigned int>& vi = m_faces[index].m_vertexIndices;
        for (size_t i = 0; i < vi.size(); ++i)
            box.expand(m_vertices[vi[i]]);
        return box;
    }

    const std::vector<Point>& vertices() const { return m_vertices; }
    const std::vector<Vector>& normals() const { return m_normals; }
    const std::vector<Face>& faces() const { return m_faces; }
    const Bvh<Mesh>& bvh() const { return m_bvh; }

    virtual bool flatten(rayito_b200::FlatScene& out, RtShape& self)
    {
        RtMesh m;
        m.first_vertex = (uint32_t)(out.vertices.size() / 3);
        m.num_vertices = (uint32_t)m_vertices.size();
        m.first_normal = (uint32_t)(out.normals.size() / 3);
        m.num_normals = (uint32_t)m_normals.size();
        m.first_face = (uint32_t)out.faceHasNormals.size();
        m.num_faces = (uint32_t)m_faces.size();
        m.first_node = (uint32_t)out.meshNodes.size();
        m.num_nodes = m_bvh.numNodes();
        m.first_cdf = (uint32_t)out.faceAreaCdf.size();
        m.total_area = m_totalArea;

        // Point / Vector are three packed floats and BvhNode is RtBvhNode bit for bit,
        // so the big arrays are block copies; the face tables are filled by index-ordered
        // chunks on the host worker threads (rayito_b200/parallel.hpp).
        static_assert(sizeof(Point) == 12 && sizeof(Vector) == 12, "Point/Vector must be three packed floats");
        static_assert(sizeof(BvhNode) == sizeof(RtBvhNode), "BvhNode must match RtBvhNode");
        appendFloats(out.vertices, m_vertices.empty() ? NULL : &m_vertices[0].m_x, m_vertices.size() * 3);
        appendFloats(out.normals, m_normals.empty() ? NULL : &m_normals[0].m_x, m_normals.size() * 3);
        // face_start is a global offset array with one closing entry per scene
        if (out.faceStart.empty())
            out.faceStart.push_back(0);

        const size_t numFaces = m_faces.size();
        const unsigned chunks = rayito_b200::chunkCount(numFaces, 1u << 15);
        std::vector<size_t> chunkIndices(chunks + 1, 0);
        std::vector<int> chunkError(chunks, 0);
        {
            const Face* faces = numFaces ? &m_faces[0] : NULL;
            const size_t numVerts = m_vertices.size(), numNormals = m_normals.size();
            size_t* counts = &chunkIndices[0];
            int* errors = &chunkError[0];
            rayito_b200::parallelChunks(numFaces, chunks, [=](unsigned c, size_t b, size_t e) {
                size_t total = 0;
                int err = 0;
                for (size_t f = b; f < e && !err; ++f)
                {
                    const Face& face = faces[f];
                    const size_t n = face.m_vertexIndices.size();
                    const bool hasNormals = !face.m_normalIndices.empty();
                    if (n < 3) { err = 1; break; }
                    if (hasNormals && face.m_normalIndices.size() != n) { err = 2; break; }
                    for (size_t i = 0; i < n; ++i)
                        if (face.m_vertexIndices[i] >= numVerts || (hasNormals && face.m_normalIndices[i] >= numNormals))
                            err = 3;
                    total += n;
                }
                counts[c + 1] = total;
                errors[c] = err;
            });
        }
        for (unsigned c = 0; c < chunks; ++c)
        {
            // first error in face order, as the serial loop reports it
            if (chunkError[c] != 0)
            {
                out.error = chunkError[c] == 1 ? "mesh face with fewer than 3 vertices"
                          : chunkError[c] == 2 ? "mesh face with mismatched normal indices"
                                               : "mesh face index out of range";
                return false;
            }
            chunkIndices[c + 1] += chunkIndices[c];
        }
        {
            const size_t indexBase = out.vertexIndex.size(), faceBase = out.faceHasNormals.size();
            out.vertexIndex.resize(indexBase + chunkIndices[chunks]);
            out.normalIndex.resize(indexBase + chunkIndices[chunks]);
            out.faceHasNormals.resize(faceBase + numFaces);
            out.faceStart.resize(faceBase + numFaces + 1);
            const Face* faces = numFaces ? &m_faces[0] : NULL;
            uint32_t* vertexIndex = out.vertexIndex.empty() ? NULL : &out.vertexIndex[0];
            uint32_t* normalIndex = out.normalIndex.empty() ? NULL : &out.normalIndex[0];
            uint32_t* faceHasNormals = out.faceHasNormals.empty() ? NULL : &out.faceHasNormals[0];
            uint32_t* faceStart = &out.faceStart[0];
            const size_t* offsets = &chunkIndices[0];
            rayito_b200::parallelChunks(numFaces, chunks, [=](unsigned c, size_t b, size_t e) {
                size_t at = indexBase + offsets[c];
                for (size_t f = b; f < e; ++f)
                {
                    const Face& face = faces[f];
                    const size_t n = face.m_vertexIndices.size();
                    const bool hasNormals = !face.m_normalIndices.empty();
                    for (size_t i = 0; i < n; ++i)
                    {
                        vertexIndex[at + i] = face.m_vertexIndices[i];
                        normalIndex[at + i] = hasNormals ? face.m_normalIndices[i] : RT_NO_INDEX;
                    }
                    at += n;
                    faceHasNormals[faceBase + f] = hasNormals ? 1u : 0u;
                    faceStart[faceBase + f + 1] = (uint32_t)at;
                }
            });
        }
        if (m_bvh.numNodes() != 0)
        {
            const size_t nodeBase = out.meshNodes.size();
            out.meshNodes.resize(nodeBase + m_bvh.numNodes());
            RtBvhNode* dst = &out.meshNodes[nodeBase];
            const BvhNode* src = m_bvh.nodes();
            const size_t numNodes = m_bvh.numNodes();
            rayito_b200::parallelChunks(numNodes, rayito_b200::chunkCount(numNodes, 1u << 16), [=](unsigned, size_t b, size_t e) {
                std::memcpy(dst + b, src + b, (e - b) * sizeof(RtBvhNode));
            });
        }
        out.faceAreaCdf.insert(out.faceAreaCdf.end(), m_faceAreaCDF.begin(), m_faceAreaCDF.end());
        out.meshes.push_back(m);
        out.meshDepth.push_back(m_bvh.maxDepth());

        self.type = RT_SHAPE_MESH;
        self.geom = (uint32_t)out.meshes.size() - 1;
        self.xform = out.addXform(m_transform);
        self.material = materialIndex(out, m_pMaterial);
        return true;
    }

protected:
    static void appendFloats(std::vector<float>& dst, const float* src, size_t n)
    {
        if (n != 0)
            dst.insert(dst.end(), src, src + n);
    }

    std::vector<Point> m_vertices;
    std::vector<Vector> m_normals;
    std::vector<Face> m_faces;
    Material* m_pMaterial;
    BBox m_bbox;
    Bvh<Mesh> m_bvh;
    std::vector<float> m_faceAreaCDF;
    float m_totalArea;
};

// Wavefront OBJ reader (OBJMesh.cpp:49-181); returns NULL for an empty mesh.
Mesh* createFromOBJFile(const char* filename);


//
// ShapeSet: the scene root (RScene.h:113-269).  Infinite shapes are kept apart and
// tested linearly; finite shapes get a top-level BVH when there are more than two.
//
class ShapeSet : public Shape
{
public:
    ShapeSet() : Shape(), m_shapes(), m_infiniteShapes(), m_bvh(*this) { }

    void addShape(Shape* pShape)
    {
        if (pShape->infiniteExtent())
            m_infiniteShapes.push_back(pShape);
        else
            m_shapes.push_back(pShape);
    }
    void clearShapes() { m_shapes.clear(); m_infiniteShapes.clear(); }

    virtual void prepare()
    {
        Shape::prepare();
        for (size_t i = 0; i < m_infiniteShapes.size(); ++i) m_infiniteShapes[i]->prepare();
        for (size_t i = 0; i < m_shapes.size(); ++i) m_shapes[i]->prepare();
        if (m_shapes.size() > 2)
            m_bvh.build();
    }

    virtual BBox bbox()
    {
        BBox total;
        for (size_t k = 0; k < m_transform.numKeys(); ++k)
        {
            float time = m_transform.keyTime(k);
            for (size_t i = 0; i < m_shapes.size(); ++i)
                total = total.combined(m_shapes[i]->bbox().transformFromLocal(time, m_transform));
        }
        return total;
    }

    virtual void findLights(std::vector<Shape*>& outLightList)
    {
        for (size_t i = 0; i < m_shapes.size(); ++i)
            m_shapes[i]->findLights(outLightList);
    }

    virtual unsigned int numElements() const { return (unsigned int)m_shapes.size(); }
    virtual BBox elementBBox(unsigned int index) const { return m_shapes[index]->bbox(); }

    const std::vector<Shape*>& finiteShapes() const { return m_shapes; }
    const std::vector<Shape*>& infiniteShapes() const { return m_infiniteShapes; }

    // Flatten the whole (prepared) scene.  `lights` is the findLights() list.
    bool flattenScene(rayito_b200::FlatScene& out, const std::vector<Shape*>& lights)
    {
        out.setXform = out.addXform(m_transform);
        out.numFinite = (unsigned)m_shapes.size();
        out.numInfinite = (unsigned)m_infiniteShapes.size();
        std::vector<Shape*> all(m_shapes);
        all.insert(all.end(), m_infiniteShapes.begin(), m_infiniteShapes.end());
        for (size_t i = 0; i < all.size(); ++i)
        {
            RtShape s;
            s.type = s.geom = s.xform = s.material = 0;
            s.light = -1;
            if (!all[i]->flatten(out, s))
                return false;
            for (size_t l = 0; l < lights.size(); ++l)
                if (lights[l] == all[i]) s.light = (int32_t)l;
            out.shapes.push_back(s);
        }
        out.lights.clear();
        for (size_t l = 0; l < lights.size(); ++l)
        {
            size_t idx = all.size();
            for (size_t i = 0; i < all.size(); ++i)
                if (all[i] == lights[l]) idx = i;
            if (idx == all.size())
            {
                out.error = "light is not a member of the scene";
                return false;
            }
            out.lights.push_back((uint32_t)idx);
        }
        out.topNodes.clear();
        out.topDepth = 0;
        if (m_shapes.size() > 2)
        {
            const BvhNode* nodes = m_bvh.nodes();
            for (unsigned int i = 0; i < m_bvh.numNodes(); ++i)
            {
                RtBvhNode n;
                n.bbox_min[0] = nodes[i].m_bbox.m_min.m_x; n.bbox_min[1] = nodes[i].m_bbox.m_min.m_y; n.bbox_min[2] = nodes[i].m_bbox.m_min.m_z;
                n.bbox_max[0] = nodes[i].m_bbox.m_max.m_x; n.bbox_max[1] = nodes[i].m_bbox.m_max.m_y; n.bbox_max[2] = nodes[i].m_bbox.m_max.m_z;
                n.first_child_or_prim = nodes[i].m_firstChild;
                n.flags = nodes[i].m_flags;
                out.topNodes.push_back(n);
            }
            out.topDepth = m_bvh.maxDepth();
        }
        return true;
    }

    virtual bool flatten(rayito_b200::FlatScene& out, RtShape& self)
    {
        (void)self;
        out.error = "nested ShapeSet is not supported by the device scene";
        return false;
    }

protected:
    std::vector<Shape*> m_shapes;
    std::vector<Shape*> m_infiniteShapes;
    Bvh<ShapeSet> m_bvh;
};

} // namespace Rayito

#endif // RAYITO_B200_SCENE_HPP
