// Scene recipes for the render hot path.
//
// These functions replay the scene construction the reference GUI performs
// (reference: Rayito_Stage7_QT/MainWindow.cpp:93-137 makeCube, :139-229 scene 1,
// :249-371 scene 2) using ONLY the public Rayito API (Shape/ShapeSet/Transform/
// Material/Light/Mesh/PerspectiveCamera).  The file is deliberately API-neutral:
// it is compiled once against this repo's host headers (the product, GPU-backed
// raytrace()) and once against the reference's own headers (oracle/_ref, the
// checker), which is the drop-in proof for the C++ surface.
//
// Include AFTER "rayito.h" and "RMesh.h" of whichever implementation is in use.
// A build may define RAYITO_RECIPE_WRAP_MESH(ptr) to substitute a Mesh subclass
// (the oracle uses this to record which face/triangle won a hit).
#ifndef RAYITO_B200_SCENE_RECIPES_H
#define RAYITO_B200_SCENE_RECIPES_H

#include <cmath>
#include <cstddef>
#include <vector>

#include "scene_store.h"

namespace rayito_recipes
{

inline Rayito::PerspectiveCamera* makeCamera(const CameraSpec& c)
{
    return new Rayito::PerspectiveCamera(c.fov,
                                         Rayito::Point(c.origin[0], c.origin[1], c.origin[2]),
                                         Rayito::Point(c.target[0], c.target[1], c.target[2]),
                                         Rayito::Point(c.up[0], c.up[1], c.up[2]),
                                         c.focalDistance, c.lensRadius, c.shutterOpen, c.shutterClose);
}

// UI defaults of the reference GUI (Rayito_Stage7_QT/MainWindow.ui: fov 30,
// focal distance 16, lens radius 0, shutter 0 -> 1).
inline CameraSpec defaultCameraScene1()
{
    CameraSpec c = { 30.0f, { -4.0f, 5.0f, 15.0f }, { 0.0f, 0.0f, 0.0f }, { 0.0f, 1.0f, 0.0f },
                     16.0f, 0.0f, 0.0f, 1.0f };
    return c;
}

inline CameraSpec defaultCameraScene2()
{
    CameraSpec c = { 30.0f, { -4.0f, 10.0f, 30.0f }, { 0.0f, 5.0f, 0.0f }, { 0.0f, 1.0f, 0.0f },
                     16.0f, 0.0f, 0.0f, 1.0f };
    return c;
}

// Unit box with the reference's face list: the top face is listed twice and the
// bottom face is missing (MainWindow.cpp:106-135); no normals => flat shading.
inline Rayito::Mesh* makeCubeMesh()
{
    static const float corner[8][3] = {
        { 0, 0, 0 }, { 1, 0, 0 }, { 1, 1, 0 }, { 0, 1, 0 },
        { 0, 0, 1 }, { 1, 0, 1 }, { 1, 1, 1 }, { 0, 1, 1 } };
    static const unsigned quad[6][4] = {
        { 0, 1, 2, 3 }, { 1, 5, 6, 2 }, { 5, 4, 7, 6 },
        { 4, 0, 3, 7 }, { 3, 2, 6, 7 }, { 3, 2, 6, 7 } };
    std::vector<Rayito::Point> verts;
    std::vector<Rayito::Vector> normals;
    std::vector<Rayito::Face> faces(6);
    for (int i = 0; i < 8; ++i)
        verts.push_back(Rayito::Point(corner[i][0], corner[i][1], corner[i][2]));
    for (int f = 0; f < 6; ++f)
        for (int k = 0; k < 4; ++k)
            faces[f].m_vertexIndices.push_back(quad[f][k]);
    return new Rayito::Mesh(verts, normals, faces, NULL);
}

// Stage 7, first scene: bullseye ground plane, four spheres (one moving, one
// mirror), a rotating box, bumpy.obj with three rotation keys, a rectangle
// light and a moving sphere light (MainWindow.cpp:143-219).  Finite-shape order
// is the insertion order and must not change (it fixes BVH prim indices).
// Returns false if the OBJ mesh could not be read.
template <typename SetT>
bool buildStage7Scene1(SetT& set, SceneStore& st, const char* objPath, bool objIsMeshLight = false)
{
    using namespace Rayito;
    Material* blueishLambert  = st.keep(new DiffuseMaterial(Color(0.6f, 0.6f, 0.9f)));
    Material* purplishLambert = st.keep(new DiffuseMaterial(Color(0.8f, 0.3f, 0.7f)));
    Material* reddishLambert  = st.keep(new DiffuseMaterial(Color(0.8f, 0.3f, 0.1f)));
    Material* bluishGlossy    = st.keep(new GlossyMaterial(Color(0.5f, 0.3f, 0.8f), 0.3f));
    Material* greenishGlossy  = st.keep(new GlossyMaterial(Color(0.3f, 0.9f, 0.3f), 0.1f));
    Material* reddishGlossy   = st.keep(new GlossyMaterial(Color(0.8f, 0.1f, 0.1f), 0.3f));
    Material* reflective      = st.keep(new ReflectionMaterial(Color(0.7f, 0.7f, 0.2f)));

    Plane* plane = st.add(new Plane(Point(), Vector(0.0f, 1.0f, 0.0f), blueishLambert, true));
    plane->transform().translate(0.0f, Vector(0.0f, -2.0f, 0.0f));
    set.addShape(plane);

    Sphere* sphere1 = st.add(new Sphere(Point(), 1.0f, purplishLambert));
    sphere1->transform().setTranslation(0.0f, Vector(2.0f, -1.0f, 0.0f));
    sphere1->transform().setTranslation(1.0f, Vector(3.0f, -1.0f, 0.0f));
    set.addShape(sphere1);

    Sphere* sphere2 = st.add(new Sphere(Point(), 2.0f, greenishGlossy));
    sphere2->transform().translate(0.0f, Vector(-3.0f, 0.0f, -2.0f));
    set.addShape(sphere2);

    Sphere* sphere3 = st.add(new Sphere(Point(), 0.5f, bluishGlossy));
    sphere3->transform().translate(0.0f, Vector(1.5f, -1.5f, 2.5f));
    set.addShape(sphere3);

    Sphere* sphere4 = st.add(new Sphere(Point(), 0.5f, reflective));
    sphere4->transform().translate(0.0f, Vector(-2.0f, -1.5f, 1.0f));
    set.addShape(sphere4);

    Mesh* cube = RAYITO_RECIPE_WRAP_MESH(makeCubeMesh());
    st.add(cube);
    cube->setMaterial(reddishLambert);
    cube->transform().translate(0.0f, Vector(0.0f, -2.0f, -2.0f));
    cube->transform().rotate(1.0f, Quaternion(Vector(0.0f, 1.0f, 0.0f), M_PI / 4.0f));
    set.addShape(cube);

    Mesh* rawObj = createFromOBJFile(objPath);
    if (rawObj == NULL)
        return false;
    Mesh* obj = RAYITO_RECIPE_WRAP_MESH(rawObj);
    st.add(obj);
    obj->setMaterial(reddishGlossy);
    obj->transform().setTranslation(0.0f, Vector(0.2f, 0.0f, 0.0f));
    obj->transform().rotate(0.5f, Quaternion(Vector(0.0f, 1.0f, 0.0f), M_PI / 4.0f));
    obj->transform().rotate(1.0f, Quaternion(Vector(0.0f, 1.0f, 0.0f), M_PI / 2.0f));
    if (objIsMeshLight)
    {
        // the reference's MAKE_OBJ_A_MESH_LIGHT variant (MainWindow.cpp:193-196)
        st.wrapped.push_back(st.shapes.back());
        st.shapes.pop_back();
        ShapeLight* meshLight = st.add(new ShapeLight(obj, Color(1.0f, 1.0f, 1.0f), 10.0f));
        set.addShape(meshLight);
    }
    else
        set.addShape(obj);

    RectangleLight* areaLight = st.add(new RectangleLight(Point(),
                                                          Vector(3.0f, 0.0f, 0.0f),
                                                          Vector(0.0f, 0.0f, 3.0f),
                                                          Color(1.0f, 1.0f, 1.0f),
                                                          5.0f));
    areaLight->transform().setTranslation(0.0f, Vector(-1.5f, 4.0f, -1.5f));
    set.addShape(areaLight);

    Sphere* bulb = st.hold(new Sphere(Point(), 0.1f, blueishLambert));
    bulb->transform().setTranslation(0.0f, Vector(0.0f, 0.5f, 4.0f));
    bulb->transform().setTranslation(0.33f, Vector(0.0f, 1.5f, 4.0f));
    bulb->transform().setTranslation(0.67f, Vector(1.0f, 1.5f, 4.0f));
    bulb->transform().setTranslation(1.0f, Vector(1.0f, 0.5f, 4.0f));
    ShapeLight* sphereLight = st.add(new ShapeLight(bulb, Color(1.0f, 1.0f, 0.3f), 100.0f));
    set.addShape(sphereLight);
    return true;
}

// Ballistic position with one elastic bounce off the plane through the origin
// perpendicular to gravity (MainWindow.cpp:249-287); float arithmetic in the
// same association as the reference, because the results become transform keys.
inline Rayito::Point ballisticPosition(const Rayito::Point& start,
                                       const Rayito::Vector& velocity,
                                       float time)
{
    using namespace Rayito;
    const Vector gravity(0.0f, -9.8f, 0.0f);
    Vector up = -gravity.normalized();
    float vUp = dot(velocity, up);
    float pUp = dot(start, up);
    float aUp = -gravity.length();
    float disc = vUp * vUp - 2.0f * aUp * pUp;
    if (disc > 0.0f)
    {
        float tHit = (-vUp - std::sqrt(disc)) / aUp;
        if (tHit < time)
        {
            Point where = start + velocity * tHit + gravity * tHit * tHit * 0.5f;
            Vector vIn = (velocity + gravity * tHit);
            Vector vOut = vIn - 2.0f * up * dot(vIn, up);
            float rest = time - tHit;
            return where + vOut * rest + gravity * rest * rest * 0.5f;
        }
    }
    return start + velocity * time + gravity * time * time * 0.5f;
}

// Stage 7, second scene: ten falling spheres and ten tumbling boxes, each with
// two translation (and rotation) keys, one strong rectangle light
// (MainWindow.cpp:289-361).
template <typename SetT>
bool buildStage7Scene2(SetT& set, SceneStore& st)
{
    using namespace Rayito;
    Material* blueishLambert  = st.keep(new DiffuseMaterial(Color(0.6f, 0.6f, 0.9f)));
    Material* yellowishGlossy = st.keep(new GlossyMaterial(Color(0.9f, 0.9f, 0.3f), 0.3f));
    Material* redLambert      = st.keep(new DiffuseMaterial(Color(1.0f, 0.2f, 0.2f)));

    Plane* plane = st.add(new Plane(Point(), Vector(0.0f, 1.0f, 0.0f), redLambert, true));
    set.addShape(plane);

    const float timeDelta = 0.2f;
    {
        Point start(-10.0f, 10.0f, 0.0f);
        Vector velocity(4.5f, 0.0f, 0.0f);
        float timeOffset = 0.0f;
        for (unsigned int i = 0; i < 10; ++i)
        {
            Point p0 = ballisticPosition(start, velocity, timeOffset);
            Point p1 = ballisticPosition(start, velocity, timeOffset + timeDelta);
            Sphere* s = st.add(new Sphere());
            s->transform().setTranslation(0.0f, p0);
            s->transform().setTranslation(1.0f, p1);
            s->setMaterial(blueishLambert);
            set.addShape(s);
            timeOffset += timeDelta * 2.0f;
        }
    }
    {
        Point start(10.0f, 10.0f, 2.0f);
        Vector velocity(-4.5f, 0.0f, 0.0f);
        float timeOffset = 0.0f;
        for (unsigned int i = 0; i < 10; ++i)
        {
            Point p0 = ballisticPosition(start, velocity, timeOffset);
            Point p1 = ballisticPosition(start, velocity, timeOffset + timeDelta);
            // The angle expressions are evaluated in double and rounded once
            // (M_PI is a double literal in the reference: MainWindow.cpp:337-340)
            float rotation0 = timeOffset * M_PI * 0.5;
            if (rotation0 > M_PI * 2.0f)
                rotation0 -= M_PI * 2.0f;
            float rotation1 = rotation0 + timeDelta * M_PI * 0.5;

            Mesh* cube = RAYITO_RECIPE_WRAP_MESH(makeCubeMesh());
            st.add(cube);
            cube->transform().setTranslation(0.0f, p0);
            cube->transform().setRotation(0.0f, Quaternion(Vector(1.0f, 0.0f, 1.0f).normalized(), rotation0));
            cube->transform().setTranslation(1.0f, p1);
            cube->transform().setRotation(1.0f, Quaternion(Vector(1.0f, 0.0f, 1.0f).normalized(), rotation1));
            cube->setMaterial(yellowishGlossy);
            set.addShape(cube);
            timeOffset += timeDelta * 2.0f;
        }
    }

    RectangleLight* areaLight = st.add(new RectangleLight(Point(),
                                                          Vector(2.0f, 0.0f, 0.0f),
                                                          Vector(0.0f, 0.0f, 2.0f),
                                                          Color(1.0f, 1.0f, 1.0f),
                                                          50.0f));
    areaLight->transform().setTranslation(0.0f, Vector(-1.0f, 15.0f, 1.0f));
    set.addShape(areaLight);
    return true;
}

// Deterministic procedural stress mesh (BASELINE.json configs[4]): a displaced
// UV sphere of gridU x gridV quads, r = 1 + .08 sin(9u) sin(7w) + .02 sin(41u+13w),
// with per-vertex normals.  2236 x 2236 gives 4 999 696 quads = 9 999 392
// triangles.  Geometry is generated in double and rounded once to float so both
// builds produce identical bits.
inline Rayito::Mesh* makeDisplacedSphereMesh(unsigned gridU, unsigned gridV, float scale)
{
    using namespace Rayito;
    std::vector<Point> verts;
    std::vector<Vector> normals;
    std::vector<Face> faces;
    verts.reserve((size_t)(gridU + 1) * (gridV + 1));
    normals.reserve((size_t)(gridU + 1) * (gridV + 1));
    for (unsigned j = 0; j <= gridV; ++j)
    {
        // Keep clear of the exact poles so no quad degenerates to zero area
        double w = (0.0005 + 0.999 * (double)j / (double)gridV) * M_PI;
        for (unsigned i = 0; i <= gridU; ++i)
        {
            double u = (double)i / (double)gridU * 2.0 * M_PI;
            double r = 1.0 + 0.08 * std::sin(9.0 * u) * std::sin(7.0 * w) + 0.02 * std::sin(41.0 * u + 13.0 * w);
            double dx = std::sin(w) * std::cos(u), dy = std::cos(w), dz = std::sin(w) * std::sin(u);
            verts.push_back(Point((float)(scale * r * dx), (float)(scale * r * dy), (float)(scale * r * dz)));
            normals.push_back(Vector((float)dx, (float)dy, (float)dz));
        }
    }
    faces.resize((size_t)gridU * gridV);
    size_t f = 0;
    for (unsigned j = 0; j < gridV; ++j)
    {
        for (unsigned i = 0; i < gridU; ++i, ++f)
        {
            unsigned a = j * (gridU + 1) + i;
            unsigned b = a + 1;
            unsigned c = a + (gridU + 1) + 1;
            unsigned d = a + (gridU + 1);
            unsigned idx[4] = { a, b, c, d };
            for (int k = 0; k < 4; ++k)
            {
                faces[f].m_vertexIndices.push_back(idx[k]);
                faces[f].m_normalIndices.push_back(idx[k]);
            }
        }
    }
    return new Mesh(verts, normals, faces, NULL);
}

// Synthetic big-mesh scene: ground plane, the displaced sphere (glossy) slowly
// rotating about y over the shutter, a diffuse companion sphere, and the same
// two area lights as Stage 7 scene 1.
template <typename SetT>
bool buildSyntheticMeshScene(SetT& set, SceneStore& st, unsigned gridU, unsigned gridV)
{
    using namespace Rayito;
    Material* ground = st.keep(new DiffuseMaterial(Color(0.6f, 0.6f, 0.9f)));
    Material* glossy = st.keep(new GlossyMaterial(Color(0.8f, 0.5f, 0.2f), 0.3f));
    Material* matte  = st.keep(new DiffuseMaterial(Color(0.3f, 0.8f, 0.4f)));

    Plane* plane = st.add(new Plane(Point(), Vector(0.0f, 1.0f, 0.0f), ground, true));
    plane->transform().translate(0.0f, Vector(0.0f, -2.0f, 0.0f));
    set.addShape(plane);

    Mesh* blob = RAYITO_RECIPE_WRAP_MESH(makeDisplacedSphereMesh(gridU, gridV, 2.0f));
    st.add(blob);
    blob->setMaterial(glossy);
    blob->transform().setTranslation(0.0f, Vector(0.0f, 0.2f, 0.0f));
    blob->transform().rotate(1.0f, Quaternion(Vector(0.0f, 1.0f, 0.0f), M_PI / 16.0f));
    set.addShape(blob);

    Sphere* side = st.add(new Sphere(Point(), 0.75f, matte));
    side->transform().translate(0.0f, Vector(3.2f, -1.25f, 1.5f));
    set.addShape(side);

    RectangleLight* areaLight = st.add(new RectangleLight(Point(),
                                                          Vector(3.0f, 0.0f, 0.0f),
                                                          Vector(0.0f, 0.0f, 3.0f),
                                                          Color(1.0f, 1.0f, 1.0f),
                                                          5.0f));
    areaLight->transform().setTranslation(0.0f, Vector(-1.5f, 4.0f, -1.5f));
    set.addShape(areaLight);

    Sphere* bulb = st.hold(new Sphere(Point(), 0.1f, ground));
    bulb->transform().setTranslation(0.0f, Vector(0.0f, 0.5f, 4.0f));
    bulb->transform().setTranslation(1.0f, Vector(1.0f, 1.5f, 4.0f));
    ShapeLight* sphereLight = st.add(new ShapeLight(bulb, Color(1.0f, 1.0f, 0.3f), 100.0f));
    set.addShape(sphereLight);
    return true;
}

// Edge-case scenes for the parity tests (not from the GUI).  They reach code the two GUI
// scenes do not: variant 0 has only two finite shapes, so the set keeps its linear list
// instead of a BVH (RScene.h:203-204), a mesh of a pentagon and a hexagon with per-vertex
// normals (fan triangulation, RMesh.h:226-238, 305-333) under scale + rotation keys, and a
// tilted rectangle light; variant 1 has no lights at all, a sphere under non-uniform scale and
// rotation keys, a mirror sphere and a statically squashed sphere; variant 2 is the empty set.
template <typename SetT>
bool buildEdgeScene(SetT& set, SceneStore& st, int variant)
{
    using namespace Rayito;
    if (variant == 2)
        return true;
    Material* grey = st.keep(new DiffuseMaterial(Color(0.7f, 0.7f, 0.7f)));
    Plane* plane = st.add(new Plane(Point(), Vector(0.0f, 1.0f, 0.0f), grey, true));
    plane->transform().translate(0.0f, Vector(0.0f, -2.0f, 0.0f));
    set.addShape(plane);
    if (variant == 0)
    {
        Material* glossy = st.keep(new GlossyMaterial(Color(0.9f, 0.5f, 0.2f), 0.2f));
        std::vector<Point> verts;
        std::vector<Vector> normals;
        std::vector<Face> faces(2);
        // a pentagon in the z = 0 plane and a hexagon folded away from it along their shared edge
        const float px[5] = { 0.0f, 1.0f, 1.4f, 0.5f, -0.4f }, py[5] = { 0.0f, 0.0f, 0.9f, 1.5f, 0.9f };
        for (int i = 0; i < 5; ++i)
        {
            verts.push_back(Point(px[i], py[i], 0.0f));
            normals.push_back(Vector(0.1f * (px[i] - 0.5f), 0.1f * (py[i] - 0.7f), 1.0f).normalized());
            faces[0].m_vertexIndices.push_back((unsigned)i);
            faces[0].m_normalIndices.push_back((unsigned)i);
        }
        const float hx[4] = { 1.2f, 0.9f, 0.1f, -0.2f }, hy[4] = { -0.6f, -1.1f, -1.1f, -0.6f }, hz[4] = { -0.3f, -0.6f, -0.6f, -0.3f };
        faces[1].m_vertexIndices.push_back(1);
        faces[1].m_vertexIndices.push_back(0);
        for (int i = 0; i < 4; ++i)
        {
            verts.push_back(Point(hx[3 - i], hy[3 - i], hz[3 - i]));
            faces[1].m_vertexIndices.push_back((unsigned)(5 + i));
        }
        Mesh* mesh = RAYITO_RECIPE_WRAP_MESH(new Mesh(verts, normals, faces, glossy));
        st.add(mesh);
        mesh->transform().setScaling(0.0f, Vector(1.5f, 0.7f, 1.2f));
        mesh->transform().setTranslation(0.0f, Vector(-0.5f, -0.5f, 0.3f));
        mesh->transform().setScaling(1.0f, Vector(1.0f, 1.0f, 1.0f));
        mesh->transform().setTranslation(1.0f, Vector(-0.2f, -0.8f, 0.0f));
        mesh->transform().rotate(1.0f, Quaternion(Vector(0.0f, 0.0f, 1.0f), M_PI / 3.0f));
        set.addShape(mesh);

        RectangleLight* light = st.add(new RectangleLight(Point(), Vector(2.0f, 0.0f, 0.0f), Vector(0.0f, 0.0f, 2.0f),
                                                          Color(1.0f, 0.9f, 0.8f), 8.0f));
        light->transform().setTranslation(0.0f, Vector(-1.0f, 3.0f, -1.0f));
        light->transform().rotate(0.0f, Quaternion(Vector(1.0f, 0.0f, 0.0f), M_PI / 7.0f));
        set.addShape(light);
        return true;
    }
    Material* blue = st.keep(new DiffuseMaterial(Color(0.3f, 0.4f, 0.9f)));
    Material* mirror = st.keep(new ReflectionMaterial(Color(0.9f, 0.9f, 0.9f)));
    Sphere* egg = st.add(new Sphere(Point(), 1.0f, blue));
    egg->transform().setScaling(0.0f, Vector(2.0f, 1.0f, 0.5f));
    egg->transform().setTranslation(0.0f, Vector(-1.0f, -0.5f, 0.0f));
    egg->transform().rotate(0.0f, Quaternion(Vector(0.0f, 1.0f, 0.0f), M_PI / 5.0f));
    egg->transform().setScaling(1.0f, Vector(1.0f, 1.5f, 1.0f));
    egg->transform().rotate(1.0f, Quaternion(Vector(0.0f, 0.0f, 1.0f), M_PI / 4.0f));
    set.addShape(egg);
    Sphere* ball = st.add(new Sphere(Point(2.0f, -1.0f, 1.0f), 1.0f, mirror));
    set.addShape(ball);
    Sphere* disc = st.add(new Sphere(Point(), 1.0f, grey));            // three finite shapes: a (small) BVH
    disc->transform().setScaling(0.0f, Vector(1.0f, 0.2f, 1.0f));
    disc->transform().setTranslation(0.0f, Vector(0.5f, 0.5f, -2.0f));
    set.addShape(disc);
    return true;
}

// Wedge of `levels` x `across` quads whose rows halve in size towards the tip: row j spans
// x in [2^-j, 2^-(j+1)] (times 2) and z in +-0.3 x.  Bvh::build splits at the MIDPOINT of the
// longest axis (RAccel.h:319-339), so every split along x peels exactly one row off and the face
// BVH is about levels + log2(across) deep -- far deeper than a balanced tree over as few faces,
// which is what the deep-stack traversal kernels need.  Even rows carry per-vertex normals, odd
// rows are flat shaded.  Geometry is generated in double and rounded once to float.
inline Rayito::Mesh* makeWedgeMesh(unsigned levels, unsigned across)
{
    using namespace Rayito;
    std::vector<Point> verts;
    std::vector<Vector> normals;
    std::vector<Face> faces;
    for (unsigned j = 0; j <= levels; ++j)
    {
        double x = std::ldexp(2.0, -(int)j);
        for (unsigned i = 0; i <= across; ++i)
        {
            double z = x * 0.6 * ((double)i / (double)across - 0.5);
            double y = 0.15 * x * std::sin(1.7 * i + 0.9 * j);
            verts.push_back(Point((float)x, (float)y, (float)z));
            normals.push_back(Vector((float)(0.1 * std::sin(0.8 * i)), 1.0f, (float)(0.1 * std::cos(1.3 * j))).normalized());
        }
    }
    faces.resize((size_t)levels * across);
    size_t f = 0;
    for (unsigned j = 0; j < levels; ++j)
    {
        for (unsigned i = 0; i < across; ++i, ++f)
        {
            unsigned a = j * (across + 1) + i;
            unsigned idx[4] = { a, a + 1, a + (across + 1) + 1, a + (across + 1) };
            for (int k = 0; k < 4; ++k)
            {
                faces[f].m_vertexIndices.push_back(idx[k]);
                if ((j & 1u) == 0)
                    faces[f].m_normalIndices.push_back(idx[k]);
            }
        }
    }
    return new Mesh(verts, normals, faces, NULL);
}

// Deep-tree edge scenes for the parity tests: the wedge mesh (face BVH deeper than 32, so the
// traversal kernels' deep-stack instantiations run) under a rotation key pair, a matte sphere, a
// rectangle light and a moving sphere light.  chain > 0 adds that many spheres whose positions
// and radii halve from one to the next, which makes the TOP-level BVH about `chain` deep as well:
// the two stacks together then need more than 64 entries.
template <typename SetT>
bool buildDeepScene(SetT& set, SceneStore& st, unsigned levels, unsigned across, unsigned chain)
{
    using namespace Rayito;
    if (levels == 0) levels = 40;
    if (across == 0) across = 8;
    Material* ground = st.keep(new DiffuseMaterial(Color(0.6f, 0.6f, 0.9f)));
    Material* glossy = st.keep(new GlossyMaterial(Color(0.8f, 0.5f, 0.2f), 0.3f));
    Material* matte  = st.keep(new DiffuseMaterial(Color(0.3f, 0.8f, 0.4f)));
    Material* mirror = st.keep(new ReflectionMaterial(Color(0.8f, 0.8f, 0.7f)));

    Plane* plane = st.add(new Plane(Point(), Vector(0.0f, 1.0f, 0.0f), ground, true));
    plane->transform().translate(0.0f, Vector(0.0f, -2.0f, 0.0f));
    set.addShape(plane);

    Mesh* wedge = RAYITO_RECIPE_WRAP_MESH(makeWedgeMesh(levels, across));
    st.add(wedge);
    wedge->setMaterial(glossy);
    wedge->transform().setTranslation(0.0f, Vector(-0.5f, -0.5f, 0.0f));
    wedge->transform().rotate(1.0f, Quaternion(Vector(0.0f, 1.0f, 0.0f), M_PI / 8.0f));
    set.addShape(wedge);

    Sphere* side = st.add(new Sphere(Point(), 0.75f, matte));
    side->transform().translate(0.0f, Vector(2.2f, -1.25f, 1.5f));
    set.addShape(side);

    for (unsigned k = 0; k < chain; ++k)
    {
        float s = std::ldexp(1.0f, -(int)k);
        Sphere* bead = st.add(new Sphere(Point(), 0.4f * s, (k & 1u) ? mirror : matte));
        bead->transform().setTranslation(0.0f, Vector(-3.0f + 6.0f * s, 1.5f * s - 1.0f, -1.0f));
        if (k == 2)
            bead->transform().setTranslation(1.0f, Vector(-3.0f + 6.0f * s, 1.5f * s - 0.6f, -1.0f));
        set.addShape(bead);
    }

    RectangleLight* areaLight = st.add(new RectangleLight(Point(),
                                                          Vector(3.0f, 0.0f, 0.0f),
                                                          Vector(0.0f, 0.0f, 3.0f),
                                                          Color(1.0f, 1.0f, 1.0f),
                                                          5.0f));
    areaLight->transform().setTranslation(0.0f, Vector(-1.5f, 4.0f, -1.5f));
    set.addShape(areaLight);

    Sphere* bulb = st.hold(new Sphere(Point(), 0.1f, ground));
    bulb->transform().setTranslation(0.0f, Vector(0.0f, 0.5f, 3.0f));
    bulb->transform().setTranslation(1.0f, Vector(1.0f, 1.5f, 3.0f));
    ShapeLight* sphereLight = st.add(new ShapeLight(bulb, Color(1.0f, 1.0f, 0.3f), 100.0f));
    set.addShape(sphereLight);
    return true;
}

} // namespace rayito_recipes

#endif // RAYITO_B200_SCENE_RECIPES_H
