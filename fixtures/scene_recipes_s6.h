// Scene recipe for the Stage 6 renderer (BASELINE config C3).
//
// Replays the scene the Stage 6 GUI builds (reference:
// Rayito_Stage6_QT/MainWindow.cpp:38-146) through the public Rayito API only.
// Stage 6 has no transforms: every shape is placed through its constructor.
// Like scene_recipes.h this file is API-neutral: it compiles against the
// reference's Stage 6 headers (oracle/_ref/libref_s6.so, the checker) and against
// this repo's host headers (the product).
//
// Include AFTER "rayito.h" and "RMesh.h" of whichever implementation is in use.
#ifndef RAYITO_B200_SCENE_RECIPES_S6_H
#define RAYITO_B200_SCENE_RECIPES_S6_H

#include <vector>

#include "scene_store.h"

namespace rayito_recipes
{

// UI defaults of the Stage 6 GUI (Rayito_Stage6_QT/MainWindow.ui: fov 30, focal
// distance 16, lens radius 0); Stage 6 has no shutter.
inline CameraSpec defaultCameraStage6()
{
    CameraSpec c = { 30.0f, { -2.0f, 5.0f, 15.0f }, { 0.0f, 0.0f, 0.0f }, { 0.0f, 1.0f, 0.0f },
                     16.0f, 0.0f, 0.0f, 0.0f };
    return c;
}

// The hand-written box (MainWindow.cpp:76-120): corners in place, the top face
// listed twice, the bottom face missing, no normals => flat shading.
inline Rayito::Mesh* makeStage6Box(Rayito::Material* material)
{
    static const float corner[8][3] = {
        { 0.0f, -2.0f, -2.0f }, { 1.0f, -2.0f, -2.0f }, { 1.0f, -1.0f, -2.0f }, { 0.0f, -1.0f, -2.0f },
        { 0.0f, -2.0f, -1.0f }, { 1.0f, -2.0f, -1.0f }, { 1.0f, -1.0f, -1.0f }, { 0.0f, -1.0f, -1.0f } };
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
    return new Rayito::Mesh(verts, normals, faces, material);
}

// Bullseye ground plane, four spheres, the box, bumpy.obj, a rectangle light and
// a sphere light.  Insertion order is the reference's and fixes BVH prim indices.
// Returns false if the OBJ mesh could not be read.
template <typename SetT>
bool buildStage6Scene(SetT& set, SceneStore& st, const char* objPath)
{
    using namespace Rayito;
    Material* blueishLambert   = st.keep(new DiffuseMaterial(Color(0.7f, 0.7f, 0.9f)));
    Material* purplishLambert  = st.keep(new DiffuseMaterial(Color(0.8f, 0.3f, 0.7f)));
    Material* yellowishLambert = st.keep(new DiffuseMaterial(Color(0.7f, 0.7f, 0.2f)));
    Material* bluishGlossy     = st.keep(new GlossyMaterial(Color(0.5f, 0.3f, 0.8f), 0.3f));
    Material* greenishGlossy   = st.keep(new GlossyMaterial(Color(0.3f, 0.9f, 0.3f), 0.1f));
    Material* reddishLambert   = st.keep(new DiffuseMaterial(Color(0.8f, 0.3f, 0.1f)));
    Material* reddishGlossy    = st.keep(new GlossyMaterial(Color(0.8f, 0.1f, 0.1f), 0.3f));

    set.addShape(st.add(new Plane(Point(0.0f, -2.0f, 0.0f), Vector(0.0f, 1.0f, 0.0f), blueishLambert, true)));
    set.addShape(st.add(new Sphere(Point(3.0f, -1.0f, 0.0f), 1.0f, purplishLambert)));
    set.addShape(st.add(new Sphere(Point(-3.0f, 0.0f, -2.0f), 2.0f, greenishGlossy)));
    set.addShape(st.add(new Sphere(Point(1.5f, -1.5f, 2.5f), 0.5f, bluishGlossy)));
    set.addShape(st.add(new Sphere(Point(-2.0f, -1.5f, 1.0f), 0.5f, yellowishLambert)));

    Mesh* box = RAYITO_RECIPE_WRAP_MESH(makeStage6Box(reddishLambert));
    set.addShape(st.add(box));

    Mesh* rawObj = createFromOBJFile(objPath);
    if (rawObj == NULL)
        return false;
    Mesh* obj = RAYITO_RECIPE_WRAP_MESH(rawObj);
    st.add(obj);
    obj->setMaterial(reddishGlossy);
    set.addShape(obj);

    set.addShape(st.add(new RectangleLight(Point(-1.5f, 4.0f, -1.5f),
                                           Vector(3.0f, 0.0f, 0.0f),
                                           Vector(0.0f, 0.0f, 3.0f),
                                           Color(1.0f, 1.0f, 1.0f),
                                           5.0f)));

    Sphere* bulb = st.hold(new Sphere(Point(1.0f, 0.5f, 2.0f), 0.5f, blueishLambert));
    set.addShape(st.add(new ShapeLight(bulb, Color(1.0f, 1.0f, 0.3f), 10.0f)));
    return true;
}

} // namespace rayito_recipes

#endif // RAYITO_B200_SCENE_RECIPES_S6_H
