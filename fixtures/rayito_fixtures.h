/*
 * rayito_fixtures -- C surface of the test / benchmark fixtures (fixtures/librayito_fixtures.so).
 *
 * Recipe scenes built with the public C++ API of the host library (fixtures/scene_recipes.h:
 * the reference GUI's scenes MainWindow.cpp:139-229 and :289-361, the Stage 6 scene, the
 * synthetic big-mesh scene of BASELINE.json, edge-case scenes), so that tools and tests written
 * in another language can (a) get the flattened RtSceneDesc that rt_scene_create() consumes and
 * (b) call Rayito::raytrace() end to end the way an application does.  Test infrastructure, not
 * product: an application brings its own scene-building code.
 */
#ifndef RAYITO_FIXTURES_H
#define RAYITO_FIXTURES_H

#include "rayito_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum
{
    RTH_RECIPE_STAGE7_SCENE1 = 1,   /* needs obj_path = .../models/bumpy.obj */
    RTH_RECIPE_STAGE7_SCENE2 = 2,
    RTH_RECIPE_STAGE7_SCENE1_MESHLIGHT = 3, /* scene 1 with bumpy.obj as a mesh light (MainWindow.cpp:193-196) */
    RTH_RECIPE_SYNTHETIC_MESH = 5,  /* grid_u x grid_v quads on a displaced sphere */
    RTH_RECIPE_EDGE_LINEAR_LIST = 7, /* edge-case scenes of the parity tests (fixtures/scene_recipes.h buildEdgeScene) */
    RTH_RECIPE_EDGE_NO_LIGHTS = 8,
    RTH_RECIPE_EDGE_EMPTY = 9,
    RTH_RECIPE_EDGE_DEEP_MESH = 10, /* wedge mesh of grid_u rows x grid_v quads (0 = 40 x 8): face BVH ~grid_u + log2(grid_v) deep */
    RTH_RECIPE_EDGE_DEEP_BOTH = 11, /* ... plus a chain of 18 halving spheres: top-level BVH ~18 deep, both stacks > 64 entries */
    RTH_RECIPE_STAGE6_SCENE = 6     /* Stage 6 scene + Stage 6 rules (Rayito_Stage6_QT/MainWindow.cpp:38-146); needs obj_path */
};

typedef struct RthScene RthScene;

/* Build a recipe scene with the C++ API, run findLights() + prepare() (host BVH
 * builds included) and flatten it.  Returns NULL on failure. */
RthScene* rth_scene_create(int recipe, const char* obj_path, unsigned grid_u, unsigned grid_v);
void rth_scene_destroy(RthScene* scene);
/* Flattened scene; valid until rth_scene_destroy(). */
const RtSceneDesc* rth_scene_desc(const RthScene* scene);
double rth_scene_prepare_seconds(const RthScene* scene);
/* Deepest leaf of the top-level BVH (mesh < 0) or of mesh #mesh (root = 0). */
unsigned rth_scene_depth(const RthScene* scene, int mesh);

/* Error text of the last failed fixture call on this thread. */
const char* rthf_last_error_string(void);

/* The camera the reference GUI uses for the recipe, at the UI default settings. */
void rth_scene_default_camera(const RthScene* scene, float* spec14);

/* Build the recipe scene and call Rayito::raytrace() on it (what the GUI's render
 * button does, MainWindow.cpp:232-238).  rgb = width*height*3 floats. */
int rth_raytrace(int recipe, const char* obj_path, unsigned grid_u, unsigned grid_v,
                 const float* spec14, unsigned width, unsigned height,
                 unsigned pixel_samples_hint, unsigned light_samples_hint, unsigned max_ray_depth,
                 int device, unsigned rank, unsigned world, int count_work,
                 float* rgb, RtRenderStats* stats);

/* The same split the way an application is written: the scene-building code runs once
 * (rth_app_create: MainWindow.cpp:143-229 or another recipe; nothing is prepared or
 * flattened), then every rth_app_raytrace() is one Rayito::raytrace() call on that scene
 * -- findLights, prepare() (host BVH build), flatten, upload, render, download -- exactly
 * what the reference redoes per call (RaytraceMain.cpp:494-497).  bench.py's e2e times this
 * call.  rgb_on_device != 0: rgb is a DEVICE pointer (width*height*3 floats on `device`)
 * and the call is rayito_b200::raytraceToDevice(): this rank's tiles stay in HBM for the
 * application's tile-assembly collective. */
typedef struct RthApp RthApp;
RthApp* rth_app_create(int recipe, const char* obj_path, unsigned grid_u, unsigned grid_v);
void rth_app_destroy(RthApp* app);
int rth_app_raytrace(RthApp* app, const float* spec14, unsigned width, unsigned height,
                     unsigned pixel_samples_hint, unsigned light_samples_hint, unsigned max_ray_depth,
                     int device, unsigned rank, unsigned world, int count_work,
                     float* rgb, int rgb_on_device, RtRenderStats* stats);

/* The same call handing back the Image raytrace() returned instead of copying it out:
 * *pixels = width*height*3 floats owned by the app handle, valid until the next
 * rth_app_raytrace_image() / rth_app_destroy() (the GUI keeps its frame the same way and
 * deletes it before the next render, MainWindow.cpp:240-244). */
int rth_app_raytrace_image(RthApp* app, const float* spec14, unsigned width, unsigned height,
                           unsigned pixel_samples_hint, unsigned light_samples_hint, unsigned max_ray_depth,
                           int device, unsigned rank, unsigned world, int count_work,
                           const float** pixels, RtRenderStats* stats);

/* One frame over all ranks of a communicator (rt_comm_create): rayito_b200::raytraceMulti() on the
 * application's scene.  On the root *pixels is the assembled frame (owned by the app handle as
 * above), on every other rank NULL. */
int rth_app_raytrace_multi(RthApp* app, const float* spec14, unsigned width, unsigned height,
                           unsigned pixel_samples_hint, unsigned light_samples_hint, unsigned max_ray_depth,
                           RtComm* comm, int root, const float** pixels, RtRenderStats* stats);

#ifdef __cplusplus
}
#endif

#endif /* RAYITO_FIXTURES_H */
