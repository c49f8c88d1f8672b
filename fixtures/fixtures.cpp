// Test and benchmark fixtures (librayito_fixtures.so): the recipe scenes -- the reference GUI's
// two scenes, the Stage 6 scene, the synthetic 10 M-triangle mesh and the edge-case scenes --
// built with the PUBLIC Rayito API of the host library, behind a small C surface
// (fixtures/rayito_fixtures.h) for tools written in another language.  Not part of the product:
// an application brings its own scene-building code.  The same recipe headers are
// compiled against the reference's own classes by the oracle, which is the drop-in proof for the
// C++ surface.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <string>
#include <vector>

#include "rayito.h"
#include "RMesh.h"
#include "scene_recipes.h"
#include "scene_recipes_s6.h"
#include "rayito_fixtures.h"

struct RthScene
{
    Rayito::ShapeSet set;
    rayito_recipes::SceneStore store;
    std::vector<Rayito::Shape*> lights;
    rayito_b200::FlatScene flat;
    RtSceneDesc desc;
    rayito_recipes::CameraSpec cameraSpec;
    double prepareSeconds;
};

// An application's scene as its scene-building code leaves it: NOT prepared, NOT flattened
struct RthApp
{
    Rayito::ShapeSet set;
    rayito_recipes::SceneStore store;
    unsigned semantics;
    Rayito::Image* frame;       // the Image of the last rth_app_raytrace_image(), owned like the GUI owns its frame
    RthApp() : semantics(RT_SEMANTICS_STAGE7), frame(NULL) { }
    ~RthApp() { delete frame; }
};

namespace
{
thread_local std::string t_fixtureError;

// The recipe entry points pick the stage rules per call and restore them afterwards
struct StageScope
{
    unsigned saved;
    explicit StageScope(unsigned semantics) : saved(rayito_b200::stageSemantics()) { rayito_b200::stageSemantics() = semantics; }
    ~StageScope() { rayito_b200::stageSemantics() = saved; }
};


// One place that maps a recipe id to its scene-building code (the same switch serves
// rth_scene_create, rth_raytrace and rth_app_create)
bool buildRecipe(Rayito::ShapeSet& set, rayito_recipes::SceneStore& store, int recipe, const char* obj_path,
                        unsigned grid_u, unsigned grid_v, rayito_recipes::CameraSpec* camera)
{
    const char* obj = obj_path ? obj_path : "";
    rayito_recipes::CameraSpec cam = rayito_recipes::defaultCameraScene1();
    bool built = false;
    switch (recipe)
    {
    case RTH_RECIPE_STAGE6_SCENE:
        cam = rayito_recipes::defaultCameraStage6();
        built = rayito_recipes::buildStage6Scene(set, store, obj);
        break;
    case RTH_RECIPE_STAGE7_SCENE1: built = rayito_recipes::buildStage7Scene1(set, store, obj); break;
    case RTH_RECIPE_STAGE7_SCENE1_MESHLIGHT: built = rayito_recipes::buildStage7Scene1(set, store, obj, true); break;
    case RTH_RECIPE_STAGE7_SCENE2:
        cam = rayito_recipes::defaultCameraScene2();
        built = rayito_recipes::buildStage7Scene2(set, store);
        break;
    case RTH_RECIPE_SYNTHETIC_MESH: built = rayito_recipes::buildSyntheticMeshScene(set, store, grid_u, grid_v); break;
    case RTH_RECIPE_EDGE_LINEAR_LIST: case RTH_RECIPE_EDGE_NO_LIGHTS: case RTH_RECIPE_EDGE_EMPTY:
        built = rayito_recipes::buildEdgeScene(set, store, recipe - RTH_RECIPE_EDGE_LINEAR_LIST);
        break;
    case RTH_RECIPE_EDGE_DEEP_MESH: built = rayito_recipes::buildDeepScene(set, store, grid_u, grid_v, 0); break;
    case RTH_RECIPE_EDGE_DEEP_BOTH: built = rayito_recipes::buildDeepScene(set, store, grid_u, grid_v, 18); break;
    default: break;
    }
    if (camera) *camera = cam;
    return built;
}

RthScene* finish(RthScene* s, bool built)
{
    if (!built)
    {
        t_fixtureError = "scene recipe failed (could not read the OBJ mesh?)";
        delete s;
        return NULL;
    }
    s->flat.semantics = rayito_b200::stageSemantics();
    s->set.findLights(s->lights);
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    s->set.prepare();
    clock_gettime(CLOCK_MONOTONIC, &b);
    s->prepareSeconds = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
    if (!s->set.flattenScene(s->flat, s->lights))
    {
        t_fixtureError = "flatten failed: " + s->flat.error;
        delete s;
        return NULL;
    }
    if (std::getenv("RAYITO_B200_TIMING") != NULL)
    {
        struct timespec c;
        clock_gettime(CLOCK_MONOTONIC, &c);
        std::fprintf(stderr, "[rayito_b200] rth_scene_create: prepare %.1f ms, flatten %.1f ms\n", 1e3 * s->prepareSeconds,
                     1e3 * (double)(c.tv_sec - b.tv_sec) + 1e-6 * (double)(c.tv_nsec - b.tv_nsec));
    }
    s->desc = s->flat.desc();
    return s;
}
}

extern "C"
{

RthScene* rth_scene_create(int recipe, const char* obj_path, unsigned grid_u, unsigned grid_v)
{
    RthScene* s = new RthScene();
    StageScope stage(recipe == RTH_RECIPE_STAGE6_SCENE ? RT_SEMANTICS_STAGE6 : RT_SEMANTICS_STAGE7);
    return finish(s, buildRecipe(s->set, s->store, recipe, obj_path, grid_u, grid_v, &s->cameraSpec));
}

void rth_scene_destroy(RthScene* s) { delete s; }

const RtSceneDesc* rth_scene_desc(const RthScene* s) { return &s->desc; }

double rth_scene_prepare_seconds(const RthScene* s) { return s->prepareSeconds; }

unsigned rth_scene_depth(const RthScene* s, int mesh)
{
    if (mesh < 0) return s->flat.topDepth;
    return (size_t)mesh < s->flat.meshDepth.size() ? s->flat.meshDepth[mesh] : 0;
}

void rth_scene_default_camera(const RthScene* s, float* spec14)
{
    const rayito_recipes::CameraSpec& c = s->cameraSpec;
    spec14[0] = c.fov;
    for (int i = 0; i < 3; ++i) { spec14[1 + i] = c.origin[i]; spec14[4 + i] = c.target[i]; spec14[7 + i] = c.up[i]; }
    spec14[10] = c.focalDistance; spec14[11] = c.lensRadius; spec14[12] = c.shutterOpen; spec14[13] = c.shutterClose;
}

int rth_raytrace(int recipe, const char* obj_path, unsigned grid_u, unsigned grid_v,
                 const float* spec14, unsigned width, unsigned height,
                 unsigned ps, unsigned ls, unsigned depth,
                 int device, unsigned rank, unsigned world, int count_work,
                 float* rgb, RtRenderStats* stats)
{
    // Mirrors MainWindow::on_renderButton_clicked: build, raytrace, hand back pixels
    try
    {
        Rayito::ShapeSet set;
        rayito_recipes::SceneStore store;
        StageScope stage(recipe == RTH_RECIPE_STAGE6_SCENE ? RT_SEMANTICS_STAGE6 : RT_SEMANTICS_STAGE7);
        bool built = buildRecipe(set, store, recipe, obj_path, grid_u, grid_v, NULL);
        if (!built)
        {
            t_fixtureError = "scene recipe failed";
            return -1;
        }
        Rayito::PerspectiveCamera cam(spec14[0],
                                      Rayito::Point(spec14[1], spec14[2], spec14[3]),
                                      Rayito::Point(spec14[4], spec14[5], spec14[6]),
                                      Rayito::Point(spec14[7], spec14[8], spec14[9]),
                                      spec14[10], spec14[11], spec14[12], spec14[13]);
        rayito_b200::RenderOptions& opt = rayito_b200::renderOptions();
        opt.device = device;
        opt.rank = rank;
        opt.world = world ? world : 1;
        opt.countWork = count_work != 0;
        Rayito::Image* image = Rayito::raytrace(set, cam, width, height, ps, ls, depth);
        std::memcpy(rgb, image->data(), (size_t)width * height * 3 * sizeof(float));
        delete image;
        if (stats) *stats = rayito_b200::lastStats();
        return 0;
    }
    catch (const std::exception& e)
    {
        t_fixtureError = e.what();
        return -1;
    }
}

RthApp* rth_app_create(int recipe, const char* obj_path, unsigned grid_u, unsigned grid_v)
{
    RthApp* app = new RthApp();
    app->semantics = recipe == RTH_RECIPE_STAGE6_SCENE ? RT_SEMANTICS_STAGE6 : RT_SEMANTICS_STAGE7;
    StageScope stage(app->semantics);
    bool built = buildRecipe(app->set, app->store, recipe, obj_path, grid_u, grid_v, NULL);
    if (!built)
    {
        t_fixtureError = "scene recipe failed (unknown recipe, or the OBJ mesh could not be read)";
        delete app;
        return NULL;
    }
    return app;
}

void rth_app_destroy(RthApp* app) { delete app; }

int rth_app_raytrace(RthApp* app, const float* spec14, unsigned width, unsigned height,
                     unsigned ps, unsigned ls, unsigned depth,
                     int device, unsigned rank, unsigned world, int count_work,
                     float* rgb, int rgb_on_device, RtRenderStats* stats)
{
    if (app == NULL || spec14 == NULL || rgb == NULL)
    {
        t_fixtureError = "null argument";
        return -1;
    }
    try
    {
        StageScope stage(app->semantics);
        Rayito::PerspectiveCamera cam(spec14[0],
                                      Rayito::Point(spec14[1], spec14[2], spec14[3]),
                                      Rayito::Point(spec14[4], spec14[5], spec14[6]),
                                      Rayito::Point(spec14[7], spec14[8], spec14[9]),
                                      spec14[10], spec14[11], spec14[12], spec14[13]);
        rayito_b200::RenderOptions& opt = rayito_b200::renderOptions();
        opt.device = device;
        opt.rank = rank;
        opt.world = world ? world : 1;
        opt.countWork = count_work != 0;
        if (rgb_on_device)
            rayito_b200::raytraceToDevice(app->set, cam, width, height, ps, ls, depth, rgb);
        else
        {
            Rayito::Image* image = Rayito::raytrace(app->set, cam, width, height, ps, ls, depth);
            std::memcpy(rgb, image->data(), (size_t)width * height * 3 * sizeof(float));
            delete image;
        }
        if (stats) *stats = rayito_b200::lastStats();
        return 0;
    }
    catch (const std::exception& e)
    {
        t_fixtureError = e.what();
        return -1;
    }
}

int rth_app_raytrace_image(RthApp* app, const float* spec14, unsigned width, unsigned height,
                           unsigned ps, unsigned ls, unsigned depth,
                           int device, unsigned rank, unsigned world, int count_work,
                           const float** pixels, RtRenderStats* stats)
{
    if (app == NULL || spec14 == NULL || pixels == NULL)
    {
        t_fixtureError = "null argument";
        return -1;
    }
    try
    {
        // the application drops the previous frame before asking for the next one (MainWindow.cpp:243)
        delete app->frame;
        app->frame = NULL;
        *pixels = NULL;
        StageScope stage(app->semantics);
        Rayito::PerspectiveCamera cam(spec14[0],
                                      Rayito::Point(spec14[1], spec14[2], spec14[3]),
                                      Rayito::Point(spec14[4], spec14[5], spec14[6]),
                                      Rayito::Point(spec14[7], spec14[8], spec14[9]),
                                      spec14[10], spec14[11], spec14[12], spec14[13]);
        rayito_b200::RenderOptions& opt = rayito_b200::renderOptions();
        opt.device = device;
        opt.rank = rank;
        opt.world = world ? world : 1;
        opt.countWork = count_work != 0;
        app->frame = Rayito::raytrace(app->set, cam, width, height, ps, ls, depth);
        *pixels = app->frame->data();
        if (stats) *stats = rayito_b200::lastStats();
        return 0;
    }
    catch (const std::exception& e)
    {
        t_fixtureError = e.what();
        return -1;
    }
}

int rth_app_raytrace_multi(RthApp* app, const float* spec14, unsigned width, unsigned height,
                            unsigned ps, unsigned ls, unsigned depth, RtComm* comm, int root,
                            const float** pixels, RtRenderStats* stats)
{
    if (app == NULL || spec14 == NULL || comm == NULL || pixels == NULL)
    {
        t_fixtureError = "null argument";
        return -1;
    }
    try
    {
        delete app->frame;
        app->frame = NULL;
        *pixels = NULL;
        StageScope stage(app->semantics);
        Rayito::PerspectiveCamera cam(spec14[0],
                                      Rayito::Point(spec14[1], spec14[2], spec14[3]),
                                      Rayito::Point(spec14[4], spec14[5], spec14[6]),
                                      Rayito::Point(spec14[7], spec14[8], spec14[9]),
                                      spec14[10], spec14[11], spec14[12], spec14[13]);
        app->frame = rayito_b200::raytraceMulti(app->set, cam, width, height, ps, ls, depth, comm, root);
        if (app->frame != NULL)
            *pixels = app->frame->data();
        if (stats) *stats = rayito_b200::lastStats();
        return 0;
    }
    catch (const std::exception& e)
    {
        t_fixtureError = e.what();
        return -1;
    }
}

const char* rthf_last_error_string(void) { return t_fixtureError.c_str(); }

} // extern "C"
