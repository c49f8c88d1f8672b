// Test driver for integration/rayito_ref_adapter.h: the adapter compiled against the REFERENCE's
// own headers and sources (where they lie under /root/reference, by oracle/Makefile into
// oracle/_ref/libref_adapter.so), with a small C surface for the tests:
//   * adapter_scene_*: build a recipe scene with the reference's classes, let the reference
//     prepare() it, flatten it with the adapter -- tests/test_adapter.py compares the result array
//     for array with what this repo's host library produces for the same recipe (no GPU needed);
//   * adapter_raytrace: rayito_b200_adapter::raytrace(), i.e. reference scene -> C ABI -> B200.
#include "rayito_ref_adapter.h"

#define RAYITO_RECIPE_WRAP_MESH(meshPtr) (meshPtr)
#include "scene_recipes.h"

namespace
{

struct AdapterScene
{
    Rayito::ShapeSet set;
    rayito_recipes::SceneStore store;
    std::vector<Rayito::Shape*> lights;
    rayito_b200_adapter::FlatRefScene flat;
    RtSceneDesc desc;
};

bool buildRecipe(Rayito::ShapeSet& set, rayito_recipes::SceneStore& store, int recipe, const char* obj, unsigned gu, unsigned gv)
{
    switch (recipe)
    {
    case 1: return rayito_recipes::buildStage7Scene1(set, store, obj ? obj : "");
    case 3: return rayito_recipes::buildStage7Scene1(set, store, obj ? obj : "", true);
    case 2: return rayito_recipes::buildStage7Scene2(set, store);
    case 5: return rayito_recipes::buildSyntheticMeshScene(set, store, gu, gv);
    case 7: case 8: case 9: return rayito_recipes::buildEdgeScene(set, store, recipe - 7);
    case 10: return rayito_recipes::buildDeepScene(set, store, gu, gv, 0);
    case 11: return rayito_recipes::buildDeepScene(set, store, gu, gv, 18);
    default: return false;
    }
}

thread_local std::string t_error;

} // namespace

extern "C"
{

const char* adapter_last_error() { return t_error.c_str(); }

void* adapter_scene_create(int recipe, const char* obj, unsigned gu, unsigned gv)
{
    AdapterScene* s = new AdapterScene();
    try
    {
        if (!buildRecipe(s->set, s->store, recipe, obj, gu, gv))
            throw std::runtime_error("scene recipe failed");
        s->set.findLights(s->lights);
        s->set.prepare();
        rayito_b200_adapter::flattenForDevice(s->set, s->lights, s->flat);
        s->desc = s->flat.desc();
        return s;
    }
    catch (const std::exception& e)
    {
        t_error = e.what();
        delete s;
        return NULL;
    }
}

void adapter_scene_destroy(void* h) { delete static_cast<AdapterScene*>(h); }

const RtSceneDesc* adapter_scene_desc(void* h) { return &static_cast<AdapterScene*>(h)->desc; }

int adapter_camera(const float* spec14, RtCamera* out)
{
    Rayito::PerspectiveCamera cam(spec14[0],
                                  Rayito::Point(spec14[1], spec14[2], spec14[3]),
                                  Rayito::Point(spec14[4], spec14[5], spec14[6]),
                                  Rayito::Point(spec14[7], spec14[8], spec14[9]),
                                  spec14[10], spec14[11], spec14[12], spec14[13]);
    rayito_b200_adapter::describeCamera(cam, *out);
    return 0;
}

// The reference application's render call with the adapter in place of Rayito::raytrace()
int adapter_raytrace(int recipe, const char* obj, unsigned gu, unsigned gv, const float* spec14,
                     unsigned width, unsigned height, unsigned ps, unsigned ls, unsigned depth,
                     int device, float* rgb, RtRenderStats* stats)
{
    try
    {
        Rayito::ShapeSet set;
        rayito_recipes::SceneStore store;
        if (!buildRecipe(set, store, recipe, obj, gu, gv))
            throw std::runtime_error("scene recipe failed");
        Rayito::PerspectiveCamera cam(spec14[0],
                                      Rayito::Point(spec14[1], spec14[2], spec14[3]),
                                      Rayito::Point(spec14[4], spec14[5], spec14[6]),
                                      Rayito::Point(spec14[7], spec14[8], spec14[9]),
                                      spec14[10], spec14[11], spec14[12], spec14[13]);
        Rayito::Image* image = rayito_b200_adapter::raytrace(set, cam, width, height, ps, ls, depth, device, stats);
        for (size_t y = 0; y < height; ++y)
            for (size_t x = 0; x < width; ++x)
            {
                const Rayito::Color& c = image->pixel(x, y);
                float* px = rgb + (y * width + x) * 3;
                px[0] = c.m_r; px[1] = c.m_g; px[2] = c.m_b;
            }
        delete image;
        return 0;
    }
    catch (const std::exception& e)
    {
        t_error = e.what();
        return -1;
    }
}

} // extern "C"
