// Host library of the Rayito API surface (librayito_host.so): the OBJ reader,
// raytrace() and a small C layer (include/rayito_b200_host.h) that lets tools and
// tests build the recipe scenes and obtain their flattened description.
//
// Everything that computes a hit or a pixel is in librayito_b200.so (CUDA); this
// file only prepares, flattens and forwards.
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
#include "rayito_b200_host.h"

namespace Rayito
{

namespace
{

bool isBlank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

const char* skipBlank(const char* p) { while (*p && isBlank(*p)) ++p; return p; }

// Parse like `istream >> float`: skip blanks, strtof; a failed extraction yields
// 0 and poisons the rest of the line (the stream would be in a failed state).
float readFloat(const char*& p, bool& ok)
{
    if (!ok) return 0.0f;
    p = skipBlank(p);
    char* end = NULL;
    float v = std::strtof(p, &end);
    if (end == p) { ok = false; return 0.0f; }
    p = end;
    return v;
}

// Parse like `istream >> int`
bool readInt(const char*& p, int& out)
{
    p = skipBlank(p);
    char* end = NULL;
    long v = std::strtol(p, &end, 10);
    if (end == p) return false;
    p = end;
    out = (int)v;
    return true;
}

} // namespace

// Reads `v`, `vn` and `f` records (forms a, a/b, a//c, a/b/c; 1-based, negative =
// relative to the end); everything else is skipped.  Behaviour follows the
// reference reader (OBJMesh.cpp:49-181) record for record: missing coordinates
// read as 0, a face's index list ends at the first token that is not an integer
// (bumpy.obj has trailing blanks), out-of-range indices are reported on stderr.
Mesh* createFromOBJFile(const char* filename)
{
    FILE* fp = std::fopen(filename, "rb");
    if (fp == NULL)
        return NULL;
    std::string text;
    char chunk[1 << 16];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof(chunk), fp)) > 0)
        text.append(chunk, got);
    std::fclose(fp);

    std::vector<Point> verts;
    std::vector<Vector> normals;
    std::vector<Face> faces;

    size_t pos = 0;
    std::string line;
    while (pos < text.size())
    {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        line.assign(text, pos, eol - pos);
        pos = eol + 1;

        const char* p = skipBlank(line.c_str());
        const char* cmdEnd = p;
        while (*cmdEnd && !isBlank(*cmdEnd)) ++cmdEnd;
        size_t cmdLen = (size_t)(cmdEnd - p);
        if (cmdLen == 0 || p[0] == '#')
            continue;
        bool isV = cmdLen == 1 && p[0] == 'v';
        bool isVN = cmdLen == 2 && p[0] == 'v' && p[1] == 'n';
        bool isF = cmdLen == 1 && p[0] == 'f';
        p = cmdEnd;
        if (isV || isVN)
        {
            bool ok = true;
            float x = readFloat(p, ok);
            float y = readFloat(p, ok);
            float z = readFloat(p, ok);
            if (isV) verts.push_back(Point(x, y, z));
            else normals.push_back(Vector(x, y, z));
        }
        else if (isF)
        {
            faces.push_back(Face());
            Face& face = faces.back();
            for (;;)
            {
                int vi;
                if (!readInt(p, vi))
                    break;
                int ni = 0, uvi = 0;
                bool gotN = false;
                bool streamOk = true;
                if (*p == '/')
                {
                    ++p;
                    if (*p == '/')
                    {
                        ++p;
                        streamOk = readInt(p, ni);
                        gotN = true;
                    }
                    else
                    {
                        streamOk = readInt(p, uvi);
                        if (streamOk && *p == '/')
                        {
                            ++p;
                            streamOk = readInt(p, ni);
                            gotN = true;
                        }
                    }
                }
                vi = vi > 0 ? vi - 1 : (int)verts.size() + vi;
                face.m_vertexIndices.push_back(vi);
                if (vi >= (int)verts.size())
                    std::fprintf(stderr, "Found out-of-range vertex index: %d\n", vi);
                if (gotN)
                {
                    ni = ni > 0 ? ni - 1 : (int)normals.size() + ni;
                    face.m_normalIndices.push_back(ni);
                    if (ni >= (int)normals.size())
                        std::fprintf(stderr, "Found out-of-range N index: %d\n", ni);
                }
                if (!streamOk)
                    break;
            }
        }
    }
    if (verts.empty() || faces.empty())
        return NULL;
    return new Mesh(verts, normals, faces, NULL);
}


Image* raytrace(ShapeSet& scene,
                const Camera& cam,
                size_t width,
                size_t height,
                unsigned int pixelSamplesHint,
                unsigned int lightSamplesHint,
                unsigned int maxRayDepth)
{
    // Same order as the reference: lights first, then prepare (RaytraceMain.cpp:494-497)
    std::vector<Shape*> lights;
    scene.findLights(lights);
    scene.prepare();

    rayito_b200::FlatScene flat;
    if (!scene.flattenScene(flat, lights))
        throw std::runtime_error("rayito_b200: cannot flatten scene: " + flat.error);
    flat.semantics = rayito_b200::stageSemantics();
    RtCamera camera;
    if (!cam.describe(camera))
        throw std::runtime_error("rayito_b200: camera has no device description");

    const rayito_b200::RenderOptions& opt = rayito_b200::renderOptions();
    RtSceneDesc desc = flat.desc();
    RtScene* dev = NULL;
    if (rt_scene_create(&desc, opt.device, &dev) != RT_OK)
        throw std::runtime_error(std::string("rayito_b200: rt_scene_create: ") + rt_last_error_string());

    RtRenderParams params;
    std::memset(&params, 0, sizeof(params));
    params.width = (uint32_t)width;
    params.height = (uint32_t)height;
    params.pixel_samples_hint = pixelSamplesHint;
    params.light_samples_hint = lightSamplesHint;
    params.max_ray_depth = maxRayDepth;
    params.tile_size = opt.tileSize;
    params.rank = opt.rank;
    params.world = opt.world;
    params.max_batch_samples = opt.maxBatchSamples;
    params.flags = opt.countWork ? RT_RENDER_COUNT_WORK : 0;

    Image* image = new Image(width, height);
    std::memset(image->data(), 0, width * height * 3 * sizeof(float));
    RtRenderStats stats;
    int rc = rt_render(dev, &camera, &params, image->data(), &stats);
    std::string err = rc == RT_OK ? "" : rt_last_error_string();
    rt_scene_destroy(dev);
    if (rc != RT_OK)
    {
        delete image;
        throw std::runtime_error("rayito_b200: rt_render: " + err);
    }
    rayito_b200::detail_setLastStats(stats);
    return image;
}

} // namespace Rayito


namespace rayito_b200
{

unsigned& stageSemantics()
{
    static unsigned semantics = RT_SEMANTICS_STAGE7;
    return semantics;
}

RenderOptions& renderOptions()
{
    static RenderOptions options;
    return options;
}

namespace
{
thread_local RtRenderStats t_lastStats;
}

void detail_setLastStats(const RtRenderStats& s) { t_lastStats = s; }
const RtRenderStats& lastStats() { return t_lastStats; }

} // namespace rayito_b200


//
// C layer for tools and tests
//
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

namespace
{
thread_local std::string t_hostError;

// The recipe entry points pick the stage rules per call and restore them afterwards
struct StageScope
{
    unsigned saved;
    explicit StageScope(unsigned semantics) : saved(rayito_b200::stageSemantics()) { rayito_b200::stageSemantics() = semantics; }
    ~StageScope() { rayito_b200::stageSemantics() = saved; }
};

RthScene* finish(RthScene* s, bool built)
{
    if (!built)
    {
        t_hostError = "scene recipe failed (could not read the OBJ mesh?)";
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
        t_hostError = "flatten failed: " + s->flat.error;
        delete s;
        return NULL;
    }
    s->desc = s->flat.desc();
    return s;
}
}

extern "C"
{

const char* rth_last_error_string(void) { return t_hostError.c_str(); }

RthScene* rth_scene_create(int recipe, const char* obj_path, unsigned grid_u, unsigned grid_v)
{
    RthScene* s = new RthScene();
    bool built = false;
    StageScope stage(recipe == RTH_RECIPE_STAGE6_SCENE ? RT_SEMANTICS_STAGE6 : RT_SEMANTICS_STAGE7);
    switch (recipe)
    {
    case RTH_RECIPE_STAGE6_SCENE:
        s->cameraSpec = rayito_recipes::defaultCameraStage6();
        built = rayito_recipes::buildStage6Scene(s->set, s->store, obj_path ? obj_path : "");
        break;
    case RTH_RECIPE_STAGE7_SCENE1:
        s->cameraSpec = rayito_recipes::defaultCameraScene1();
        built = rayito_recipes::buildStage7Scene1(s->set, s->store, obj_path ? obj_path : "");
        break;
    case RTH_RECIPE_STAGE7_SCENE1_MESHLIGHT:
        s->cameraSpec = rayito_recipes::defaultCameraScene1();
        built = rayito_recipes::buildStage7Scene1(s->set, s->store, obj_path ? obj_path : "", true);
        break;
    case RTH_RECIPE_STAGE7_SCENE2:
        s->cameraSpec = rayito_recipes::defaultCameraScene2();
        built = rayito_recipes::buildStage7Scene2(s->set, s->store);
        break;
    case RTH_RECIPE_SYNTHETIC_MESH:
        s->cameraSpec = rayito_recipes::defaultCameraScene1();
        built = rayito_recipes::buildSyntheticMeshScene(s->set, s->store, grid_u, grid_v);
        break;
    default:
        t_hostError = "unknown recipe";
        delete s;
        return NULL;
    }
    return finish(s, built);
}

void rth_scene_destroy(RthScene* s) { delete s; }

const RtSceneDesc* rth_scene_desc(const RthScene* s) { return &s->desc; }

double rth_scene_prepare_seconds(const RthScene* s) { return s->prepareSeconds; }

unsigned rth_scene_depth(const RthScene* s, int mesh)
{
    if (mesh < 0) return s->flat.topDepth;
    return (size_t)mesh < s->flat.meshDepth.size() ? s->flat.meshDepth[mesh] : 0;
}

int rth_camera(const float* spec14, RtCamera* out)
{
    Rayito::PerspectiveCamera cam(spec14[0],
                                  Rayito::Point(spec14[1], spec14[2], spec14[3]),
                                  Rayito::Point(spec14[4], spec14[5], spec14[6]),
                                  Rayito::Point(spec14[7], spec14[8], spec14[9]),
                                  spec14[10], spec14[11], spec14[12], spec14[13]);
    return cam.describe(*out) ? 0 : -1;
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
        bool built = false;
        StageScope stage(recipe == RTH_RECIPE_STAGE6_SCENE ? RT_SEMANTICS_STAGE6 : RT_SEMANTICS_STAGE7);
        switch (recipe)
        {
        case RTH_RECIPE_STAGE6_SCENE: built = rayito_recipes::buildStage6Scene(set, store, obj_path ? obj_path : ""); break;
        case RTH_RECIPE_STAGE7_SCENE1: built = rayito_recipes::buildStage7Scene1(set, store, obj_path ? obj_path : ""); break;
        case RTH_RECIPE_STAGE7_SCENE1_MESHLIGHT: built = rayito_recipes::buildStage7Scene1(set, store, obj_path ? obj_path : "", true); break;
        case RTH_RECIPE_STAGE7_SCENE2: built = rayito_recipes::buildStage7Scene2(set, store); break;
        case RTH_RECIPE_SYNTHETIC_MESH: built = rayito_recipes::buildSyntheticMeshScene(set, store, grid_u, grid_v); break;
        default: break;
        }
        if (!built)
        {
            t_hostError = "scene recipe failed";
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
        t_hostError = e.what();
        return -1;
    }
}

int rth_stage1_render(int device, unsigned width, unsigned height, unsigned char* rgb8)
{
    using namespace Rayito;
    // Scene of Rayito_Stage1/main.cpp:68-74
    RtStage1Plane plane;
    Vector n = Vector(0.0f, 1.0f, 0.0f);
    {   // Stage 1's normalize divides unconditionally (rayito.h:194)
        float len = n.length();
        n = Vector(n.m_x / len, n.m_y / len, n.m_z / len);
    }
    plane.position[0] = 0.0f; plane.position[1] = -2.0f; plane.position[2] = 0.0f;
    plane.normal[0] = n.m_x; plane.normal[1] = n.m_y; plane.normal[2] = n.m_z;
    plane.color[0] = 1.0f; plane.color[1] = 0.5f; plane.color[2] = 0.8f;
    // makeCameraRay's per-call basis (main.cpp:35-40), hoisted: it does not depend on the pixel
    Point origin(0.0f, 0.0f, 0.0f), target(0.0f, 0.0f, 1.0f), upDir(0.0f, 1.0f, 0.0f);
    Vector forward = target - origin;
    float fl = forward.length();
    forward = Vector(forward.m_x / fl, forward.m_y / fl, forward.m_z / fl);
    Vector right = cross(forward, upDir);
    float rl = right.length();
    right = Vector(right.m_x / rl, right.m_y / rl, right.m_z / rl);
    Vector up = cross(right, forward);
    float ul = up.length();
    up = Vector(up.m_x / ul, up.m_y / ul, up.m_z / ul);
    RtCamera cam;
    std::memset(&cam, 0, sizeof(cam));
    cam.origin[0] = origin.m_x; cam.origin[1] = origin.m_y; cam.origin[2] = origin.m_z;
    cam.forward[0] = forward.m_x; cam.forward[1] = forward.m_y; cam.forward[2] = forward.m_z;
    cam.right[0] = right.m_x; cam.right[1] = right.m_y; cam.right[2] = right.m_z;
    cam.up[0] = up.m_x; cam.up[1] = up.m_y; cam.up[2] = up.m_z;
    cam.tan_fov = std::tan(30.0f * M_PI / 180.0f);
    int rc = rt_stage1_render(device, &plane, 1, &cam, width, height, rgb8);
    if (rc != RT_OK)
        t_hostError = rt_last_error_string();
    return rc;
}

} // extern "C"
