// Host library of the Rayito API surface (librayito_host.so): the OBJ reader,
// raytrace() / raytraceToDevice() / raytraceMulti() and a small C layer
// (include/rayito_b200_host.h): the camera description and the Stage 1-3 programs.
// (The recipe scenes of the tests and of bench.py live in fixtures/, not here.)
//
// Everything that computes a hit or a pixel is in librayito_b200.so (CUDA); this
// file only prepares, flattens and forwards.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "rayito.h"
#include "RMesh.h"
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


// Body of raytrace().  deviceImage == NULL: the reference's contract, a new host Image.
// Otherwise the pixels of this rank's tiles are left in device memory (returns NULL).
static Image* raytraceImpl(ShapeSet& scene,
                           const Camera& cam,
                           size_t width,
                           size_t height,
                           unsigned int pixelSamplesHint,
                           unsigned int lightSamplesHint,
                           unsigned int maxRayDepth,
                           float* deviceImage,
                           RtComm* comm = NULL,
                           int root = 0)
{
    // RAYITO_B200_TIMING=1 prints where a raytrace() call spends its wall time (stderr)
    const bool timing = std::getenv("RAYITO_B200_TIMING") != NULL;
    struct timespec tp[6];
    clock_gettime(CLOCK_MONOTONIC, &tp[0]);
    // Same order as the reference: lights first, then prepare (RaytraceMain.cpp:494-497)
    std::vector<Shape*> lights;
    scene.findLights(lights);
    scene.prepare();
    clock_gettime(CLOCK_MONOTONIC, &tp[1]);

    rayito_b200::FlatScene& flat = rayito_b200::detail_flatCache();
    flat.reset();
    if (!scene.flattenScene(flat, lights))
        throw std::runtime_error("rayito_b200: cannot flatten scene: " + flat.error);
    flat.semantics = rayito_b200::stageSemantics();
    RtCamera camera;
    if (!cam.describe(camera))
        throw std::runtime_error("rayito_b200: camera has no device description");

    rayito_b200::RenderOptions opt = rayito_b200::renderOptions();
    if (comm != NULL)
    {
        int rank = 0, world = 1, device = 0;
        if (rt_comm_rank(comm, &rank, &world, &device) != RT_OK)
            throw std::runtime_error("rayito_b200: bad communicator");
        opt.rank = (unsigned)rank;
        opt.world = (unsigned)world;
        opt.device = device;
    }
    RtSceneDesc desc = flat.desc();
    RtScene* dev = NULL;
    clock_gettime(CLOCK_MONOTONIC, &tp[2]);
    // (kTreeDevice / kTreeAuto: prepare() left the face BVH of some meshes to the GPU; meshes that came with
    // their nodes keep them)
    const unsigned treeMode = rayito_b200::treeMode();
    const bool deviceTrees = (treeMode == rayito_b200::kTreeDevice || treeMode == rayito_b200::kTreeAuto) &&
                             flat.semantics == RT_SEMANTICS_STAGE7;
    int created = rt_scene_create_ex(&desc, opt.device, deviceTrees ? (uint32_t)RT_SCENE_BUILD_MESH_BVH : 0u, &dev);
    if (created == RT_ERR_UNSUPPORTED && deviceTrees)
    {
        // a mesh the device build declines (see rt_build.cuh): build every tree on the host after all
        rayito_b200::treeMode() = rayito_b200::kTreeReference;
        try
        {
            scene.prepare();
            flat.reset();
            if (!scene.flattenScene(flat, lights))
                throw std::runtime_error("rayito_b200: cannot flatten scene: " + flat.error);
            flat.semantics = rayito_b200::stageSemantics();
        }
        catch (...)
        {
            rayito_b200::treeMode() = treeMode;
            throw;
        }
        rayito_b200::treeMode() = treeMode;
        desc = flat.desc();
        created = rt_scene_create(&desc, opt.device, &dev);
    }
    if (created != RT_OK)
        throw std::runtime_error(std::string("rayito_b200: rt_scene_create: ") + rt_last_error_string());
    clock_gettime(CLOCK_MONOTONIC, &tp[3]);

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

    Image* image = NULL;
    RtRenderStats stats;
    int rc;
    if (comm != NULL)
    {
        const bool isRoot = (int)opt.rank == root;
        if (isRoot)
            image = new Image(width, height, Image::Uncleared());
        rc = rt_render_multi_host(dev, &camera, &params, comm, root, isRoot ? image->data() : NULL, &stats, NULL);
    }
    else if (deviceImage != NULL)
        rc = rt_render_device(dev, &camera, &params, deviceImage, &stats, NULL);
    else
    {
        // one rank: every pixel is overwritten by the download; several: the others' tiles stay black
        image = opt.world > 1 ? new Image(width, height) : new Image(width, height, Image::Uncleared());
        rc = rt_render(dev, &camera, &params, image->data(), &stats);
    }
    clock_gettime(CLOCK_MONOTONIC, &tp[4]);
    std::string err = rc == RT_OK ? "" : rt_last_error_string();
    rt_scene_destroy(dev);
    clock_gettime(CLOCK_MONOTONIC, &tp[5]);
    if (timing)
    {
        double ms[5];
        for (int i = 0; i < 5; ++i)
            ms[i] = 1e3 * (double)(tp[i + 1].tv_sec - tp[i].tv_sec) + 1e-6 * (double)(tp[i + 1].tv_nsec - tp[i].tv_nsec);
        std::fprintf(stderr, "[rayito_b200] raytrace: prepare %.1f ms, flatten %.1f, scene upload %.1f, rt_render %.1f "
                             "(device %.1f, image download %.1f), destroy %.1f\n",
                     ms[0], ms[1], ms[2], ms[3], stats.render_ms, stats.download_ms, ms[4]);
    }
    if (rc != RT_OK)
    {
        delete image;
        throw std::runtime_error("rayito_b200: rt_render: " + err);
    }
    rayito_b200::detail_setLastStats(stats);
    return image;
}

Image* raytrace(ShapeSet& scene,
                const Camera& cam,
                size_t width,
                size_t height,
                unsigned int pixelSamplesHint,
                unsigned int lightSamplesHint,
                unsigned int maxRayDepth)
{
    return raytraceImpl(scene, cam, width, height, pixelSamplesHint, lightSamplesHint, maxRayDepth, NULL);
}

} // namespace Rayito


namespace rayito_b200
{

void raytraceToDevice(Rayito::ShapeSet& scene,
                      const Rayito::Camera& cam,
                      size_t width,
                      size_t height,
                      unsigned int pixelSamplesHint,
                      unsigned int lightSamplesHint,
                      unsigned int maxRayDepth,
                      float* deviceImage)
{
    if (deviceImage == NULL)
        throw std::runtime_error("rayito_b200: raytraceToDevice needs a device buffer");
    Rayito::raytraceImpl(scene, cam, width, height, pixelSamplesHint, lightSamplesHint, maxRayDepth, deviceImage);
}

Rayito::Image* raytraceMulti(Rayito::ShapeSet& scene,
                             const Rayito::Camera& cam,
                             size_t width,
                             size_t height,
                             unsigned int pixelSamplesHint,
                             unsigned int lightSamplesHint,
                             unsigned int maxRayDepth,
                             RtComm* comm,
                             int root)
{
    if (comm == NULL)
        throw std::runtime_error("rayito_b200: raytraceMulti needs a communicator");
    return Rayito::raytraceImpl(scene, cam, width, height, pixelSamplesHint, lightSamplesHint, maxRayDepth, NULL, comm, root);
}

namespace
{
std::mutex g_pixelLock;
void* g_sparePixels = NULL;
size_t g_sparePixelBytes = 0;
}

void* acquirePixels(size_t bytes)
{
    if (bytes == 0)
        bytes = sizeof(float);
    {
        std::lock_guard<std::mutex> guard(g_pixelLock);
        if (g_sparePixels != NULL && g_sparePixelBytes == bytes)
        {
            void* block = g_sparePixels;
            g_sparePixels = NULL;
            g_sparePixelBytes = 0;
            return block;
        }
    }
    void* block = std::malloc(bytes);
    if (block == NULL)
        throw std::bad_alloc();
    return block;
}

void releasePixels(void* block, size_t bytes)
{
    if (block == NULL)
        return;
    if (bytes == 0)
        bytes = sizeof(float);
    void* drop = block;
    {
        // keep the larger of the spare and this one
        std::lock_guard<std::mutex> guard(g_pixelLock);
        if (g_sparePixels == NULL || g_sparePixelBytes < bytes)
        {
            drop = g_sparePixels;
            g_sparePixels = block;
            g_sparePixelBytes = bytes;
        }
    }
    std::free(drop);
}

FlatScene& detail_flatCache()
{
    static thread_local FlatScene cache;
    return cache;
}

void releaseHostCaches()
{
    detail_flatCache().shrink();
    buildScratch().release();
    void* spare = NULL;
    {
        std::lock_guard<std::mutex> guard(g_pixelLock);
        spare = g_sparePixels;
        g_sparePixels = NULL;
        g_sparePixelBytes = 0;
    }
    std::free(spare);
}

unsigned& stageSemantics()
{
    static thread_local unsigned semantics = RT_SEMANTICS_STAGE7;
    return semantics;
}

unsigned& treeMode()
{
    static thread_local unsigned mode = kTreeAuto;
    return mode;
}

RenderOptions& renderOptions()
{
    static thread_local RenderOptions options;
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
// C layer of the host library (include/rayito_b200_host.h): camera description and the Stage 1-3 programs
//
namespace
{
thread_local std::string t_hostError;
}

extern "C"
{

const char* rth_last_error_string(void) { return t_hostError.c_str(); }

int rth_set_tree_mode(unsigned mode)
{
    if (mode > rayito_b200::kTreeAuto)
    {
        t_hostError = "rth_set_tree_mode: unknown mode";
        return -1;
    }
    rayito_b200::treeMode() = mode;
    return 0;
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

int rth_stage1_render_float(int device, unsigned width, unsigned height, float* rgb, unsigned char* rgb8)
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
    int rc = rgb != NULL ? rt_stage1_render_float(device, &plane, 1, &cam, width, height, rgb, rgb8)
                         : rt_stage1_render(device, &plane, 1, &cam, width, height, rgb8);
    if (rc != RT_OK)
        t_hostError = rt_last_error_string();
    return rc;
}

int rth_stage1_render(int device, unsigned width, unsigned height, unsigned char* rgb8)
{
    return rth_stage1_render_float(device, width, height, NULL, rgb8);
}

} // extern "C"

namespace
{

// Stage 1-3 vectors normalise by dividing unconditionally (Rayito_Stage3/rayito.h:194)
Rayito::Vector dividedByLength(const Rayito::Vector& v)
{
    float len = v.length();
    return Rayito::Vector(v.m_x / len, v.m_y / len, v.m_z / len);
}

RtS23Shape s23Shape(unsigned type, unsigned material, const Rayito::Point& p)
{
    RtS23Shape s;
    std::memset(&s, 0, sizeof(s));
    s.type = type;
    s.material = material;
    s.position[0] = p.m_x; s.position[1] = p.m_y; s.position[2] = p.m_z;
    return s;
}

RtS23Shape s23Plane(unsigned material, const Rayito::Point& p, const Rayito::Vector& normal, bool bullseye)
{
    RtS23Shape s = s23Shape(RT_S23_PLANE, material, p);
    Rayito::Vector n = dividedByLength(normal);          // Plane ctor: normal.normalized()
    s.normal[0] = n.m_x; s.normal[1] = n.m_y; s.normal[2] = n.m_z;
    s.bullseye = bullseye ? 1 : 0;
    return s;
}

RtS23Shape s23Sphere(unsigned material, const Rayito::Point& p, float radius)
{
    RtS23Shape s = s23Shape(RT_S23_SPHERE, material, p);
    s.radius = radius;
    return s;
}

RtS23Shape s23Rect(unsigned material, const Rayito::Point& p, const Rayito::Vector& a, const Rayito::Vector& b)
{
    RtS23Shape s = s23Shape(RT_S23_RECT, material, p);
    s.side1[0] = a.m_x; s.side1[1] = a.m_y; s.side1[2] = a.m_z;
    s.side2[0] = b.m_x; s.side2[1] = b.m_y; s.side2[2] = b.m_z;
    return s;
}

RtS23Material s23Material(unsigned kind, const Rayito::Color& c, float exponent = 0.0f)
{
    RtS23Material m;
    std::memset(&m, 0, sizeof(m));
    m.kind = kind;
    m.color[0] = c.m_r; m.color[1] = c.m_g; m.color[2] = c.m_b;
    m.exponent = exponent;
    return m;
}

// Emitter(colour, power): emittance() = m_color * m_power (rayito.h:490)
RtS23Material s23Emitter(const Rayito::Color& c, float power)
{
    RtS23Material m = s23Material(RT_S23_MAT_EMITTER, Rayito::Color(0.0f, 0.0f, 0.0f));
    Rayito::Color e = c * power;
    m.emittance[0] = e.m_r; m.emittance[1] = e.m_g; m.emittance[2] = e.m_b;
    return m;
}

} // namespace

extern "C" int rth_stage23_render(int device, int stage, unsigned width, unsigned height, unsigned samples_u,
                                  unsigned samples_v, float* rgb, unsigned char* rgb8, RtRenderStats* stats)
{
    using namespace Rayito;
    std::vector<RtS23Shape> shapes;
    std::vector<RtS23Material> materials;
    std::vector<uint32_t> lights;
    if (stage == 2)
    {
        // Rayito_Stage2/main.cpp:96-122: white bullseye plane, two rectangle lights
        materials.push_back(s23Material(RT_S23_MAT_LAMBERT, Color(1.0f, 1.0f, 1.0f)));
        materials.push_back(s23Emitter(Color(1.0f, 0.5f, 1.0f), 3.0f));
        materials.push_back(s23Emitter(Color(1.0f, 1.0f, 0.5f), 0.75f));
        shapes.push_back(s23Plane(0, Point(0.0f, -2.0f, 0.0f), Vector(0.0f, 1.0f, 0.0f), true));
        shapes.push_back(s23Rect(1, Point(-2.5f, 2.0f, -2.5f), Vector(5.0f, 0.0f, 0.0f), Vector(0.0f, 0.0f, 5.0f)));
        shapes.push_back(s23Rect(2, Point(-2.0f, -1.0f, -2.0f), Vector(4.0f, 0.0f, 0.0f), Vector(0.0f, 0.0f, 4.0f)));
        lights.push_back(1);
        lights.push_back(2);
    }
    else if (stage == 3)
    {
        // Rayito_Stage3/main.cpp:165-201
        materials.push_back(s23Material(RT_S23_MAT_LAMBERT, Color(0.9f, 0.9f, 1.0f)));
        materials.push_back(s23Material(RT_S23_MAT_LAMBERT, Color(0.9f, 0.7f, 0.8f)));
        materials.push_back(s23Material(RT_S23_MAT_PHONG, Color(0.7f, 0.9f, 0.7f), 16.0f));
        materials.push_back(s23Emitter(Color(1.0f, 1.0f, 1.0f), 1.0f));
        materials.push_back(s23Emitter(Color(1.0f, 1.0f, 0.1f), 4.0f));
        shapes.push_back(s23Plane(0, Point(0.0f, -2.0f, 0.0f), Vector(0.0f, 1.0f, 0.0f), true));
        shapes.push_back(s23Sphere(1, Point(3.0f, -1.0f, 0.0f), 1.0f));
        shapes.push_back(s23Sphere(2, Point(-3.0f, 0.0f, -2.0f), 2.0f));
        shapes.push_back(s23Rect(3, Point(-2.5f, 4.0f, -2.5f), Vector(5.0f, 0.0f, 0.0f), Vector(0.0f, 0.0f, 5.0f)));
        shapes.push_back(s23Sphere(4, Point(0.0f, 0.0f, 2.0f), 1.0f));      // ShapeLight around a sphere
        lights.push_back(3);
        lights.push_back(4);
    }
    else
    {
        t_hostError = "stage must be 2 or 3";
        return -1;
    }
    // makeCameraRay's basis (main.cpp:62-67), which does not depend on the pixel
    Point origin(0.0f, 5.0f, 15.0f), target(0.0f, 0.0f, 0.0f), upDir(0.0f, 1.0f, 0.0f);
    Vector forward = dividedByLength(target - origin);
    Vector right = dividedByLength(cross(forward, upDir));
    Vector up = dividedByLength(cross(right, forward));
    RtCamera cam;
    std::memset(&cam, 0, sizeof(cam));
    cam.origin[0] = origin.m_x; cam.origin[1] = origin.m_y; cam.origin[2] = origin.m_z;
    cam.forward[0] = forward.m_x; cam.forward[1] = forward.m_y; cam.forward[2] = forward.m_z;
    cam.right[0] = right.m_x; cam.right[1] = right.m_y; cam.right[2] = right.m_z;
    cam.up[0] = up.m_x; cam.up[1] = up.m_y; cam.up[2] = up.m_z;
    cam.tan_fov = std::tan(45.0f * M_PI / 180.0f);

    RtS23Scene scene;
    scene.num_shapes = (uint32_t)shapes.size();       scene.shapes = &shapes[0];
    scene.num_materials = (uint32_t)materials.size(); scene.materials = &materials[0];
    scene.num_lights = (uint32_t)lights.size();       scene.lights = &lights[0];
    RtS23Params prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.stage = (uint32_t)stage;
    prm.width = width;
    prm.height = height;
    prm.pixel_samples_u = samples_u;
    prm.pixel_samples_v = stage == 2 ? 1 : samples_v;
    prm.light_samples_u = prm.light_samples_v = 4;    // kNumLightSamplesU/V (Rayito_Stage3/main.cpp:92-93)
    prm.seed_z = 362436069u;
    prm.seed_w = 521288629u;
    int rc = rt_stage23_render(device, &scene, &cam, &prm, rgb, rgb8, stats);
    if (rc != RT_OK)
        t_hostError = rt_last_error_string();
    return rc;
}
