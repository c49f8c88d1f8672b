// Image, cameras and raytrace(): the top of the Rayito API surface
// (reference: Rayito_Stage7_QT/rayito.h, RaytraceMain.cpp:205-267, 485-579).
//
// raytrace() keeps the reference signature.  It prepares the scene on the host
// exactly as the reference does, flattens it (scene.hpp), uploads it through the
// C ABI (include/rayito_b200.h) and renders on the GPU; there is no CPU renderer
// behind it and it throws std::runtime_error when the CUDA core fails.
#ifndef RAYITO_B200_RENDER_HPP
#define RAYITO_B200_RENDER_HPP

#include <cstddef>
#include <cstring>

#include "scene.hpp"

namespace rayito_b200
{
// Pixel storage of Rayito::Image.  raytrace() hands the caller a new Image per call and the
// caller deletes it after display (MainWindow.cpp:243); a 4K float frame is 100 MB, which the C
// library would map, page-fault in (~25 000 faults) and unmap again on every call.  One spare block
// is kept per process instead, so a render loop reuses the same pages (releaseHostCaches() frees it).
void* acquirePixels(size_t bytes);
void releasePixels(void* block, size_t bytes);
}

namespace Rayito
{

// rayito.h:25-44
class Image
{
public:
    // Black image, as the reference's `new Color[width * height]`
    Image(size_t width, size_t height)
        : m_width(width), m_height(height),
          m_pixels(static_cast<Color*>(rayito_b200::acquirePixels(width * height * sizeof(Color))))
    {
        std::memset(static_cast<void*>(m_pixels), 0, width * height * sizeof(Color));     // 0.0f is all-zero bits
    }
    // raytrace()'s own frames when every pixel is about to be overwritten by the download
    struct Uncleared { };
    Image(size_t width, size_t height, Uncleared)
        : m_width(width), m_height(height),
          m_pixels(static_cast<Color*>(rayito_b200::acquirePixels(width * height * sizeof(Color)))) { }
    virtual ~Image() { rayito_b200::releasePixels(m_pixels, m_width * m_height * sizeof(Color)); }

    size_t width() const { return m_width; }
    size_t height() const { return m_height; }
    Color& pixel(size_t x, size_t y) { return m_pixels[y * m_width + x]; }
    // Contiguous float RGB storage (Color is three floats)
    float* data() { return &m_pixels[0].m_r; }

protected:
    size_t m_width, m_height;
    Color* m_pixels;

private:
    Image(const Image&);
    Image& operator=(const Image&);
};

// rayito.h:51-67
class Camera
{
public:
    Camera(float shutterOpen = 0.0f, float shutterClose = 0.0f)
        : m_shutterOpen(shutterOpen), m_shutterClose(shutterClose) { }
    virtual ~Camera() { }

    // Device description of this camera; false if it has none
    virtual bool describe(RtCamera& out) const = 0;

protected:
    float m_shutterOpen;
    float m_shutterClose;
};

class PerspectiveCamera : public Camera
{
public:
    // Look-at basis and tan(fov) computed as in RaytraceMain.cpp:205-222: the
    // tangent takes the FULL field of view as the half angle and is evaluated in
    // double (M_PI) before rounding to float; right/up are not re-normalised.
    PerspectiveCamera(float fieldOfViewInDegrees,
                      const Point& origin,
                      const Vector& target,
                      const Vector& targetUpDirection,
                      float focalDistance,
                      float lensRadius,
                      float shutterOpen,
                      float shutterClose)
        : Camera(shutterOpen, shutterClose),
          m_origin(origin),
          m_forward((target - origin).normalized()),
          m_tanFov(std::tan(fieldOfViewInDegrees * M_PI / 180.0f)),
          m_focalDistance(focalDistance),
          m_lensRadius(lensRadius)
    {
        m_right = cross(m_forward, targetUpDirection);
        m_up = cross(m_right, m_forward);
    }

    // Stage 6 signature (S6 RaytraceMain.cpp:152-157): no shutter
    PerspectiveCamera(float fieldOfViewInDegrees,
                      const Point& origin,
                      const Vector& target,
                      const Vector& targetUpDirection,
                      float focalDistance,
                      float lensRadius)
        : Camera(0.0f, 0.0f),
          m_origin(origin),
          m_forward((target - origin).normalized()),
          m_tanFov(std::tan(fieldOfViewInDegrees * M_PI / 180.0f)),
          m_focalDistance(focalDistance),
          m_lensRadius(lensRadius)
    {
        m_right = cross(m_forward, targetUpDirection);
        m_up = cross(m_right, m_forward);
    }

    virtual bool describe(RtCamera& out) const
    {
        out.origin[0] = m_origin.m_x; out.origin[1] = m_origin.m_y; out.origin[2] = m_origin.m_z;
        out.forward[0] = m_forward.m_x; out.forward[1] = m_forward.m_y; out.forward[2] = m_forward.m_z;
        out.right[0] = m_right.m_x; out.right[1] = m_right.m_y; out.right[2] = m_right.m_z;
        out.up[0] = m_up.m_x; out.up[1] = m_up.m_y; out.up[2] = m_up.m_z;
        out.tan_fov = m_tanFov;
        out.focal_distance = m_focalDistance;
        out.lens_radius = m_lensRadius;
        out.shutter_open = m_shutterOpen;
        out.shutter_close = m_shutterClose;
        return true;
    }

protected:
    Point m_origin;
    Vector m_forward;
    Vector m_right;
    Vector m_up;
    float m_tanFov;
    float m_focalDistance;
    float m_lensRadius;
};

// rayito.h:138-144.  Caller owns (deletes) the returned Image.
Image* raytrace(ShapeSet& scene,
                const Camera& cam,
                size_t width,
                size_t height,
                unsigned int pixelSamplesHint,
                unsigned int lightSamplesHint,
                unsigned int maxRayDepth);

} // namespace Rayito

namespace rayito_b200
{

// Knobs of the GPU backend behind raytrace(), per calling thread (one host thread drives one
// GPU: a thread's settings never leak into another thread's calls)
struct RenderOptions
{
    int device;              // CUDA device ordinal
    unsigned rank, world;    // screen-tile shard rendered by this process
    unsigned tileSize;       // 0 = core default
    unsigned maxBatchSamples;// 0 = core default
    bool countWork;          // fill node/triangle counters in lastStats()
    RenderOptions() : device(0), rank(0), world(1), tileSize(0), maxBatchSamples(0), countWork(false) { }
};

RenderOptions& renderOptions();

// raytrace() for a multi-GPU application: same preparation, same render of this rank's
// screen tiles (renderOptions().rank / .world), but the float image stays in DEVICE memory:
// deviceImage is width*height*3 floats on renderOptions().device, rows top-down like
// Image::pixel; pixels of other ranks' tiles are left untouched (zero the buffer first).
// The application then assembles the frame with one collective over the device buffers
// (every pixel has exactly one contributing rank, so a sum-reduce is exact) and downloads
// once, instead of every rank staging a full frame through host memory.  INTEGRATION.md.
void raytraceToDevice(Rayito::ShapeSet& scene,
                      const Rayito::Camera& cam,
                      size_t width,
                      size_t height,
                      unsigned int pixelSamplesHint,
                      unsigned int lightSamplesHint,
                      unsigned int maxRayDepth,
                      float* deviceImage);
// raytrace() of ONE frame by all the ranks of a communicator (one process per GPU; rt_comm_create
// / rt_comm_from_nccl in include/rayito_b200.h): every rank prepares the same scene and renders
// its screen tiles, the tiles travel to the root rank over NCCL and are assembled there.  The
// root gets the Image (caller deletes it, as for raytrace()); every other rank gets NULL.  Rank,
// world size and device come from the communicator, not from renderOptions().
Rayito::Image* raytraceMulti(Rayito::ShapeSet& scene,
                             const Rayito::Camera& cam,
                             size_t width,
                             size_t height,
                             unsigned int pixelSamplesHint,
                             unsigned int lightSamplesHint,
                             unsigned int maxRayDepth,
                             RtComm* comm,
                             int root = 0);
// raytrace() keeps its host-side working storage (the flattened scene, BVH build scratch) per
// calling thread between calls; this hands it back to the allocator.  Device memory cached by
// the render core is released by rt_release_cached_memory() (include/rayito_b200.h).
void releaseHostCaches();
FlatScene& detail_flatCache();

// Statistics of the most recent raytrace() on this thread
const RtRenderStats& lastStats();
void detail_setLastStats(const RtRenderStats& stats);

} // namespace rayito_b200

#endif // RAYITO_B200_RENDER_HPP
