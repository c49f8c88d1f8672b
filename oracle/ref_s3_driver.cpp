// ORACLE (test infrastructure, never shipped, never on the product path).
//
// C-ABI driver around the UNMODIFIED reference Stage 3 program (config C2, the spp
// sweep).  Stage 3 is a single main.cpp + rayito.h with compile-time constants; this
// file #includes that main.cpp WHERE IT LIES (its main() renamed by a macro), so the
// reference's own Rng, makeCameraRay(), trace() and every class of its rayito.h are
// used as they are.  The only restated part is the pixel loop of main()
// (Rayito_Stage3/main.cpp:227-268), because the pixel-sample counts it sweeps are
// constants there; ref3_render(4, 4) is pinned against the unmodified binary
// (oracle/_ref/stage3 -> out.ppm, 0 differing bytes) by tests/test_stage23.py.
// Built by oracle/Makefile into oracle/_ref/libref_s3.so.
#include <cstddef>
#include <cstdint>
#include <cstring>

#define main rayito_stage3_reference_main
#include "main.cpp"          // -I/root/reference/Rayito_Stage3
#undef main

namespace
{

// Counts ShapeSet::intersect calls: every primary and shadow ray of trace()
// (main.cpp:101,138) -- the definition of a "ray" in BASELINE.md.
class CountingSet : public ShapeSet
{
public:
    CountingSet() : m_calls(0) { }
    virtual bool intersect(Intersection& isect)
    {
        ++m_calls;
        return ShapeSet::intersect(isect);
    }
    uint64_t m_calls;
};

} // namespace

extern "C"
{

// The scene and camera of main.cpp:165-201,245-250 rendered with nU x nV stratified
// pixel samples (the reference's constants are 4 x 4) and the reference's fixed 4 x 4
// light samples.  Any output pointer may be NULL.
//   rgb      W*H*3 floats: pixelColor after the box-filter division, before clamp()
//   rgb8     W*H*3 bytes: exactly what the reference streams into out.ppm
//   hitFlags one byte per pixel sample in scan order: did trace() hit anything
//            (this decides how many Rng draws the sample consumed: 2 or 2 + 64)
//   rays     number of ShapeSet::intersect calls
void ref3_render(unsigned width, unsigned height, unsigned nU, unsigned nV,
                 float* rgb, unsigned char* rgb8, unsigned char* hitFlags, uint64_t* rays)
{
    Lambert blueishLambert(Color(0.9f, 0.9f, 1.0f));
    Lambert purplishLambert(Color(0.9f, 0.7f, 0.8f));
    Phong greenishPhong(Color(0.7f, 0.9f, 0.7f), 16.0f);
    CountingSet masterSet;
    Plane plane(Point(0.0f, -2.0f, 0.0f), Vector(0.0f, 1.0f, 0.0f), &blueishLambert, true);
    masterSet.addShape(&plane);
    Sphere sphere1(Point(3.0f, -1.0f, 0.0f), 1.0f, &purplishLambert);
    masterSet.addShape(&sphere1);
    Sphere sphere2(Point(-3.0f, 0.0f, -2.0f), 2.0f, &greenishPhong);
    masterSet.addShape(&sphere2);
    RectangleLight areaLight(Point(-2.5f, 4.0f, -2.5f), Vector(5.0f, 0.0f, 0.0f), Vector(0.0f, 0.0f, 5.0f),
                             Color(1.0f, 1.0f, 1.0f), 1.0f);
    masterSet.addShape(&areaLight);
    Sphere sphereForLight(Point(0.0f, 0.0f, 2.0f), 1.0f, &blueishLambert);
    ShapeLight sphereLight(&sphereForLight, Color(1.0f, 1.0f, 0.1f), 4.0f);
    masterSet.addShape(&sphereLight);

    std::list<Shape*> lights;
    masterSet.findLights(lights);
    Rng rng;

    size_t sample = 0;
    for (size_t y = 0; y < height; ++y)
    {
        for (size_t x = 0; x < width; ++x)
        {
            Color pixelColor(0.0f, 0.0f, 0.0f);
            for (size_t vsi = 0; vsi < nV; ++vsi)
            {
                for (size_t usi = 0; usi < nU; ++usi, ++sample)
                {
                    float yu = 1.0f - ((y + (vsi + rng.nextFloat()) / float(nV)) / float(height));
                    float xu = (x + (usi + rng.nextFloat()) / float(nU)) / float(width);
                    Ray ray = makeCameraRay(45.0f, Point(0.0f, 5.0f, 15.0f), Point(0.0f, 0.0f, 0.0f),
                                            Point(0.0f, 1.0f, 0.0f), xu, yu);
                    Rng before = rng;
                    pixelColor += trace(ray, masterSet, lights, rng);
                    if (hitFlags)
                        hitFlags[sample] = (before.m_z != rng.m_z || before.m_w != rng.m_w) ? 1 : 0;
                }
            }
            pixelColor /= size_t(nU) * size_t(nV);
            if (rgb)
            {
                float* o = rgb + (y * width + x) * 3;
                o[0] = pixelColor.m_r; o[1] = pixelColor.m_g; o[2] = pixelColor.m_b;
            }
            pixelColor.clamp();
            if (rgb8)
            {
                unsigned char* o = rgb8 + (y * width + x) * 3;
                o[0] = static_cast<unsigned char>(pixelColor.m_r * 255.0f);
                o[1] = static_cast<unsigned char>(pixelColor.m_g * 255.0f);
                o[2] = static_cast<unsigned char>(pixelColor.m_b * 255.0f);
            }
        }
    }
    if (rays)
        *rays = masterSet.m_calls;
}

// The reference's compile-time constants, so tests can assert what they pin
void ref3_constants(unsigned* out6)
{
    out6[0] = (unsigned)kWidth; out6[1] = (unsigned)kHeight;
    out6[2] = (unsigned)kNumPixelSamplesU; out6[3] = (unsigned)kNumPixelSamplesV;
    out6[4] = (unsigned)kNumLightSamplesU; out6[5] = (unsigned)kNumLightSamplesV;
}

const char* ref3_build_info()
{
    return "reference: Rayito_Stage3 (unmodified main.cpp + rayito.h; pixel loop restated for the spp sweep), g++ "
           __VERSION__ ", -O3, no -march, no -ffast-math";
}

} // extern "C"
