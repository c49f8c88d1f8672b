// ORACLE (test infrastructure, never shipped, never on the product path).
//
// C-ABI driver around the UNMODIFIED reference Stage 2 program.  As for Stage 3 the
// program's main.cpp is #included WHERE IT LIES (main() renamed), which brings in the
// reference's own Rng, makeCameraRay() and the classes of its rayito.h.  Stage 2 does
// its shading inline in main() (Rayito_Stage2/main.cpp:145-223), so that loop is the
// restated part here, with the pixel-sample count as a parameter; ref2_render(.., 64)
// is pinned against the reference's golden image Rayito_Stage2/out_ref.ppm and against
// the unmodified binary (oracle/_ref/stage2), 0 differing bytes (tests/test_stage23.py).
// Built by oracle/Makefile into oracle/_ref/libref_s2.so.
#include <cstddef>
#include <cstdint>
#include <cstring>

#define main rayito_stage2_reference_main
#include "main.cpp"          // -I/root/reference/Rayito_Stage2
#undef main

namespace
{

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

// Scene of main.cpp:96-122, numSamples purely random samples per pixel (reference: 64).
// Outputs as in ref3_render (ref_s3_driver.cpp).
void ref2_render(unsigned width, unsigned height, unsigned numSamples,
                 float* rgb, unsigned char* rgb8, unsigned char* hitFlags, uint64_t* rays)
{
    CountingSet masterSet;
    Plane plane(Point(0.0f, -2.0f, 0.0f), Vector(0.0f, 1.0f, 0.0f), Color(1.0f, 1.0f, 1.0f), true);
    masterSet.addShape(&plane);
    RectangleLight areaLight(Point(-2.5f, 2.0f, -2.5f), Vector(5.0f, 0.0f, 0.0f), Vector(0.0f, 0.0f, 5.0f),
                             Color(1.0f, 0.5f, 1.0f), 3.0f);
    masterSet.addShape(&areaLight);
    RectangleLight smallAreaLight(Point(-2.0f, -1.0f, -2.0f), Vector(4.0f, 0.0f, 0.0f), Vector(0.0f, 0.0f, 4.0f),
                                  Color(1.0f, 1.0f, 0.5f), 0.75f);
    masterSet.addShape(&smallAreaLight);

    std::list<Shape*> lights;
    masterSet.findLights(lights);
    Rng rng;

    size_t sample = 0;
    for (size_t y = 0; y < height; ++y)
    {
        for (size_t x = 0; x < width; ++x)
        {
            Color pixelColor(0.0f, 0.0f, 0.0f);
            for (size_t si = 0; si < numSamples; ++si, ++sample)
            {
                float yu = 1.0f - ((y + rng.nextFloat()) / float(height - 1));
                float xu = (x + rng.nextFloat()) / float(width - 1);
                Ray ray = makeCameraRay(45.0f, Point(0.0f, 5.0f, 15.0f), Point(0.0f, 0.0f, 0.0f),
                                        Point(0.0f, 1.0f, 0.0f), xu, yu);
                Intersection intersection(ray);
                bool hit = masterSet.intersect(intersection);
                if (hitFlags)
                    hitFlags[sample] = hit ? 1 : 0;
                if (hit)
                {
                    pixelColor += intersection.m_emitted;
                    Point position = intersection.position();
                    for (std::list<Shape*>::iterator iter = lights.begin(); iter != lights.end(); ++iter)
                    {
                        Point lightPoint;
                        Vector lightNormal;
                        Light* pLightShape = dynamic_cast<Light*>(*iter);
                        pLightShape->sampleSurface(rng.nextFloat(),
                                                   rng.nextFloat(),
                                                   position,
                                                   lightPoint,
                                                   lightNormal);
                        Vector toLight = lightPoint - position;
                        float lightDistance = toLight.normalize();
                        Ray shadowRay(position, toLight, lightDistance);
                        Intersection shadowIntersection(shadowRay);
                        bool intersected = masterSet.intersect(shadowIntersection);
                        if (!intersected || shadowIntersection.m_pShape == pLightShape)
                        {
                            float lightAttenuation = std::max(0.0f, dot(intersection.m_normal, toLight));
                            pixelColor += intersection.m_color * pLightShape->emitted() * lightAttenuation;
                        }
                    }
                }
            }
            pixelColor /= size_t(numSamples);
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

void ref2_constants(unsigned* out3)
{
    out3[0] = (unsigned)kWidth; out3[1] = (unsigned)kHeight; out3[2] = (unsigned)kNumPixelSamples;
}

const char* ref2_build_info()
{
    return "reference: Rayito_Stage2 (unmodified rayito.h, Rng and makeCameraRay; main()'s inline shading loop restated), g++ "
           __VERSION__ ", -O3, no -march, no -ffast-math";
}

} // extern "C"
