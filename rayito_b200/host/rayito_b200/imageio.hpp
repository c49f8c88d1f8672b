// Image file writers of the tutorial's command-line stages.
//
// Stages 1-4 stream their image into out.ppm while they render
// (Rayito_Stage1/main.cpp:78-90, 126-133; Rayito_Stage3/main.cpp:211-225, 259-271): header
// "P6\n<W> <H>\n255\n", then one byte per channel, rows top-down, each channel clamped to
// [0,1] and truncated (not rounded) to 8 bits.  Built with WRITE_PFM they write out.pfm
// instead: header "PF\n<W> <H>\n-1.0\n" followed by `fileStream << r << g << b`, which is
// ostream's TEXT formatting of each float (six significant digits, no separators, rows
// top-down) -- not a readable PFM.  Both forms are reproduced byte for byte (writePPM,
// writeReferencePFM), and writePFM writes the standard binary PFM (little-endian floats,
// rows bottom-up) that image tools can read.
#ifndef RAYITO_B200_IMAGEIO_HPP
#define RAYITO_B200_IMAGEIO_HPP

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace rayito_b200
{

// Color::clamp() then static_cast<unsigned char>(c * 255.0f)  (Rayito_Stage1/main.cpp:126-132)
inline void quantise8(const float* rgb, size_t numPixels, unsigned char* rgb8)
{
    for (size_t i = 0; i < numPixels * 3; ++i)
    {
        float c = rgb[i];
        c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);       // std::max(minValue, std::min(maxValue, c)); NaN stays NaN
        rgb8[i] = static_cast<unsigned char>(c * 255.0f);
    }
}

inline bool writePPM(const char* path, size_t width, size_t height, const unsigned char* rgb8)
{
    FILE* fp = std::fopen(path, "wb");
    if (fp == NULL)
        return false;
    std::fprintf(fp, "P6\n%zu %zu\n255\n", width, height);
    size_t n = width * height * 3;
    bool ok = std::fwrite(rgb8, 1, n, fp) == n;
    return std::fclose(fp) == 0 && ok;
}

// Standard PFM: "PF", negative scale = little-endian, rows from the BOTTOM of the image up
inline bool writePFM(const char* path, size_t width, size_t height, const float* rgb)
{
    FILE* fp = std::fopen(path, "wb");
    if (fp == NULL)
        return false;
    std::fprintf(fp, "PF\n%zu %zu\n-1.0\n", width, height);
    bool ok = true;
    for (size_t y = height; y > 0 && ok; --y)
        ok = std::fwrite(rgb + (y - 1) * width * 3, sizeof(float), width * 3, fp) == width * 3;
    return std::fclose(fp) == 0 && ok;
}

// The bytes the reference's WRITE_PFM build produces: ostream << float is printf("%g")
inline bool writeReferencePFM(const char* path, size_t width, size_t height, const float* rgb)
{
    FILE* fp = std::fopen(path, "wb");
    if (fp == NULL)
        return false;
    std::fprintf(fp, "PF\n%zu %zu\n-1.0\n", width, height);
    bool ok = true;
    for (size_t i = 0; i < width * height * 3 && ok; ++i)
        ok = std::fprintf(fp, "%g", (double)rgb[i]) > 0;
    return std::fclose(fp) == 0 && ok;
}

} // namespace rayito_b200

#endif // RAYITO_B200_IMAGEIO_HPP
