// Shared by cli/rayito_stage{1,2,3}/main.cpp: option parsing and output writing of the
// command-line stages.  The programs keep the reference's contract -- no arguments, 512 x 512,
// out.ppm in the working directory (Rayito_Stage1/main.cpp:60-62, 86) -- and add a few options:
//   --width N --height N    image size (reference: compile-time 512 x 512)
//   --samples N [M]         Stage 2: N samples per pixel (64); Stage 3: N x M strata (4 x 4)
//   --pfm                   also write out.pfm, a standard binary PFM of the unclamped image
//   --reference-pfm         write out.pfm with the bytes of the reference's WRITE_PFM build instead
//   --device N              CUDA device ordinal
//   -o PATH                 output path of the PPM
#ifndef RAYITO_B200_STAGE_CLI_HPP
#define RAYITO_B200_STAGE_CLI_HPP

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rayito_b200_host.h"
#include "rayito_b200/imageio.hpp"

struct StageOptions
{
    unsigned width, height, samplesU, samplesV;
    int device;
    bool pfm, referencePfm;
    std::string out;
    StageOptions() : width(512), height(512), samplesU(0), samplesV(0), device(0), pfm(false), referencePfm(false), out("out.ppm") { }
};

inline bool parseStageOptions(int argc, char** argv, StageOptions& o)
{
    for (int i = 1; i < argc; ++i)
    {
        std::string a = argv[i];
        if (a == "--width" && i + 1 < argc) o.width = (unsigned)std::atoi(argv[++i]);
        else if (a == "--height" && i + 1 < argc) o.height = (unsigned)std::atoi(argv[++i]);
        else if (a == "--samples" && i + 1 < argc)
        {
            o.samplesU = (unsigned)std::atoi(argv[++i]);
            if (i + 1 < argc && argv[i + 1][0] != '-') o.samplesV = (unsigned)std::atoi(argv[++i]);
        }
        else if (a == "--device" && i + 1 < argc) o.device = std::atoi(argv[++i]);
        else if (a == "--pfm") o.pfm = true;
        else if (a == "--reference-pfm") o.referencePfm = true;
        else if (a == "-o" && i + 1 < argc) o.out = argv[++i];
        else
        {
            std::fprintf(stderr, "usage: %s [--width N] [--height N] [--samples N [M]] [--pfm | --reference-pfm] [--device N] [-o out.ppm]\n", argv[0]);
            return false;
        }
    }
    return o.width >= 2 && o.height >= 2;
}

inline int writeStageOutputs(const StageOptions& o, const std::vector<float>& rgb, const std::vector<unsigned char>& rgb8)
{
    if (!rayito_b200::writePPM(o.out.c_str(), o.width, o.height, &rgb8[0]))
    {
        std::fprintf(stderr, "cannot write %s\n", o.out.c_str());
        return 1;
    }
    if (o.pfm || o.referencePfm)
    {
        std::string path = o.out.size() > 4 && o.out.compare(o.out.size() - 4, 4, ".ppm") == 0
                               ? o.out.substr(0, o.out.size() - 4) + ".pfm" : o.out + ".pfm";
        bool ok = o.referencePfm ? rayito_b200::writeReferencePFM(path.c_str(), o.width, o.height, &rgb[0])
                                 : rayito_b200::writePFM(path.c_str(), o.width, o.height, &rgb[0]);
        if (!ok)
        {
            std::fprintf(stderr, "cannot write %s\n", path.c_str());
            return 1;
        }
    }
    return 0;
}

#endif // RAYITO_B200_STAGE_CLI_HPP
