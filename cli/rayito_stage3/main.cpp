// rayito (Stage 3) on the B200: the program of Rayito_Stage3/main.cpp rendered by the CUDA core
// (rt_stage23_render: the program's ONE serial random stream is located on the device, every
// sample then runs in parallel at its exact position in it).  `make && ./rayito` writes the very
// out.ppm the reference program writes.
#include "stage_cli.hpp"

int main(int argc, char** argv)
{
    StageOptions opt;
    if (!parseStageOptions(argc, argv, opt))
        return 2;
    // kNumPixelSamples (Stage 2: 64) / kNumPixelSamplesU x V (Stage 3: 4 x 4)
    const unsigned su = opt.samplesU ? opt.samplesU : (3 == 2 ? 64u : 4u);
    const unsigned sv = opt.samplesV ? opt.samplesV : (3 == 2 ? 1u : (opt.samplesU ? opt.samplesU : 4u));
    std::vector<float> rgb((size_t)opt.width * opt.height * 3);
    std::vector<unsigned char> rgb8((size_t)opt.width * opt.height * 3);
    RtRenderStats stats;
    if (rth_stage23_render(opt.device, 3, opt.width, opt.height, su, sv, &rgb[0], &rgb8[0], &stats) != 0)
    {
        std::fprintf(stderr, "rayito: %s\n", rth_last_error_string());
        return 1;
    }
    return writeStageOutputs(opt, rgb, rgb8);
}
