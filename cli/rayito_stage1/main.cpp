// rayito (Stage 1) on the B200: the program of Rayito_Stage1/main.cpp -- one pink plane, one ray
// through every pixel corner, out.ppm -- rendered by the CUDA core (rt_stage1_render) instead of
// the scalar loop.  `make && ./rayito` writes the reference's out_ref.ppm byte for byte.
#include "stage_cli.hpp"

int main(int argc, char** argv)
{
    StageOptions opt;
    if (!parseStageOptions(argc, argv, opt))
        return 2;
    std::vector<float> rgb((size_t)opt.width * opt.height * 3);
    std::vector<unsigned char> rgb8((size_t)opt.width * opt.height * 3);
    if (rth_stage1_render_float(opt.device, opt.width, opt.height, &rgb[0], &rgb8[0]) != 0)
    {
        std::fprintf(stderr, "rayito: %s\n", rth_last_error_string());
        return 1;
    }
    return writeStageOutputs(opt, rgb, rgb8);
}
