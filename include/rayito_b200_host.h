/*
 * rayito_b200_host -- C entry points of the host library (librayito_host.so).
 *
 * The host library is the C++ mirror of the Rayito API (namespace Rayito: Shape, ShapeSet,
 * Plane, Sphere, Mesh, lights, materials, PerspectiveCamera, raytrace(); plus
 * rayito_b200::raytraceToDevice / raytraceMulti for several GPUs and the image writers of
 * rayito_b200/imageio.hpp): applications use it from C++ by including its rayito.h.  The few
 * C functions here are what has no C++ class behind it in the reference: the camera
 * description and the self-contained Stage 1 / 2 / 3 programs (their scenes are part of the
 * programs, Rayito_Stage1/main.cpp:65-75, Rayito_Stage3/main.cpp:165-201).
 * Scenes for tests and benchmarks (the GUI's scenes, the synthetic mesh, edge cases) are NOT
 * here: see fixtures/rayito_fixtures.h.
 */
#ifndef RAYITO_B200_HOST_H
#define RAYITO_B200_HOST_H

#include "rayito_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* rth_last_error_string(void);

/* Which face BVH Mesh::prepare() builds from now on, on the calling thread (C++: rayito_b200::treeMode()).
 * 3 = DEFAULT: per mesh, mode 2 from 65 536 faces up and mode 0 below (the same tree either way);
 * 0 = the reference's tree, node for node (Bvh<T>::buildRange, Rayito_Stage7_QT/RAccel.h:290-374; default:
 * hit records bit-equal to the reference); 1 = PERF MODE, a binned-SAH tree in the same node format, traversed
 * by the same kernels.  The reference's slab test is not watertight, so another tree may decide a grazing
 * ray differently: with mode 1 parity is measured (tests/test_gpu_perf_tree.py), not bit-exact. */
int rth_set_tree_mode(unsigned mode);

/* PerspectiveCamera constructor (RaytraceMain.cpp:205-222).  spec14 = fov degrees,
 * origin xyz, target xyz, up xyz, focal distance, lens radius, shutter open, close. */
int rth_camera(const float* spec14, RtCamera* out);
/* Stage 1 program (Rayito_Stage1/main.cpp:65-135): builds its scene (one pink plane at
 * y = -2) and camera (fov 30 at the origin looking down +z) with makeCameraRay's basis
 * arithmetic (main.cpp:28-52), renders on the GPU and returns the P6 payload
 * (width*height*3 bytes); what `make && ./rayito` writes after the "P6" header. */
int rth_stage1_render(int device, unsigned width, unsigned height, unsigned char* rgb8);
/* ... also handing back pixelColor before clamp() (rgb: width*height*3 floats; what WRITE_PFM streams out) */
int rth_stage1_render_float(int device, unsigned width, unsigned height, float* rgb, unsigned char* rgb8);

/* Stage 2 and Stage 3 programs (Rayito_Stage2/main.cpp:93-228, Rayito_Stage3/main.cpp:162-279):
 * build the program's scene and camera (fov 45 at (0,5,15) looking at the origin) with
 * the reference's own constructor arithmetic and render on the GPU.  stage = 2: samples_u
 * random samples per pixel (reference 64), samples_v ignored; stage = 3: samples_u x
 * samples_v stratified samples per pixel (reference 4 x 4; config C2 sweeps 1..16 squared),
 * 4 x 4 samples per light.  rgb / rgb8 / stats as for rt_stage23_render. */
int rth_stage23_render(int device, int stage, unsigned width, unsigned height, unsigned samples_u, unsigned samples_v,
                       float* rgb, unsigned char* rgb8, RtRenderStats* stats);

#ifdef __cplusplus
}
#endif

#endif /* RAYITO_B200_HOST_H */
