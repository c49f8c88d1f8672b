// placeholder until the wavefront renderer lands
#ifndef RAYITO_B200_RT_RENDER_CUH
#define RAYITO_B200_RT_RENDER_CUH
#include "rt_scene.cuh"
inline void rt_render_release(RtScene*) { }
inline int rt_render_impl(RtScene*, const RtCamera*, const RtRenderParams*, float*, bool, RtRenderStats*, cudaStream_t) { return rt_fail(RT_ERR_UNSUPPORTED, "render not built"); }
inline int rt_camera_rays_impl(RtScene*, const RtCamera*, const RtRenderParams*, uint32_t, RtRay*) { return rt_fail(RT_ERR_UNSUPPORTED, "render not built"); }
inline int rt_tonemap_impl(int, const float*, size_t, float, float, uint8_t*) { return rt_fail(RT_ERR_UNSUPPORTED, "render not built"); }
#endif
