// Drop-in header name of the Rayito API; the implementation lives in rayito_b200/render.hpp.
//
// An application written against the Stage 6 API (no transforms, no shutter) compiles
// with -DRAYITO_B200_STAGE=6 to get Stage 6 rendering rules (RtSceneDesc.semantics).
#ifndef RAYITO_B200_COMPAT_RAYITO_H
#define RAYITO_B200_COMPAT_RAYITO_H
#include "rayito_b200/render.hpp"
#if defined(RAYITO_B200_STAGE) && RAYITO_B200_STAGE == 6
namespace
{
struct RayitoB200SelectStage6
{
    RayitoB200SelectStage6() { rayito_b200::stageSemantics() = RT_SEMANTICS_STAGE6; }
} g_rayitoB200SelectStage6;
}
#endif
#endif
