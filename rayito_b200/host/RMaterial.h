// Drop-in header name of the Rayito API; the implementation lives in rayito_b200/scene.hpp.
#ifndef RAYITO_B200_COMPAT_RMATERIAL_H
#define RAYITO_B200_COMPAT_RMATERIAL_H
#include "rayito_b200/scene.hpp"
#endif
