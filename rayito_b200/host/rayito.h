// Drop-in header name of the Rayito API; the implementation lives in rayito_b200/render.hpp.
#ifndef RAYITO_B200_COMPAT_RAYITO_H
#define RAYITO_B200_COMPAT_RAYITO_H
#include "rayito_b200/render.hpp"
#endif
