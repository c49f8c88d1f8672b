// Drop-in header name of the Rayito API; the implementation lives in rayito_b200/math.hpp.
#ifndef RAYITO_B200_COMPAT_RMATH_H
#define RAYITO_B200_COMPAT_RMATH_H
#include "rayito_b200/math.hpp"
#endif
