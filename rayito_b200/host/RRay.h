// Drop-in header name of the Rayito API; the implementation lives in rayito_b200/accel.hpp.
#ifndef RAYITO_B200_COMPAT_RRAY_H
#define RAYITO_B200_COMPAT_RRAY_H
#include "rayito_b200/accel.hpp"
#endif
