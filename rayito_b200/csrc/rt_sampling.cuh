// Counter-based reproduction of the reference's sample stream, and the sample
// warps (Rayito_Stage7_QT/RSampling.h).
//
// Stage 7 draws ALL its random numbers from CorrelatedMultiJitterSampler, a
// stateless integer hash of (index, permutation).  The only sequential state is
// one multiply-with-carry Rng per image chunk that hands out a fixed 5*depth+3
// permutations per pixel (RaytraceMain.cpp:69-108,159-169).  MWC with lag 1 is a
// linear congruential step modulo a*2^16-1, so the state for any pixel is reached
// by modular exponentiation instead of stepping through every earlier pixel:
// that is what lets the GPU render any tile of the image with the exact stream.
#ifndef RAYITO_B200_RT_SAMPLING_CUH
#define RAYITO_B200_RT_SAMPLING_CUH

#include "rt_device.cuh"
#include "rt_libm.cuh"

#define RT_PI_D 3.14159265358979323846   /* M_PI: a double, as on the reference's libc */

// ---------------------------------------------------------------------------
// Rng (RSampling.h:27-58): z = 36969*(z & 65535) + (z >> 16); w likewise with
// 18000; output (z << 16) + w.
// ---------------------------------------------------------------------------
#define RT_MWC_AZ 36969u
#define RT_MWC_AW 18000u
#define RT_MWC_MZ 2422800383ull   /* 36969 * 2^16 - 1 */
#define RT_MWC_MW 1179647999ull   /* 18000 * 2^16 - 1 */

struct MwcState
{
    uint32_t z, w;
};

__host__ __device__ __forceinline__ uint32_t mwc_next(MwcState& s)
{
    s.z = RT_MWC_AZ * (s.z & 65535u) + (s.z >> 16);
    s.w = RT_MWC_AW * (s.w & 65535u) + (s.w >> 16);
    return (s.z << 16) + s.w;
}

__host__ __device__ __forceinline__ uint64_t mwc_powmod(uint64_t base, uint64_t exp, uint64_t mod)
{
    uint64_t result = 1, b = base % mod;
    while (exp)
    {
        if (exp & 1) result = (result * b) % mod;
        b = (b * b) % mod;
        exp >>= 1;
    }
    return result;
}

// State after `steps` calls of nextUInt32, given the state after exactly two
// calls (two literal steps make any 32-bit seed canonical).
// steps >= 2.  A state congruent to 0 is a fixed point (0 or the modulus itself).
__host__ __device__ __forceinline__ MwcState mwc_jump(MwcState after2, uint64_t steps)
{
    MwcState s;
    uint64_t e = steps - 2;
    if (e == 0)
        return after2;      // z2 may still be the non-canonical representative M+1
    uint64_t z = (mwc_powmod(RT_MWC_AZ, e, RT_MWC_MZ) * (after2.z % RT_MWC_MZ)) % RT_MWC_MZ;
    uint64_t w = (mwc_powmod(RT_MWC_AW, e, RT_MWC_MW) * (after2.w % RT_MWC_MW)) % RT_MWC_MW;
    s.z = (uint32_t)((z == 0 && after2.z != 0) ? RT_MWC_MZ : z);
    s.w = (uint32_t)((w == 0 && after2.w != 0) ? RT_MWC_MW : w);
    return s;
}

// Chunk geometry of the reference renderer (RaytraceMain.cpp:504-547): up to 4x4
// chunks, one Rng each, seeded from the chunk bounds (:69-70).
struct ChunkGrid
{
    uint32_t width, height;
    uint32_t cw, ch;        // chunk size
    uint32_t nx, ny;        // chunk counts
};

__host__ __device__ __forceinline__ ChunkGrid chunk_grid(uint32_t width, uint32_t height)
{
    ChunkGrid g;
    g.width = width;
    g.height = height;
    g.cw = width >= 4 ? width / 4 : 1;
    g.ch = height >= 4 ? height / 4 : 1;
    g.nx = width > 4 ? width / g.cw : 1;
    g.ny = height > 4 ? height / g.ch : 1;
    if (g.nx * g.cw < width) g.nx++;
    if (g.ny * g.ch < height) g.ny++;
    return g;
}

// QUIRK: for images narrower (or lower) than 4 pixels the reference makes at most two
// one-pixel chunks per axis (RaytraceMain.cpp:508-516), so pixels beyond them are
// never rendered and keep Image's default black.
__host__ __device__ __forceinline__ bool chunk_covers(const ChunkGrid& g, uint32_t x, uint32_t y)
{
    return x < g.nx * g.cw && y < g.ny * g.ch;
}

// The 5*depth+3 permutations pixel (x, y) renders with.  out[] order:
//   per bounce b: out[5b+0] bounce, +1 light selection, +2 light element, +3 light, +4 brdf
//   then out[5D+0] time, out[5D+1] lens, out[5D+2] subpixel
// (RaytraceMain.cpp:82-108 for the first pixel of a chunk, :159-169 for the rest:
// note the tail order differs -- time,lens,subpixel vs lens,time,subpixel).
__host__ __device__ inline void pixel_permutations(const ChunkGrid& g, uint32_t x, uint32_t y, uint32_t depth, uint32_t* out)
{
    uint32_t cx = x / g.cw, cy = y / g.ch;
    if (cx >= g.nx) cx = g.nx - 1;     // cannot happen for reference chunking; keeps indices sane
    if (cy >= g.ny) cy = g.ny - 1;
    uint64_t xs = (uint64_t)cx * g.cw, ys = (uint64_t)cy * g.ch;
    uint64_t xe = xs + g.cw < g.width ? xs + g.cw : g.width;
    uint64_t ye = ys + g.ch < g.height ? ys + g.ch : g.height;
    MwcState s;
    s.z = (uint32_t)(((xs << 16) | xe) ^ xs);
    s.w = (uint32_t)(((ys << 16) | ye) ^ ys);
    uint64_t k = (uint64_t)(y - ys) * (xe - xs) + (x - xs);     // pixel index inside the chunk
    uint32_t per_pixel = 5 * depth + 3;
    uint64_t skip = k * per_pixel;                              // draws consumed by earlier pixels
    if (skip >= 2)
    {
        MwcState a2 = s;
        mwc_next(a2);
        mwc_next(a2);
        s = mwc_jump(a2, skip);
    }
    else
    {
        for (uint64_t i = 0; i < skip; ++i) mwc_next(s);
    }
    for (uint32_t i = 0; i < 5 * depth; ++i)
        out[i] = mwc_next(s);
    uint32_t t0 = mwc_next(s), t1 = mwc_next(s), t2 = mwc_next(s);
    if (k == 0) { out[5 * depth + 0] = t0; out[5 * depth + 1] = t1; }
    else        { out[5 * depth + 0] = t1; out[5 * depth + 1] = t0; }
    out[5 * depth + 2] = t2;
}

// ---------------------------------------------------------------------------
// CorrelatedMultiJitterSampler (RSampling.h:253-375), Kensler's CMJ
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t cmj_permute(uint32_t i, uint32_t num, uint32_t p)
{
    uint32_t w = num - 1;
    w |= w >> 1;
    w |= w >> 2;
    w |= w >> 4;
    w |= w >> 8;
    w |= w >> 16;
    do
    {
        i ^= p;
        i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;
        i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1;
        i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11;
        i *= 0x74dcb303u;
        i ^= (i & w) >> 2;
        i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;
        i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= num);
    return (i + p) % num;
}

__host__ __device__ __forceinline__ float cmj_rand01(uint32_t i, uint32_t p)
{
    i ^= p;
    i ^= i >> 17;
    i ^= i >> 10;
    i *= 0xb36534e5u;
    i ^= i >> 12;
    i ^= i >> 21;
    i *= 0x93fc4795u;
    i ^= 0xdf6e307fu;
    i ^= i >> 17;
    i *= 1u | p >> 18;
    return (float)i * 2.328306e-10f;
}

// sample1D (RSampling.h:272-279); index < samples is guaranteed by the caller
__host__ __device__ __forceinline__ float cmj_sample1d(uint32_t index, uint32_t samples, uint32_t perm)
{
    uint32_t s = cmj_permute(index, samples, perm * 0x8ff3cd11u);
    float sx = cmj_rand01(s, perm * 0xa399d265u);
    return ((float)s + sx) / (float)samples;
}

// sample2D (RSampling.h:288-306); index < xs*ys is guaranteed by the caller
__host__ __device__ __forceinline__ void cmj_sample2d(uint32_t index, uint32_t xs, uint32_t ys, uint32_t perm, float& u, float& v)
{
    uint32_t s = cmj_permute(index, xs * ys, perm * 0xc2d3c8fbu);
    int ix = (int)cmj_permute(s % xs, xs, perm * 0xa511e9b3u);
    int iy = (int)cmj_permute(s / xs, ys, perm * 0x63d83595u);
    float sx = cmj_rand01(s, perm * 0xa399d265u);
    float sy = cmj_rand01(s, perm * 0x711ad6a5u);
    u = ((float)ix + ((float)iy + sx) / (float)ys) / (float)xs;
    v = ((float)s + sy) / (float)(xs * ys);
}

#ifdef __CUDACC__
// Entry points used by the renderer's kernels; -DRT_CMJ_CALL=__noinline__ makes them real calls
// (measured slower together with RT_SHADE_CALL, see rt_shade.cuh)
#ifndef RT_CMJ_CALL
#define RT_CMJ_CALL __forceinline__
#endif
// Power-of-two sample counts (the GUI's default 16 x 16 pixel samples, any power-of-two hint) take an
// exact short cut: with num = 2^k the mask w is num - 1, so `i &= w` leaves i < num and the cycle walk
// ends after its first round, and (i + p) % num, s % xs, s / xs are a mask and a shift.  Same integers,
// without the five 32-bit divisions of a 2-D sample (~25 instructions each) and without the loop.  Other
// counts go through the general code, kept out of line so that the kernels carry it once.
__device__ __forceinline__ uint32_t cmj_permute_pow2(uint32_t i, uint32_t w, uint32_t p)
{
    i ^= p;
    i *= 0xe170893du;
    i ^= p >> 16;
    i ^= (i & w) >> 4;
    i ^= p >> 8;
    i *= 0x0929eb3fu;
    i ^= p >> 23;
    i ^= (i & w) >> 1;
    i *= 1u | p >> 27;
    i *= 0x6935fa69u;
    i ^= (i & w) >> 11;
    i *= 0x74dcb303u;
    i ^= (i & w) >> 2;
    i *= 0x9e501cc3u;
    i ^= (i & w) >> 2;
    i *= 0xc860a3dfu;
    i &= w;
    i ^= i >> 5;
    return (i + p) & w;
}
__device__ __noinline__ float cmj1d_general(uint32_t index, uint32_t samples, uint32_t perm) { return cmj_sample1d(index, samples, perm); }
__device__ __noinline__ void cmj2d_general(uint32_t index, uint32_t xs, uint32_t ys, uint32_t perm, float& u, float& v)
{
    cmj_sample2d(index, xs, ys, perm, u, v);
}
#ifndef RT_CMJ_POW2
#define RT_CMJ_POW2 1       /* 0: always the general code (A/B runs) */
#endif
__device__ RT_CMJ_CALL float cmj1d(uint32_t index, uint32_t samples, uint32_t perm)
{
    if (!RT_CMJ_POW2 || (samples & (samples - 1u)) != 0u)
        return cmj1d_general(index, samples, perm);
    uint32_t s = cmj_permute_pow2(index, samples - 1u, perm * 0x8ff3cd11u);
    float sx = cmj_rand01(s, perm * 0xa399d265u);
    return ((float)s + sx) / (float)samples;
}
__device__ RT_CMJ_CALL void cmj2d(uint32_t index, uint32_t xs, uint32_t ys, uint32_t perm, float& u, float& v)
{
    if (!RT_CMJ_POW2 || ((xs & (xs - 1u)) | (ys & (ys - 1u))) != 0u)
    {
        cmj2d_general(index, xs, ys, perm, u, v);
        return;
    }
    const uint32_t shift = 31u - (uint32_t)__clz((int)xs);      // log2(xs)
    uint32_t s = cmj_permute_pow2(index, xs * ys - 1u, perm * 0xc2d3c8fbu);
    int ix = (int)cmj_permute_pow2(s & (xs - 1u), xs - 1u, perm * 0xa511e9b3u);
    int iy = (int)cmj_permute_pow2(s >> shift, ys - 1u, perm * 0x63d83595u);
    float sx = cmj_rand01(s, perm * 0xa399d265u);
    float sy = cmj_rand01(s, perm * 0x711ad6a5u);
    u = ((float)ix + ((float)iy + sx) / (float)ys) / (float)xs;
    v = ((float)s + sy) / (float)(xs * ys);
}

// ---------------------------------------------------------------------------
// libm.  The reference calls the C library's float cos/sin/pow; they feed sample
// DIRECTIONS, so they are reproduced bit for bit (rt_libm.cuh restates glibc 2.39's
// algorithms; tests/test_libm_cpu.py pins the host build of that file to the C
// library).  Kept out of line: they are called from a dozen places.
// ---------------------------------------------------------------------------
__device__ __noinline__ void ref_sincosf(float x, float& s, float& c)
{
    rtm_sincosf(x, s, c);
}
__device__ __forceinline__ float ref_cosf(float x) { return rtm_cosf(x); }
__device__ __forceinline__ float ref_sinf(float x) { return rtm_sinf(x); }
__device__ __noinline__ float ref_powf(float x, float y) { return rtm_powf(x, y); }

// Vector(0,1,0)-or-(1,0,0) frame around a direction (RMath.h:946-955)
__device__ __forceinline__ void make_frame(V3 ref, V3& x, V3& y, V3& z)
{
    z = normalized3(ref);
    V3 v2 = (z.x != 0.0f || z.z != 0.0f) ? mk(0.0f, 1.0f, 0.0f) : mk(1.0f, 0.0f, 0.0f);
    x = normalized3(cross3(v2, z));
    y = cross3(z, x);
}

// transformFromLocalCoordinateSpace (RMath.h:978-986)
__device__ __forceinline__ V3 frame_to_world(V3 v, V3 x, V3 y, V3 z)
{
    return mk(v.x * x.x + v.y * y.x + v.z * z.x,
              v.x * x.y + v.y * y.y + v.z * z.y,
              v.x * x.z + v.y * y.z + v.z * z.z);
}

// concentricSampleDisk (RSampling.h:400-452)
__device__ __forceinline__ void concentric_disk(float u1, float u2, float& dx, float& dy)
{
    float r, theta;
    float sx = 2.0f * u1 - 1.0f;
    float sy = 2.0f * u2 - 1.0f;
    if (sx == 0.0f && sy == 0.0f)
    {
        dx = 0.0f;
        dy = 0.0f;
        return;
    }
    if (sx >= -sy)
    {
        if (sx > sy)
        {
            r = sx;
            if (sy > 0.0f) theta = sy / r;
            else theta = 8.0f + sy / r;
        }
        else
        {
            r = sy;
            theta = 2.0f - sx / r;
        }
    }
    else
    {
        if (sx <= sy)
        {
            r = -sx;
            theta = 4.0f - sy / r;
        }
        else
        {
            r = -sy;
            theta = 6.0f + sx / r;
        }
    }
    theta = (float)((double)theta * (RT_PI_D / 4.0));      // theta *= M_PI / 4.0f, in double
    float sn, cs;
    ref_sincosf(theta, sn, cs);
    dx = r * cs;
    dy = r * sn;
}

// uniformToCosineHemisphere (RSampling.h:500-508)
__device__ __forceinline__ V3 cosine_hemisphere(float u1, float u2)
{
    float dx, dy;
    concentric_disk(u1, u2, dx, dy);
    float z = sqrtf(std_max(0.0f, 1.0f - dx * dx - dy * dy));
    return mk(dx, dy, z);
}

// uniformToSphere (RSampling.h:456-466)
__device__ __forceinline__ V3 uniform_sphere(float u1, float u2)
{
    float z = 1.0f - 2.0f * u1;
    float radius = sqrtf(std_max(0.0f, 1.0f - z * z));
    float phi = (float)((RT_PI_D * 2.0) * (double)u2);     // M_PI * 2.0f * u2
    float sn, cs;
    ref_sincosf(phi, sn, cs);
    return mk(radius * cs, radius * sn, z);
}

// uniformToCone / uniformConePdf (RSampling.h:512-523)
__device__ __forceinline__ V3 uniform_cone(float u1, float u2, float cos_theta_max)
{
    float cos_theta = u1 * (cos_theta_max - 1.0f) + 1.0f;
    float sin_theta = sqrtf(std_max(0.0f, 1.0f - cos_theta * cos_theta));
    float phi = (float)(((double)u2 * RT_PI_D) * 2.0);     // u2 * M_PI * 2.0f
    float sn, cs;
    ref_sincosf(phi, sn, cs);
    return mk(cs * sin_theta, sn * sin_theta, cos_theta);
}

__device__ __forceinline__ float uniform_cone_pdf(float cos_theta_max)
{
    // cosThetaMax >= 1.0f ? 0.0 : 1.0f / (2.0f * M_PI * (1.0f - cosThetaMax))  -- a double expression
    if (cos_theta_max >= 1.0f)
        return 0.0f;
    return (float)(1.0 / ((2.0 * RT_PI_D) * (double)(1.0f - cos_theta_max)));
}

// uniformToUniformDisk (RSampling.h:470-485), used by depth of field
__device__ __forceinline__ void uniform_disk(float u1, float u2, float& dx, float& dy)
{
    float radius = sqrtf(u1);
    float theta = (float)((RT_PI_D * 2.0) * (double)u2);
    float sn, cs;
    ref_sincosf(theta, sn, cs);
    dx = radius * cs;
    dy = radius * sn;
}

// uniformToBarycentricTriangle (RSampling.h:527-532)
__device__ __forceinline__ void uniform_barycentric(float u1, float u2, float& a, float& b)
{
    float s = sqrtf(u1);
    a = 1.0f - s;
    b = u2 * s;
}

// powerHeuristic(1, pdf1, 1, pdf2) (RSampling.h:387-392)
__device__ __forceinline__ float power_heuristic(float pdf1, float pdf2)
{
    float w1 = 1.0f * pdf1;
    float w2 = 1.0f * pdf2;
    return w1 * w1 / (w1 * w1 + w2 * w2);
}
#endif // __CUDACC__

#endif // RAYITO_B200_RT_SAMPLING_CUH
