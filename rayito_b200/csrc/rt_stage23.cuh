// Stage 2 and Stage 3 of the tutorial (BASELINE config C2: the Stage 3 spp sweep):
// the small serial-Rng programs Rayito_Stage2/main.cpp and Rayito_Stage3/main.cpp.
//
// Both draw every random number from ONE multiply-with-carry Rng in scan order, and a
// pixel sample consumes 2 draws when its camera ray misses the scene and 2 + D when it
// hits (D = 2 per light sample: Rayito_Stage2/main.cpp:156-157,183-184,
// Rayito_Stage3/main.cpp:117-127,241-242).  The stream position of sample k is therefore
//     offset(k) = 2k + D * (number of earlier samples whose camera ray hit),
// a prefix sum over a predicate that itself depends on the offset (the jitter decides
// hit or miss on silhouette pixels).  The reference resolves this by being sequential.
// Here:
//   1. k_s23_guess    -- every sample's hit predicate with centred jitter (parallel);
//   2. k_s23_prepass  -- per image segment, a windowed fix-point: evaluate a window of
//      samples with offsets from the assumed flags, accept everything up to the first
//      disagreement, restart behind it.  Samples whose hit does not depend on the jitter
//      (all but silhouettes) are accepted a whole window at a time;
//   3. k_s23_segfix   -- chains the segments' hit counts; segments whose incoming count
//      changed are redone (their silhouette samples may flip), until nothing changes;
//   4. k_s23_shade    -- one thread per pixel sample: MWC jump-ahead to offset(k), then the
//      reference's arithmetic for the camera ray, the closest hit in list order, every
//      light sample and shadow ray, Lambert / Phong shading;
//   5. k_s23_resolve  -- per pixel: the samples' terms added in the reference's order,
//      box-filter division, clamp, 8-bit truncation.
// Everything is the reference's float arithmetic (no FMA, IEEE div/sqrt, the C library's
// sinf/cosf/powf from rt_libm.cuh), so the image is the reference's bit for bit.
#ifndef RAYITO_B200_RT_STAGE23_CUH
#define RAYITO_B200_RT_STAGE23_CUH

#include "rt_sampling.cuh"
#include "rt_shade.cuh"

namespace rt_detail
{

#define RT_S23_TMIN 0.00001f          /* kRayTMin of Stages 1-3 (Rayito_Stage3/rayito.h:303) */
#define RT_S23_PRE_THREADS 256
#define RT_S23_PRE_PER_THREAD 4
#define RT_S23_WINDOW (RT_S23_PRE_THREADS * RT_S23_PRE_PER_THREAD)
#define RT_S23_MAX_TERMS 3            /* Stage 2: emitted + one term per light (<= 2 lights) */

struct S23Ctx
{
    RtS23Shape shapes[RT_S23_MAX_SHAPES];
    RtS23Material materials[RT_S23_MAX_SHAPES];
    uint32_t lights[RT_S23_MAX_LIGHTS];         // shape indices, list order
    uint32_t num_shapes, num_lights;
    RtCamera cam;
    uint32_t stage;                              // 2 or 3
    uint32_t width, height;
    uint32_t nu, nv, spp;                        // pixel samples (Stage 2: nu = spp, nv = 1)
    uint32_t lu, lv;                             // light samples per light
    uint32_t draws_per_hit;                      // D
    uint32_t terms;                              // colour terms stored per sample
    uint64_t num_samples;
    MwcState seed_after2;                        // Rng state after two draws (canonical form for the jump)
    MwcState seed;

    uint8_t* flags;                              // per sample: camera ray hit something
    uint32_t* hits_before;                       // per sample: hits among earlier samples
    uint64_t seg_len;
    uint32_t num_segs;
    unsigned long long* seg_hin;                 // hits before the segment (as last run)
    unsigned long long* seg_hout;                // ... after it
    uint32_t* seg_dirty;
    uint32_t* any_dirty;
    float4* sample_terms;                        // [terms][num_samples]
    float* rgb;                                  // width*height*3, before clamp
    uint8_t* rgb8;
};

struct S23Hit
{
    float t;
    int shape;          // list index of the winning entry
    bool is_light_self; // m_pShape == the Light object itself (RectangleLight yes, ShapeLight no: rayito.h:710-719)
    V3 normal;
    float cmod;         // colour modifier (grey)
};

// Rng state after `steps` draws
__device__ __forceinline__ MwcState s23_rng_at(const S23Ctx& c, uint64_t steps)
{
    if (steps >= 2)
        return mwc_jump(c.seed_after2, steps);
    MwcState s = c.seed;
    if (steps == 1) mwc_next(s);
    return s;
}

__device__ __forceinline__ float s23_next_float(MwcState& s)
{
    return (float)mwc_next(s) * 2.328306e-10f;       // Rng::nextFloat (main.cpp:38-42)
}

// Vector::normalize of Stages 1-3 divides unconditionally (rayito.h:194)
__device__ __forceinline__ V3 s23_normalized(V3 v, float* len_out = NULL)
{
    float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (len_out) *len_out = len;
    return mk(v.x / len, v.y / len, v.z / len);
}

// makeCameraRay (Rayito_Stage3/main.cpp:55-79) with the pixel-independent basis hoisted
__device__ __forceinline__ void s23_camera_ray(const S23Ctx& c, float xu, float yu, V3& o, V3& d)
{
    V3 fwd = mk(c.cam.forward[0], c.cam.forward[1], c.cam.forward[2]);
    V3 right = mk(c.cam.right[0], c.cam.right[1], c.cam.right[2]);
    V3 up = mk(c.cam.up[0], c.cam.up[1], c.cam.up[2]);
    o = mk(c.cam.origin[0], c.cam.origin[1], c.cam.origin[2]);
    d = fwd + right * ((xu - 0.5f) * c.cam.tan_fov) + up * ((yu - 0.5f) * c.cam.tan_fov);
    d = s23_normalized(d);
}

// Screen position of sample k given its two jitter draws (first draw -> yu, second -> xu)
__device__ __forceinline__ void s23_screen(const S23Ctx& c, uint64_t k, float r_first, float r_second, float& xu, float& yu)
{
    uint32_t s = (uint32_t)(k % c.spp);
    uint64_t p = k / c.spp;
    uint32_t x = (uint32_t)(p % c.width), y = (uint32_t)(p / c.width);
    if (c.stage == 2)
    {
        // Rayito_Stage2/main.cpp:156-157
        yu = 1.0f - (((float)y + r_first) / (float)(c.height - 1));
        xu = ((float)x + r_second) / (float)(c.width - 1);
    }
    else
    {
        // Rayito_Stage3/main.cpp:241-242
        uint32_t vsi = s / c.nu, usi = s % c.nu;
        yu = 1.0f - (((float)y + ((float)vsi + r_first) / (float)c.nv) / (float)c.height);
        xu = ((float)x + ((float)usi + r_second) / (float)c.nu) / (float)c.width;
    }
}

// ShapeSet::intersect (rayito.h:543-558): every list entry in order, each accepting
// only t < current m_t.  SHADING = also fill normal / colour modifier.
template <bool SHADING>
__device__ __forceinline__ bool s23_intersect(const S23Ctx& c, V3 o, V3 d, float tmax, S23Hit& h)
{
    h.t = tmax;
    h.shape = -1;
    h.is_light_self = false;
    h.cmod = 1.0f;
    h.normal = mk(0.0f, 0.0f, 0.0f);
    bool any = false;
    for (uint32_t i = 0; i < c.num_shapes; ++i)
    {
        const RtS23Shape& sh = c.shapes[i];
        V3 pos = mk(sh.position[0], sh.position[1], sh.position[2]);
        if (sh.type == RT_S23_PLANE)
        {
            // Plane::intersect (rayito.h:745-780): one-sided
            V3 n = mk(sh.normal[0], sh.normal[1], sh.normal[2]);
            float n_dot_d = dot3(n, d);
            if (n_dot_d >= 0.0f)
                continue;
            float t = (dot3(pos, n) - dot3(o, n)) / dot3(d, n);
            if (t >= h.t || t < RT_S23_TMIN)
                continue;
            h.t = t;
            h.shape = (int)i;
            h.is_light_self = false;
            any = true;
            if (SHADING)
            {
                h.normal = n;
                h.cmod = 1.0f;
                if (sh.bullseye)
                {
                    V3 at = o + t * d;
                    if (fmodf(length3(at - pos) * 0.25f, 1.0f) > 0.5f)
                        h.cmod = 1.0f * 0.2f;
                }
            }
        }
        else if (sh.type == RT_S23_SPHERE)
        {
            // Sphere::intersect (rayito.h:815-901)
            V3 lo = o - pos;
            float a = d.x * d.x + d.y * d.y + d.z * d.z;
            float b = 2.0f * dot3(d, lo);
            float cc = (lo.x * lo.x + lo.y * lo.y + lo.z * lo.z) - sh.radius * sh.radius;
            float disc = b * b - 4.0f * a * cc;
            if (disc < 0.0f)
                continue;
            disc = sqrtf(disc);
            float q = b < 0.0f ? -0.5f * (b - disc) : -0.5f * (b + disc);
            float t0 = q / a;
            float t1 = q != 0.0f ? cc / q : h.t;
            if (t0 > t1) { float tmp = t1; t1 = t0; t0 = tmp; }
            if (t0 >= h.t || t1 < RT_S23_TMIN)
                continue;
            if (t0 >= RT_S23_TMIN)
                h.t = t0;
            else if (t1 < h.t)
                h.t = t1;
            else
                continue;
            h.shape = (int)i;
            h.is_light_self = false;          // a ShapeLight leaves m_pShape = the inner sphere
            any = true;
            if (SHADING)
            {
                h.normal = s23_normalized(lo + h.t * d);
                h.cmod = 1.0f;
            }
        }
        else
        {
            // RectangleLight::intersect (rayito.h:616-669): two-sided
            V3 s1 = mk(sh.side1[0], sh.side1[1], sh.side1[2]), s2 = mk(sh.side2[0], sh.side2[1], sh.side2[2]);
            V3 n = s23_normalized(cross3(s1, s2));
            float n_dot_d = dot3(n, d);
            if (n_dot_d == 0.0f)
                continue;
            float t = (dot3(pos, n) - dot3(o, n)) / dot3(d, n);
            if (t >= h.t || t < RT_S23_TMIN)
                continue;
            float len1, len2;
            V3 s1n = s23_normalized(s1, &len1), s2n = s23_normalized(s2, &len2);
            V3 rel = (o + t * d) - pos;
            float lx = dot3(rel, s1n), ly = dot3(rel, s2n);
            if (lx < 0.0f || lx > len1 || ly < 0.0f || ly > len2)
                continue;
            h.t = t;
            h.shape = (int)i;
            h.is_light_self = true;
            any = true;
            if (SHADING)
            {
                h.cmod = 1.0f;
                h.normal = dot3(n, d) > 0.0f ? n * -1.0f : n;
            }
        }
    }
    return any;
}

// Did the camera ray of sample k hit anything, were its jitter drawn at stream position `offset`
__device__ __forceinline__ bool s23_primary_hits(const S23Ctx& c, uint64_t k, uint64_t offset)
{
    MwcState s = s23_rng_at(c, offset);
    float r1 = s23_next_float(s), r2 = s23_next_float(s);
    float xu, yu;
    s23_screen(c, k, r1, r2, xu, yu);
    V3 o, d;
    s23_camera_ray(c, xu, yu, o, d);
    S23Hit h;
    return s23_intersect<false>(c, o, d, RT_RAY_TMAX, h);
}

// ---- 1. geometric guess ------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_s23_guess(const __grid_constant__ S23Ctx c)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    if (k < c.num_samples)
    {
        float xu, yu;
        s23_screen(c, k, 0.5f, 0.5f, xu, yu);
        V3 o, d;
        s23_camera_ray(c, xu, yu, o, d);
        S23Hit h;
        hit = s23_intersect<false>(c, o, d, RT_RAY_TMAX, h);
        c.flags[k] = hit ? 1 : 0;
    }
    // per-segment totals of the guess seed the segment chain (seg_hin = 0 here)
    uint32_t votes = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && votes)
    {
        uint64_t k0 = k;                         // first sample of this warp; a warp may straddle two segments
        uint64_t seg0 = k0 / c.seg_len;
        uint64_t boundary = (seg0 + 1) * c.seg_len;
        uint32_t in_first = boundary - k0 >= 32 ? 32u : (uint32_t)(boundary - k0);
        uint32_t lo = in_first == 32 ? votes : (votes & ((1u << in_first) - 1u));
        uint32_t hi = votes & ~lo;
        if (lo) atomicAdd(c.seg_hout + seg0, (unsigned long long)__popc(lo));
        if (hi) atomicAdd(c.seg_hout + seg0 + 1, (unsigned long long)__popc(hi));
    }
}

// ---- 2. windowed fix-point over one segment -------------------------------------------
struct S23Block
{
    uint32_t warp_a[RT_S23_PRE_THREADS / 32];
    unsigned long long warp_b[RT_S23_PRE_THREADS / 32];
    uint32_t total;
    unsigned long long first_bad;
    uint32_t confirmed;
};

__global__ void __launch_bounds__(RT_S23_PRE_THREADS)
k_s23_prepass(const __grid_constant__ S23Ctx c)
{
    const uint32_t seg = blockIdx.x;
    if (!c.seg_dirty[seg])
        return;
    __shared__ S23Block sm;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t begin = (uint64_t)seg * c.seg_len;
    const uint64_t end = begin + c.seg_len < c.num_samples ? begin + c.seg_len : c.num_samples;
    uint64_t H = c.seg_hin[seg];
    uint64_t P = begin;
    const unsigned long long NONE = ~0ull;
    while (P < end)
    {
        const uint64_t wend = P + RT_S23_WINDOW < end ? P + RT_S23_WINDOW : end;
        const uint64_t base = P + (uint64_t)tid * RT_S23_PRE_PER_THREAD;
        uint32_t a[RT_S23_PRE_PER_THREAD], cnt = 0;
        #pragma unroll
        for (int u = 0; u < RT_S23_PRE_PER_THREAD; ++u)
        {
            a[u] = base + u < wend ? c.flags[base + u] : 0;
            cnt += a[u];
        }
        // exclusive scan of the assumed counts over the block
        uint32_t incl = cnt;
        #pragma unroll
        for (int off = 1; off < 32; off <<= 1)
        {
            uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= (uint32_t)off) incl += v;
        }
        if (lane == 31) sm.warp_a[warp] = incl;
        __syncthreads();
        uint32_t before = incl - cnt;
        for (uint32_t w = 0; w < warp; ++w) before += sm.warp_a[w];

        // evaluate with the offsets the assumption implies
        uint32_t e[RT_S23_PRE_PER_THREAD];
        unsigned long long bad = NONE;
        uint64_t h = H + before;
        #pragma unroll
        for (int u = 0; u < RT_S23_PRE_PER_THREAD; ++u)
        {
            uint64_t k = base + u;
            e[u] = 0;
            if (k < wend)
            {
                e[u] = s23_primary_hits(c, k, 2ull * k + (uint64_t)c.draws_per_hit * h) ? 1u : 0u;
                if (e[u] != a[u] && bad == NONE) bad = k;
                h += a[u];
            }
        }
        // first disagreement in the window
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1)
        {
            unsigned long long v = __shfl_xor_sync(0xffffffffu, bad, off);
            bad = v < bad ? v : bad;
        }
        if (lane == 0) sm.warp_b[warp] = bad;
        __syncthreads();
        unsigned long long m = NONE;
        for (uint32_t w = 0; w < RT_S23_PRE_THREADS / 32; ++w) m = sm.warp_b[w] < m ? sm.warp_b[w] : m;

        // samples <= m were evaluated at their true offsets: final.  The rest keep the
        // evaluated flags as the next assumption.
        uint32_t conf = 0;
        h = H + before;
        #pragma unroll
        for (int u = 0; u < RT_S23_PRE_PER_THREAD; ++u)
        {
            uint64_t k = base + u;
            if (k < wend)
            {
                c.flags[k] = (uint8_t)e[u];
                if (k <= m)
                {
                    c.hits_before[k] = (uint32_t)h;
                    conf += e[u];
                }
                h += a[u];
            }
        }
        #pragma unroll
        for (int off = 16; off > 0; off >>= 1) conf += __shfl_xor_sync(0xffffffffu, conf, off);
        __syncthreads();                     // warp_a reads above are done
        if (lane == 0) sm.warp_a[warp] = conf;
        __syncthreads();
        uint32_t total_conf = 0;
        for (uint32_t w = 0; w < RT_S23_PRE_THREADS / 32; ++w) total_conf += sm.warp_a[w];
        H += total_conf;
        P = m == NONE ? wend : (uint64_t)m + 1;
        __syncthreads();
    }
    if (tid == 0)
        c.seg_hout[seg] = H;
}

// ---- 3. chain the segments -------------------------------------------------------------
__global__ void k_s23_segfix(const __grid_constant__ S23Ctx c)
{
    unsigned long long h = 0;
    uint32_t dirty = 0;
    for (uint32_t s = 0; s < c.num_segs; ++s)
    {
        unsigned long long count = c.seg_hout[s] - c.seg_hin[s];
        uint32_t d = c.seg_hin[s] != h ? 1u : 0u;
        c.seg_dirty[s] = d;
        dirty |= d;
        c.seg_hin[s] = h;
        c.seg_hout[s] = h + count;
        h += count;
    }
    *c.any_dirty = dirty;
}

// ---- 4. one pixel sample --------------------------------------------------------------
__device__ __forceinline__ Color3 s23_shade(const RtS23Material& m, V3 normal, V3 ray_dir, V3 to_light)
{
    Color3 col = mkc(m.color[0], m.color[1], m.color[2]);
    if (m.kind == RT_S23_MAT_LAMBERT)
    {
        // Lambert::shade (rayito.h:448-456)
        float f = std_max(0.0f, dot3(to_light, normal));
        return mkc(f * col.r, f * col.g, f * col.b);
    }
    if (m.kind == RT_S23_MAT_PHONG)
    {
        // Phong::shade (rayito.h:469-476)
        V3 half = s23_normalized(to_light - ray_dir);
        float f = ref_powf(std_max(0.0f, dot3(half, normal)), m.exponent);
        return mkc(f * col.r, f * col.g, f * col.b);
    }
    return mkc(0.0f, 0.0f, 0.0f);       // Emitter::shade
}

// Light::sampleSurface of entry `li` (RectangleLight rayito.h:672-681, Sphere :904-915)
__device__ __forceinline__ V3 s23_sample_light(const RtS23Shape& sh, float u1, float u2, V3 ref)
{
    V3 pos = mk(sh.position[0], sh.position[1], sh.position[2]);
    if (sh.type == RT_S23_RECT)
    {
        V3 s1 = mk(sh.side1[0], sh.side1[1], sh.side1[2]), s2 = mk(sh.side2[0], sh.side2[1], sh.side2[2]);
        return pos + s1 * u1 + s2 * u2;
    }
    // uniformToSphere (rayito.h:925-933): phi is computed in double (M_PI) and rounded
    float z = 1.0f - 2.0f * u1;
    float radius = sqrtf(std_max(0.0f, 1.0f - z * z));
    float phi = (float)(RT_PI_D * 2.0 * (double)u2);
    float sn, cs;
    ref_sincosf(phi, sn, cs);
    V3 n = mk(radius * cs, radius * sn, z);
    V3 p = n * sh.radius + pos;
    if (dot3(n, ref - p) < 0.0f)
    {
        n = n * -1.0f;
        p = n * sh.radius + pos;
    }
    return p;
}

__global__ void __launch_bounds__(128)
k_s23_shade(const __grid_constant__ S23Ctx c)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= c.num_samples)
        return;
    uint64_t offset = 2ull * k + (uint64_t)c.draws_per_hit * c.hits_before[k];
    MwcState rng = s23_rng_at(c, offset);
    float r1 = s23_next_float(rng), r2 = s23_next_float(rng);
    float xu, yu;
    s23_screen(c, k, r1, r2, xu, yu);
    V3 o, d;
    s23_camera_ray(c, xu, yu, o, d);
    S23Hit hit;
    Color3 term[RT_S23_MAX_TERMS];
    for (int t = 0; t < RT_S23_MAX_TERMS; ++t) term[t] = mkc(0.0f, 0.0f, 0.0f);
    if (s23_intersect<true>(c, o, d, RT_RAY_TMAX, hit))
    {
        const RtS23Material& mat = c.materials[c.shapes[hit.shape].material];
        Color3 emit = mkc(mat.emittance[0], mat.emittance[1], mat.emittance[2]);
        V3 position = o + hit.t * d;
        Color3 cm = mkc(hit.cmod, hit.cmod, hit.cmod);
        if (c.stage == 2)
        {
            // Rayito_Stage2/main.cpp:169-203: one purely random sample per light, every
            // contribution added straight into the pixel accumulator
            term[0] = emit;
            Color3 surf = mkc(mat.color[0], mat.color[1], mat.color[2]) * cm;     // m_color (bullseye folded in, rayito.h:657)
            for (uint32_t l = 0; l < c.num_lights; ++l)
            {
                const RtS23Shape& ls = c.shapes[c.lights[l]];
                const RtS23Material& lm = c.materials[ls.material];
                // g++ evaluates the two nextFloat() arguments right to left: u2 is drawn first
                float u2 = s23_next_float(rng);
                float u1 = s23_next_float(rng);
                V3 lp = s23_sample_light(ls, u1, u2, position);
                float dist;
                V3 to_light = s23_normalized(lp - position, &dist);
                S23Hit sh;
                bool blocked = s23_intersect<false>(c, position, to_light, dist, sh);
                if (!blocked || (sh.is_light_self && sh.shape == (int)c.lights[l]))
                {
                    float atten = std_max(0.0f, dot3(hit.normal, to_light));
                    Color3 e = mkc(lm.emittance[0], lm.emittance[1], lm.emittance[2]);
                    Color3 v = surf * e;
                    term[1 + l] = mkc(atten * v.r, atten * v.g, atten * v.b);
                }
            }
        }
        else
        {
            // trace() (Rayito_Stage3/main.cpp:96-159)
            Color3 result = mkc(0.0f, 0.0f, 0.0f) + emit;
            for (uint32_t l = 0; l < c.num_lights; ++l)
            {
                const RtS23Shape& ls = c.shapes[c.lights[l]];
                const RtS23Material& lm = c.materials[ls.material];
                Color3 e = mkc(lm.emittance[0], lm.emittance[1], lm.emittance[2]);
                Color3 light_result = mkc(0.0f, 0.0f, 0.0f);
                for (uint32_t lsv = 0; lsv < c.lv; ++lsv)
                {
                    for (uint32_t lsu = 0; lsu < c.lu; ++lsu)
                    {
                        // arguments evaluated right to left: the v sample is drawn first
                        float u2 = ((float)lsv + s23_next_float(rng)) / (float)c.lv;
                        float u1 = ((float)lsu + s23_next_float(rng)) / (float)c.lu;
                        V3 lp = s23_sample_light(ls, u1, u2, position);
                        float dist;
                        V3 to_light = s23_normalized(lp - position, &dist);
                        S23Hit sh;
                        bool blocked = s23_intersect<false>(c, position, to_light, dist, sh);
                        if (!blocked || (sh.is_light_self && sh.shape == (int)c.lights[l]))
                            light_result = light_result + e * cm * s23_shade(mat, hit.normal, d, to_light);
                    }
                }
                light_result = light_result / (float)(c.lu * c.lv);
                result = result + light_result;
            }
            term[0] = result;
        }
    }
    for (uint32_t t = 0; t < c.terms; ++t)
        c.sample_terms[(uint64_t)t * c.num_samples + k] = make_float4(term[t].r, term[t].g, term[t].b, 0.0f);
}

// ---- 5. pixel: ordered sum, box filter, clamp, quantise (main.cpp:256-268) ---------------
__global__ void __launch_bounds__(128)
k_s23_resolve(const __grid_constant__ S23Ctx c)
{
    uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (uint64_t)c.width * c.height)
        return;
    Color3 sum = mkc(0.0f, 0.0f, 0.0f);
    for (uint32_t s = 0; s < c.spp; ++s)
    {
        uint64_t k = p * c.spp + s;
        for (uint32_t t = 0; t < c.terms; ++t)
        {
            float4 v = c.sample_terms[(uint64_t)t * c.num_samples + k];
            sum = sum + mkc(v.x, v.y, v.z);
        }
    }
    sum = sum / (float)c.spp;
    if (c.rgb)
    {
        c.rgb[3 * p + 0] = sum.r; c.rgb[3 * p + 1] = sum.g; c.rgb[3 * p + 2] = sum.b;
    }
    if (c.rgb8)
    {
        float r = std_max(0.0f, std_min(1.0f, sum.r));
        float g = std_max(0.0f, std_min(1.0f, sum.g));
        float b = std_max(0.0f, std_min(1.0f, sum.b));
        c.rgb8[3 * p + 0] = (uint8_t)(r * 255.0f);
        c.rgb8[3 * p + 1] = (uint8_t)(g * 255.0f);
        c.rgb8[3 * p + 2] = (uint8_t)(b * 255.0f);
    }
}

// ---- host side -----------------------------------------------------------------------
inline int rt_stage23_impl(int device, const RtS23Scene* scene, const RtCamera* cam, const RtS23Params* prm,
                           float* rgb, uint8_t* rgb8, RtRenderStats* stats)
{
    if (scene == NULL || cam == NULL || prm == NULL || (rgb == NULL && rgb8 == NULL))
        return rt_fail(RT_ERR_ARG, "null argument");
    if (prm->stage != 2 && prm->stage != 3)
        return rt_fail(RT_ERR_ARG, "RtS23Params.stage must be 2 or 3");
    if (scene->num_shapes == 0 || scene->num_shapes > RT_S23_MAX_SHAPES || scene->num_lights > RT_S23_MAX_LIGHTS ||
        scene->num_materials == 0 || scene->num_materials > RT_S23_MAX_SHAPES)
        return rt_fail(RT_ERR_ARG, "Stage 2/3 scenes hold 1..16 shapes and materials and at most 4 lights");
    if (prm->width < 2 || prm->height < 2 || prm->pixel_samples_u == 0 || prm->pixel_samples_v == 0)
        return rt_fail(RT_ERR_ARG, "bad image size or sample count");
    if (prm->stage == 2 && scene->num_lights + 1 > RT_S23_MAX_TERMS)
        return rt_fail(RT_ERR_ARG, "Stage 2 supports at most 2 lights");
    if (prm->stage == 3 && (prm->light_samples_u == 0 || prm->light_samples_v == 0))
        return rt_fail(RT_ERR_ARG, "Stage 3 needs light_samples_u, light_samples_v >= 1");
    for (uint32_t i = 0; i < scene->num_shapes; ++i)
        if (scene->shapes[i].type > RT_S23_RECT || scene->shapes[i].material >= scene->num_materials)
            return rt_fail(RT_ERR_ARG, "bad shape type or material index");
    for (uint32_t i = 0; i < scene->num_lights; ++i)
        if (scene->lights[i] >= scene->num_shapes)
            return rt_fail(RT_ERR_ARG, "light index out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return rt_fail(RT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    RT_CUDA(cudaSetDevice(device));

    S23Ctx c;
    std::memset(&c, 0, sizeof(c));
    std::memcpy(c.shapes, scene->shapes, sizeof(RtS23Shape) * scene->num_shapes);
    std::memcpy(c.materials, scene->materials, sizeof(RtS23Material) * scene->num_materials);
    std::memcpy(c.lights, scene->lights, sizeof(uint32_t) * scene->num_lights);
    c.num_shapes = scene->num_shapes;
    c.num_lights = scene->num_lights;
    c.cam = *cam;
    c.stage = prm->stage;
    c.width = prm->width;
    c.height = prm->height;
    c.nu = prm->pixel_samples_u;
    c.nv = prm->stage == 2 ? 1 : prm->pixel_samples_v;
    c.spp = c.nu * c.nv;
    c.lu = prm->stage == 2 ? 1 : prm->light_samples_u;
    c.lv = prm->stage == 2 ? 1 : prm->light_samples_v;
    c.draws_per_hit = 2 * c.lu * c.lv * c.num_lights;
    c.terms = prm->stage == 2 ? 1 + c.num_lights : 1;
    c.num_samples = (uint64_t)c.width * c.height * c.spp;
    if (c.num_samples >= (1ull << 32))
        return rt_fail(RT_ERR_ARG, "too many pixel samples for one call (2^32)");
    c.seed.z = prm->seed_z;
    c.seed.w = prm->seed_w;
    c.seed_after2 = c.seed;
    mwc_next(c.seed_after2);
    mwc_next(c.seed_after2);

    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    // one resident block per segment: the segments of a round run side by side
    uint32_t want = (uint32_t)sms * 4;
    c.seg_len = (c.num_samples + want - 1) / want;
    c.seg_len = ((c.seg_len + RT_S23_WINDOW - 1) / RT_S23_WINDOW) * RT_S23_WINDOW;
    c.num_segs = (uint32_t)((c.num_samples + c.seg_len - 1) / c.seg_len);

    const size_t n = (size_t)c.num_samples, px = (size_t)c.width * c.height;
    size_t bytes_flags = (n + 255) & ~(size_t)255;
    size_t bytes_hb = n * 4, bytes_terms = n * 16 * c.terms, bytes_seg = (size_t)(c.num_segs + 1) * 8;
    size_t total = bytes_flags + bytes_hb + bytes_terms + 2 * bytes_seg + (size_t)(c.num_segs + 64) * 4 + px * 12 + px * 3 + 1024;
    char* block = NULL;
    cudaEvent_t ev[3] = { NULL, NULL, NULL };
    cudaError_t err = cudaMalloc((void**)&block, total);
    if (err != cudaSuccess)
        return rt_fail(RT_ERR_CUDA, cudaGetErrorString(err));
    int rc = RT_OK;
    uint64_t launches = 0, rounds = 0;
    do
    {
        char* q = block;
        c.sample_terms = (float4*)q;                q += bytes_terms;
        c.hits_before = (uint32_t*)q;               q += bytes_hb;
        c.seg_hin = (unsigned long long*)q;         q += bytes_seg;
        c.seg_hout = (unsigned long long*)q;        q += bytes_seg;
        c.rgb = (float*)q;                          q += px * 12;
        c.seg_dirty = (uint32_t*)q;                 q += (size_t)c.num_segs * 4;
        c.any_dirty = (uint32_t*)q;                 q += 64 * 4;
        c.flags = (uint8_t*)q;                      q += bytes_flags;
        c.rgb8 = (uint8_t*)q;
        for (int i = 0; i < 3; ++i)
            if ((err = cudaEventCreate(&ev[i])) != cudaSuccess) break;
        if (err != cudaSuccess) break;
        cudaEventRecord(ev[0]);
        if ((err = cudaMemsetAsync(c.seg_hin, 0, 2 * bytes_seg)) != cudaSuccess) break;
        unsigned sample_blocks = (unsigned)((n + 255) / 256);
        k_s23_guess<<<sample_blocks, 256>>>(c);
        k_s23_segfix<<<1, 1>>>(c);
        launches += 2;
        // every segment runs at least once (segfix marks only changed ones)
        if ((err = cudaMemsetAsync(c.seg_dirty, 1, (size_t)c.num_segs * 4)) != cudaSuccess) break;
        for (;;)
        {
            k_s23_prepass<<<c.num_segs, RT_S23_PRE_THREADS>>>(c);
            k_s23_segfix<<<1, 1>>>(c);
            launches += 2;
            ++rounds;
            uint32_t dirty = 0;
            if ((err = cudaMemcpy(&dirty, c.any_dirty, 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
            if (!dirty) break;
            if (rounds > c.num_segs + 2) { rc = rt_fail(RT_ERR_UNSUPPORTED, "Stage 2/3 stream pre-pass did not converge"); break; }
        }
        if (err != cudaSuccess || rc != RT_OK) break;
        cudaEventRecord(ev[1]);
        k_s23_shade<<<(unsigned)((n + 127) / 128), 128>>>(c);
        k_s23_resolve<<<(unsigned)((px + 127) / 128), 128>>>(c);
        launches += 2;
        cudaEventRecord(ev[2]);
        if ((err = cudaGetLastError()) != cudaSuccess) break;
        if (rgb && (err = cudaMemcpy(rgb, c.rgb, px * 12, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (rgb8 && (err = cudaMemcpy(rgb8, c.rgb8, px * 3, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if ((err = cudaDeviceSynchronize()) != cudaSuccess) break;
        if (stats)
        {
            std::memset(stats, 0, sizeof(*stats));
            unsigned long long hits = 0;
            if ((err = cudaMemcpy(&hits, c.seg_hout + (c.num_segs - 1), 8, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
            float pre_ms = 0.0f, shade_ms = 0.0f;
            cudaEventElapsedTime(&pre_ms, ev[0], ev[1]);
            cudaEventElapsedTime(&shade_ms, ev[1], ev[2]);
            stats->samples = c.num_samples;
            // every ShapeSet::intersect call: camera rays + one shadow ray per light sample of a hit
            stats->closest_rays = c.num_samples + hits * (uint64_t)(c.draws_per_hit / 2);
            stats->shape_tests = stats->closest_rays * c.num_shapes;
            stats->kernel_launches = launches;
            stats->trace_launches = rounds;          // pre-pass rounds
            stats->render_ms = pre_ms + shade_ms;
            stats->trace_ms = shade_ms;
            stats->upload_ms = pre_ms;               // stream pre-pass (guess + fix-point rounds)
        }
    } while (0);
    for (int i = 0; i < 3; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
    cudaFree(block);
    if (err != cudaSuccess)
    {
        cudaGetLastError();
        return rt_fail(RT_ERR_CUDA, cudaGetErrorString(err));
    }
    return rc;
}

} // namespace rt_detail

#endif // RAYITO_B200_RT_STAGE23_CUH
