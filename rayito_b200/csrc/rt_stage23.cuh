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
//   2. k_s23_prepass  -- per image segment (one warp each), a windowed fix-point: evaluate
//      a window of 32 samples with offsets from the assumed flags, accept everything up to
//      the first disagreement, restart behind it.  Samples whose hit does not depend on the jitter
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

#include <mutex>

#include "rt_pool.cuh"
#include "rt_sampling.cuh"
#include "rt_shade.cuh"

namespace rt_detail
{

#define RT_S23_TMIN 0.00001f          /* kRayTMin of Stages 1-3 (Rayito_Stage3/rayito.h:303) */
#define RT_S23_MAX_TERMS 3            /* Stage 2: emitted + one term per light (<= 2 lights) */

// Pure functions of a shape's constants that the reference evaluates inside every intersect
// call (rayito.h:618,634-637,747,826).  Same inputs, same IEEE operations => same bits, so they
// are computed once per render (on the host, in float, without contraction).
struct S23Derived
{
    float n[3];             // rectangle: cross(side1, side2).normalized()
    float pos_dot_n;        // dot(m_position, normal) (plane and rectangle)
    float s1n[3], s2n[3];   // rectangle: normalised sides
    float len1, len2;       // ... and their lengths
    float r2;               // sphere: m_radius * m_radius
};

struct S23Ctx
{
    RtS23Shape shapes[RT_S23_MAX_SHAPES];
    RtS23Material materials[RT_S23_MAX_SHAPES];
    S23Derived derived[RT_S23_MAX_SHAPES];      // per-shape values the reference recomputes on every call
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
    uint32_t* seg_sensitive;                     // scheduling hint from k_s23_guess
    const uint32_t* pow_table;                   // MWC multiplier powers for short jumps (see k_s23_prepass)
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
    // num_samples < 2^32: 32-bit divisions (64-bit ones cost ten times as much on the serial chain)
    uint32_t k32 = (uint32_t)k;
    uint32_t p = k32 / c.spp, s = k32 - p * c.spp;
    uint32_t y = p / c.width, x = p - y * c.width;
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
// only t < current m_t.  SHADING = also fill normal / colour modifier.  ANY = only the
// boolean result is wanted: the first accepted hit decides it.
template <bool SHADING, bool ANY>
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
        const S23Derived& dv = c.derived[i];
        V3 pos = mk(sh.position[0], sh.position[1], sh.position[2]);
        if (sh.type == RT_S23_PLANE)
        {
            // Plane::intersect (rayito.h:745-780): one-sided
            V3 n = mk(sh.normal[0], sh.normal[1], sh.normal[2]);
            float n_dot_d = dot3(n, d);
            if (n_dot_d >= 0.0f)
                continue;
            float t = (dv.pos_dot_n - dot3(o, n)) / dot3(d, n);
            if (t >= h.t || t < RT_S23_TMIN)
                continue;
            if (ANY) return true;
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
            float cc = (lo.x * lo.x + lo.y * lo.y + lo.z * lo.z) - dv.r2;
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
            if (ANY) return true;
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
            V3 n = mk(dv.n[0], dv.n[1], dv.n[2]);
            float n_dot_d = dot3(n, d);
            if (n_dot_d == 0.0f)
                continue;
            float t = (dv.pos_dot_n - dot3(o, n)) / dot3(d, n);
            if (t >= h.t || t < RT_S23_TMIN)
                continue;
            V3 s1n = mk(dv.s1n[0], dv.s1n[1], dv.s1n[2]), s2n = mk(dv.s2n[0], dv.s2n[1], dv.s2n[2]);
            V3 rel = (o + t * d) - pos;
            float lx = dot3(rel, s1n), ly = dot3(rel, s2n);
            if (lx < 0.0f || lx > dv.len1 || ly < 0.0f || ly > dv.len2)
                continue;
            if (ANY) return true;
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

// ---- 1. geometric guess ------------------------------------------------------------
// Hit predicate with centred jitter, and whether the four jitter corners agree with it.
// A segment with a disagreeing sample is "sensitive": its hit count may depend on where
// in the stream it starts.  This is only a scheduling hint (k_s23_chain); every sample
// is still confirmed at its true stream position by k_s23_prepass.
__device__ __forceinline__ bool s23_hits_with_jitter(const S23Ctx& c, uint64_t k, float r1, float r2)
{
    float xu, yu;
    s23_screen(c, k, r1, r2, xu, yu);
    V3 o, d;
    s23_camera_ray(c, xu, yu, o, d);
    S23Hit h;
    return s23_intersect<false, true>(c, o, d, RT_RAY_TMAX, h);
}

__global__ void __launch_bounds__(256)
k_s23_guess(const __grid_constant__ S23Ctx c)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false, sensitive = false;
    if (k < c.num_samples)
    {
        hit = s23_hits_with_jitter(c, k, 0.5f, 0.5f);
        const float lo = 0.0f, hi = 0.99999994f;
        sensitive = s23_hits_with_jitter(c, k, lo, lo) != hit || s23_hits_with_jitter(c, k, lo, hi) != hit ||
                    s23_hits_with_jitter(c, k, hi, lo) != hit || s23_hits_with_jitter(c, k, hi, hi) != hit;
        c.flags[k] = hit ? 1 : 0;
    }
    // per-segment totals of the guess seed the segment chain (seg_hin = 0 here); a warp's 32
    // samples lie in one segment (segment lengths are multiples of 32)
    uint32_t votes = __ballot_sync(0xffffffffu, hit);
    uint32_t touchy = __ballot_sync(0xffffffffu, sensitive);
    if ((threadIdx.x & 31) == 0 && k < c.num_samples)
    {
        uint64_t seg = k / c.seg_len;
        if (votes) atomicAdd(c.seg_hout + seg, (unsigned long long)__popc(votes));
        if (touchy) atomicOr(c.seg_sensitive + seg, 1u);
    }
}

// ---- 2. windowed fix-point over one segment -------------------------------------------
// One WARP per segment, windows of 32 samples (one per lane): no shared-memory traffic and
// no block barriers on the serial chain.  Assumed flags travel in registers from one window
// to the next.  Rng states come from one state per segment moved by table look-ups: a draw
// distance below 2^20 costs three modular multiplications per generator
// (a^lo * a^(64 mid) * a^(4096 hi) * state) instead of a 34-step modular exponentiation.
#define RT_S23_PRE_WARPS 4                                   /* warps (= segments) per block */
#define RT_S23_TABLE 384                                     /* 64 + 64 + 256 powers per generator */
#define RT_S23_TABLE_SPAN (1u << 20)

__device__ __forceinline__ uint64_t s23_mulmod(uint64_t a, uint64_t b, uint64_t m) { return (a * b) % m; }

struct S23Base
{
    uint64_t off;           // draws consumed before this state (>= 2: canonical)
    uint64_t zr, wr;        // state residues
    bool z_nonzero, w_nonzero;
};

__device__ __forceinline__ S23Base s23_base_at(const S23Ctx& c, uint64_t off)
{
    S23Base b;
    b.off = off < 2 ? 2 : off;
    MwcState s = s23_rng_at(c, b.off);
    b.zr = s.z % RT_MWC_MZ; b.wr = s.w % RT_MWC_MW;
    b.z_nonzero = s.z != 0; b.w_nonzero = s.w != 0;
    return b;
}

// State `delta` draws after the base (delta < 2^20)
__device__ __forceinline__ MwcState s23_rng_ahead(const S23Base& b, uint32_t delta, const uint32_t* tz, const uint32_t* tw)
{
    uint32_t lo = delta & 63u, mid = 64u + ((delta >> 6) & 63u), hi = 128u + (delta >> 12);
    uint64_t z = s23_mulmod(s23_mulmod(s23_mulmod(tz[lo], tz[mid], RT_MWC_MZ), tz[hi], RT_MWC_MZ), b.zr, RT_MWC_MZ);
    uint64_t w = s23_mulmod(s23_mulmod(s23_mulmod(tw[lo], tw[mid], RT_MWC_MW), tw[hi], RT_MWC_MW), b.wr, RT_MWC_MW);
    MwcState r;
    r.z = (uint32_t)((z == 0 && b.z_nonzero) ? RT_MWC_MZ : z);     // same representative rule as mwc_jump
    r.w = (uint32_t)((w == 0 && b.w_nonzero) ? RT_MWC_MW : w);
    return r;
}

// Resolves samples [begin, end) given H hits before `begin`; returns the hits before `end`.
// All 32 lanes of the calling warp take part.
__device__ __forceinline__ uint64_t s23_resolve_segment(const S23Ctx& c, const uint32_t* tz, const uint32_t* tw,
                                                        uint64_t begin, uint64_t end, uint64_t H)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t D = c.draws_per_hit;
    const uint32_t reach = 32u * (2u + D);                   // draws one window can span
    uint64_t P = begin;
    S23Base base = s23_base_at(c, 2ull * P + (uint64_t)D * H);
    uint32_t a = P + lane < end ? c.flags[P + lane] : 0u;
    while (P < end)
    {
        const uint64_t win_off = 2ull * P + (uint64_t)D * H;
        if (win_off - base.off + reach >= RT_S23_TABLE_SPAN && win_off >= base.off)
            base = s23_base_at(c, win_off);
        const uint64_t k = P + lane;
        const bool live = k < end;
        // stored flags the next window may need (it starts at most 32 samples further on)
        const uint32_t ahead = P + 32 + lane < end ? c.flags[P + 32 + lane] : 0u;
        const uint32_t votes_a = __ballot_sync(0xffffffffu, a != 0);
        const uint32_t before = __popc(votes_a & ((1u << lane) - 1u));
        const uint64_t off = 2ull * k + (uint64_t)D * (H + before);
        bool e = false;
        if (live)
        {
            MwcState st;
            if (off < base.off)               st = s23_rng_at(c, off);                  // the very first sample
            else if (reach < RT_S23_TABLE_SPAN) st = s23_rng_ahead(base, (uint32_t)(off - base.off), tz, tw);
            else                              st = s23_rng_at(c, off);                  // absurdly many light samples
            float r1 = s23_next_float(st), r2 = s23_next_float(st);
            e = s23_hits_with_jitter(c, k, r1, r2);
        }
        const uint32_t votes_e = __ballot_sync(0xffffffffu, e);
        const uint32_t live_mask = __ballot_sync(0xffffffffu, live);
        const uint32_t diff = (votes_e ^ votes_a) & live_mask;
        // lanes up to and including the first disagreement ran at their true offsets: final
        const uint32_t m = diff ? (uint32_t)__ffs(diff) - 1u : 31u;
        const uint32_t final_mask = (m == 31u ? 0xffffffffu : ((2u << m) - 1u)) & live_mask;
        if (live && lane <= m)
        {
            c.flags[k] = e ? 1 : 0;
            c.hits_before[k] = (uint32_t)(H + before);
        }
        const uint32_t adv = __popc(final_mask);
        H += __popc(votes_e & final_mask);
        P += adv;
        // next window's assumption: what the lanes behind m just evaluated, then stored flags
        const uint32_t keep = 32u - adv;
        const uint32_t carried = __shfl_down_sync(0xffffffffu, e ? 1u : 0u, adv & 31u);
        const uint32_t fresh = __shfl_sync(0xffffffffu, ahead, (lane - keep) & 31u);
        a = lane < keep ? carried : fresh;
    }
    return H;
}

__device__ __forceinline__ void s23_load_tables(const S23Ctx& c, uint32_t* tz, uint32_t* tw)
{
    for (uint32_t i = threadIdx.x; i < RT_S23_TABLE; i += blockDim.x)
    {
        tz[i] = c.pow_table[i];
        tw[i] = c.pow_table[RT_S23_TABLE + i];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RT_S23_PRE_WARPS * 32)
k_s23_prepass(const __grid_constant__ S23Ctx c)
{
    __shared__ uint32_t tz[RT_S23_TABLE], tw[RT_S23_TABLE];
    s23_load_tables(c, tz, tw);
    const uint32_t seg = blockIdx.x * RT_S23_PRE_WARPS + (threadIdx.x >> 5);
    if (seg >= c.num_segs || !c.seg_dirty[seg])
        return;
    const uint64_t begin = (uint64_t)seg * c.seg_len;
    const uint64_t end = begin + c.seg_len < c.num_samples ? begin + c.seg_len : c.num_samples;
    uint64_t H = s23_resolve_segment(c, tz, tw, begin, end, c.seg_hin[seg]);
    if ((threadIdx.x & 31) == 0)
        c.seg_hout[seg] = H;
}

// One warp walks the segments in order: insensitive ones contribute their guessed count,
// sensitive ones are resolved in place with the exact incoming count.  Afterwards seg_hin /
// seg_hout are exact provided the hint was right; k_s23_prepass + k_s23_segfix check that.
__global__ void __launch_bounds__(32)
k_s23_chain(const __grid_constant__ S23Ctx c)
{
    __shared__ uint32_t tz[RT_S23_TABLE], tw[RT_S23_TABLE];
    s23_load_tables(c, tz, tw);
    const uint32_t lane = threadIdx.x & 31;
    // with silhouettes everywhere a serial walk would cost more than the parallel rounds it saves
    uint32_t sensitive_segs = 0;
    for (uint32_t s0 = 0; s0 < c.num_segs; s0 += 32)
        sensitive_segs += __popc(__ballot_sync(0xffffffffu, s0 + lane < c.num_segs && c.seg_sensitive[s0 + lane] != 0));
    if (sensitive_segs > 64 && sensitive_segs > c.num_segs / 8)
        return;
    unsigned long long H = 0;
    for (uint32_t s0 = 0; s0 < c.num_segs; s0 += 32)
    {
        const uint32_t s = s0 + lane;
        const bool in = s < c.num_segs;
        unsigned long long count = in ? c.seg_hout[s] - c.seg_hin[s] : 0ull;
        const uint32_t touchy = __ballot_sync(0xffffffffu, in && c.seg_sensitive[s] != 0);
        for (uint32_t j = 0; j < 32 && s0 + j < c.num_segs; ++j)
        {
            const uint32_t sj = s0 + j;
            unsigned long long cj = __shfl_sync(0xffffffffu, count, j);
            if (lane == 0) c.seg_hin[sj] = H;
            if (touchy & (1u << j))
            {
                const uint64_t begin = (uint64_t)sj * c.seg_len;
                const uint64_t end = begin + c.seg_len < c.num_samples ? begin + c.seg_len : c.num_samples;
                H = s23_resolve_segment(c, tz, tw, begin, end, H);
            }
            else
                H += cj;
            if (lane == 0) c.seg_hout[sj] = H;
        }
    }
}

// ---- 3. chain the segments -------------------------------------------------------------
// One block: exclusive scan of the segments' hit counts -> hits before each segment; a
// segment whose incoming count changed must be redone.
#define RT_S23_FIX_THREADS 1024
__global__ void __launch_bounds__(RT_S23_FIX_THREADS)
k_s23_segfix(const __grid_constant__ S23Ctx c)
{
    __shared__ unsigned long long warp_sum[RT_S23_FIX_THREADS / 32];
    __shared__ uint32_t dirty_any;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) dirty_any = 0;
    const uint32_t per = (c.num_segs + RT_S23_FIX_THREADS - 1) / RT_S23_FIX_THREADS;
    const uint32_t first = tid * per, last = first + per < c.num_segs ? first + per : c.num_segs;
    unsigned long long mine = 0;
    for (uint32_t s = first; s < last; ++s)
        mine += c.seg_hout[s] - c.seg_hin[s];
    unsigned long long incl = mine;
    #pragma unroll
    for (int off = 1; off < 32; off <<= 1)
    {
        unsigned long long v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (uint32_t)off) incl += v;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    unsigned long long h = incl - mine;
    for (uint32_t w = 0; w < warp; ++w) h += warp_sum[w];
    uint32_t dirty = 0;
    for (uint32_t s = first; s < last; ++s)
    {
        unsigned long long count = c.seg_hout[s] - c.seg_hin[s];
        uint32_t d = c.seg_hin[s] != h ? 1u : 0u;
        c.seg_dirty[s] = d;
        dirty |= d;
        c.seg_hin[s] = h;
        c.seg_hout[s] = h + count;
        h += count;
    }
    if (dirty) atomicOr(&dirty_any, 1u);
    __syncthreads();
    if (tid == 0) *c.any_dirty = dirty_any;
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
    if (s23_intersect<true, false>(c, o, d, RT_RAY_TMAX, hit))
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
                bool blocked = s23_intersect<false, false>(c, position, to_light, dist, sh);
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
                        bool blocked = s23_intersect<false, false>(c, position, to_light, dist, sh);
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
// volatile stores keep each intermediate a rounded float (no excess precision, no contraction)
inline float s23_host_dot(const float* a, const float* b)
{
    volatile float x = a[0] * b[0], y = a[1] * b[1], z = a[2] * b[2];
    volatile float xy = x + y;
    volatile float r = xy + z;
    return r;
}

inline float s23_host_normalize(const float* v, float* out)
{
    volatile float len = sqrtf(s23_host_dot(v, v));
    for (int i = 0; i < 3; ++i) { volatile float q = v[i] / len; out[i] = q; }
    return len;
}

inline S23Derived s23_derive(const RtS23Shape& sh)
{
    S23Derived d;
    std::memset(&d, 0, sizeof(d));
    if (sh.type == RT_S23_PLANE)
        d.pos_dot_n = s23_host_dot(sh.position, sh.normal);
    else if (sh.type == RT_S23_SPHERE)
    {
        volatile float r2 = sh.radius * sh.radius;
        d.r2 = r2;
    }
    else
    {
        const float* a = sh.side1; const float* b = sh.side2;
        volatile float cx1 = a[1] * b[2], cx2 = a[2] * b[1], cy1 = a[2] * b[0], cy2 = a[0] * b[2], cz1 = a[0] * b[1], cz2 = a[1] * b[0];
        volatile float cx = cx1 - cx2, cy = cy1 - cy2, cz = cz1 - cz2;
        float cr[3] = { cx, cy, cz };
        s23_host_normalize(cr, d.n);
        d.pos_dot_n = s23_host_dot(sh.position, d.n);
        d.len1 = s23_host_normalize(sh.side1, d.s1n);
        d.len2 = s23_host_normalize(sh.side2, d.s2n);
    }
    return d;
}

// Working set kept between calls (one block, grow only)
struct S23Cache
{
    std::mutex lock;
    char* block;
    size_t bytes;
    int device;
    S23Cache() : block(NULL), bytes(0), device(-1) { }
};
inline S23Cache& s23_cache() { static S23Cache c; return c; }
inline void s23_release_locked(S23Cache& c)
{
    if (c.block != NULL)
    {
        cudaSetDevice(c.device);
        cudaDeviceSynchronize();
        pool_free(c.device, c.block, c.bytes);
    }
    c.block = NULL;
    c.bytes = 0;
    c.device = -1;
}
// rt_release_cached_memory(): park the Stage 2/3 working set in the pool, which is released next
inline void s23_release()
{
    S23Cache& c = s23_cache();
    std::lock_guard<std::mutex> guard(c.lock);
    s23_release_locked(c);
}

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
    for (uint32_t i = 0; i < scene->num_shapes; ++i)
        c.derived[i] = s23_derive(scene->shapes[i]);
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
    // one resident warp per segment: the segments of a round run side by side
    uint32_t want = (uint32_t)sms * 32;
    c.seg_len = (c.num_samples + want - 1) / want;
    c.seg_len = ((c.seg_len + 31) / 32) * 32;
    c.num_segs = (uint32_t)((c.num_samples + c.seg_len - 1) / c.seg_len);

    // a^i (i < 64), a^(64 j) (j < 64) and a^(4096 l) (l < 256) for both generators, filled once
    static uint32_t table[2 * RT_S23_TABLE];
    static std::once_flag table_once;
    std::call_once(table_once, []() {
        for (int g = 0; g < 2; ++g)
        {
            uint64_t a = g == 0 ? RT_MWC_AZ : RT_MWC_AW, m = g == 0 ? RT_MWC_MZ : RT_MWC_MW;
            for (uint32_t i = 0; i < 64; ++i)
            {
                table[g * RT_S23_TABLE + i] = (uint32_t)mwc_powmod(a, i, m);
                table[g * RT_S23_TABLE + 64 + i] = (uint32_t)mwc_powmod(a, 64ull * i, m);
            }
            for (uint32_t i = 0; i < 256; ++i)
                table[g * RT_S23_TABLE + 128 + i] = (uint32_t)mwc_powmod(a, 4096ull * i, m);
        }
    });

    const size_t n = (size_t)c.num_samples, px = (size_t)c.width * c.height;
    const size_t A = 255;       // every sub-buffer starts on a 256-byte boundary
    size_t bytes_flags = (n + A) & ~A;
    size_t bytes_hb = (n * 4 + A) & ~A, bytes_terms = (n * 16 * c.terms + A) & ~A;
    size_t bytes_seg = ((size_t)(c.num_segs + 1) * 8 + A) & ~A, bytes_dirty = ((size_t)c.num_segs * 4 + A) & ~A;
    size_t bytes_rgb = (px * 12 + A) & ~A, bytes_rgb8 = (px * 3 + A) & ~A;
    size_t total = bytes_flags + bytes_hb + bytes_terms + 2 * bytes_seg + 2 * bytes_dirty + 256 + bytes_rgb + bytes_rgb8 +
                   sizeof(table);
    // working memory is kept between calls (grow only): a sweep re-renders the same image many times.
    // The block comes from the device pool, so rt_release_cached_memory() can hand it back (s23_release).
    S23Cache& cache = s23_cache();
    std::lock_guard<std::mutex> guard(cache.lock);
    cudaError_t err = cudaSuccess;
    if (cache.device != device || cache.bytes < total)
    {
        s23_release_locked(cache);
        err = pool_alloc(device, (void**)&cache.block, total, &cache.bytes);
        if (err != cudaSuccess)
        {
            cache.block = NULL;
            cache.bytes = 0;
            cudaGetLastError();
            return rt_fail(RT_ERR_CUDA, cudaGetErrorString(err));
        }
        cache.device = device;
    }
    char* block = cache.block;
    cudaEvent_t ev[3] = { NULL, NULL, NULL };
    int rc = RT_OK;
    uint64_t launches = 0, rounds = 0;
    do
    {
        char* q = block;
        c.sample_terms = (float4*)q;                q += bytes_terms;
        c.hits_before = (uint32_t*)q;               q += bytes_hb;
        c.seg_hin = (unsigned long long*)q;         q += bytes_seg;
        c.seg_hout = (unsigned long long*)q;        q += bytes_seg;
        c.rgb = (float*)q;                          q += bytes_rgb;
        c.seg_dirty = (uint32_t*)q;                 q += bytes_dirty;
        c.seg_sensitive = (uint32_t*)q;             q += bytes_dirty;
        c.any_dirty = (uint32_t*)q;                 q += 256;
        c.pow_table = (const uint32_t*)q;           q += sizeof(table);
        c.flags = (uint8_t*)q;                      q += bytes_flags;
        c.rgb8 = (uint8_t*)q;
        for (int i = 0; i < 3; ++i)
            if ((err = cudaEventCreate(&ev[i])) != cudaSuccess) break;
        if (err != cudaSuccess) break;
        if ((err = cudaMemcpy((void*)c.pow_table, table, sizeof(table), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        cudaEventRecord(ev[0]);
        if ((err = cudaMemsetAsync(c.seg_hin, 0, 2 * bytes_seg)) != cudaSuccess) break;
        if ((err = cudaMemsetAsync(c.seg_sensitive, 0, bytes_dirty)) != cudaSuccess) break;
        unsigned sample_blocks = (unsigned)((n + 255) / 256);
        k_s23_guess<<<sample_blocks, 256>>>(c);
        k_s23_segfix<<<1, RT_S23_FIX_THREADS>>>(c);
        k_s23_chain<<<1, 32>>>(c);
        launches += 3;
        // every segment runs at least once (segfix marks only changed ones)
        if ((err = cudaMemsetAsync(c.seg_dirty, 1, (size_t)c.num_segs * 4)) != cudaSuccess) break;
        for (;;)
        {
            k_s23_prepass<<<(c.num_segs + RT_S23_PRE_WARPS - 1) / RT_S23_PRE_WARPS, RT_S23_PRE_WARPS * 32>>>(c);
            k_s23_segfix<<<1, RT_S23_FIX_THREADS>>>(c);
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
    if (err != cudaSuccess)
    {
        cudaGetLastError();
        return rt_fail(RT_ERR_CUDA, cudaGetErrorString(err));
    }
    return rc;
}

} // namespace rt_detail

#endif // RAYITO_B200_RT_STAGE23_CUH
