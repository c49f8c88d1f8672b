// Wavefront path tracer: raytrace() / RenderThread::run() / pathTrace() of the
// reference (Rayito_Stage7_QT/RaytraceMain.cpp:64-185, 270-482, 485-579) as a
// sequence of wide kernels over a batch of pixel samples.
//
//   pixel setup   per pixel: the 5*depth+3 sampler permutations by MWC jump-ahead
//   ray gen       per sample: CMJ subpixel/lens/time -> camera ray          (:120-142)
//   per bounce b:
//     trace path  closest hit + shading inputs                              (:293-298)
//     shade       emission rule, material, bounce sampling, throughput      (:303-328, 451-477)
//     per light sample l:
//       light     pick light, sample its surface, BRDF-sample -> 2 rays     (:358-422)
//       trace     shadow rays (any hit) and BRDF-MIS probes (closest hit)   (:395, :423)
//       resolve   visibility, light pdf, MIS weights, accumulate            (:396-447)
//   accumulate    ordered per-pixel sum over samples, box filter            (:145-156)
//
// Paths live in fixed slots (sample index = pixel-in-batch * spp + psi, so the 32
// lanes of a warp start as 32 samples of one pixel); stages communicate through
// index queues filled with warp-aggregated atomics, so every kernel runs only on
// live work (ray compaction) and dead lanes never reach the traversal loop.
// All per-path arithmetic is order-independent of the queues: images are
// deterministic, and per-pixel sums run in ascending sample order like the CPU.
#ifndef RAYITO_B200_RT_RENDER_CUH
#define RAYITO_B200_RT_RENDER_CUH

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "rt_scene.cuh"
#include "rt_shade.cuh"
// threads per block of every kernel of the renderer (the split passes size shared arrays with it)
#ifndef RT_BLOCK
#define RT_BLOCK 128
#endif

#include "rt_wave.cuh"
#include "rt_split.cuh"

#define RT_DEFAULT_TILE 32u
#define RT_DEFAULT_BATCH (128u << 20)   /* samples per wavefront batch: ~88 GB of state on a 180 GB B200 (shrunk to fit
                                           60 % of the free memory elsewhere); 16 Mi -> 128 Mi: +5.7 % on C4 (kernel tails) */
#define RT_MAX_DEPTH 16u

// Queue counters.  Ray queues are binned by direction octant (8 bins) so that the
// lanes of a traversal warp share the near/far child order; the shade queue is binned
// by hit shape (16 bins) so that the lanes of a shading warp share shape type and BRDF.
#ifndef RT_QBINS
#define RT_QBINS 8
#endif
#define RT_SBINS 16
#define RT_LBINS 8
// minimum resident blocks per SM asked of the shading kernels (register cap = 65536 / (blocks * RT_BLOCK)):
// 6 blocks = 80 registers; k_light_sample wanted 96 (5 blocks), +2 % on the frame
#ifndef RT_SHADE_MINBLOCKS
#define RT_SHADE_MINBLOCKS 6
#endif
// Grid-stride kernels (ray generation, shading, light sampling, resolve): blocks per SM of their grids.  Round 1 used 32
// (606 k threads, each walking ~220 queue entries of a 128 Mi-sample batch).  Same-session A/B on C4 (sessions AB-AD,
// profiles/README.md): 6 / 12 / 24 per SM (whole waves of the 6 resident blocks) 4862-4878, 32: 4879-4925, 48: 4913-4960,
// 128: 4988, 512: 4980-5013, 1024: 4943, 2048: 4850, one entry per thread: 4344; with k_raygen at 128-512 as well 5025-5040
// (3840x2160: 4930 -> 5069); C5 2541 -> 2609.  More, shorter-lived blocks keep more gathers in flight and even out
// the end of every launch; past ~1000 per SM block scheduling itself starts to cost.
#ifndef RT_WIDE_SHADE_PER_SM
#define RT_WIDE_SHADE_PER_SM 512
#endif
#ifndef RT_WIDE_GEN_PER_SM
#define RT_WIDE_GEN_PER_SM 256
#endif
enum { CTL_PATH_A = 0, CTL_PATH_B = 8, CTL_SHADOW = 16, CTL_MIS = 24, CTL_SHADE = 32, CTL_LIT = 48,   // (room for 8 ray bins)
       CTL_CUR_PATH = 49, CTL_CUR_SHADOW = 50, CTL_CUR_MIS = 51,
       // split traversal: mesh / resume queue counts and cursors, double buffered
       CTL_MESH_N = 52, CTL_MESH_CUR = 54, CTL_RES_N = 56, CTL_RES_CUR = 58,
       // lit paths regrouped by (light, BRDF kind) for one light-sample iteration
       CTL_LITB = 64, CTL_WORDS = 80 };
#define CTL_PATH(cur) ((cur) ? CTL_PATH_B : CTL_PATH_A)

// Device pointers and constants of one render call
struct RenderCtx
{
    DScene sc;
    RtCamera cam;
    uint32_t width, height;
    uint32_t ps, ls, depth;
    uint32_t spp;               // ps * ps
    uint32_t spp_shift;         // log2(spp) when spp is a power of two, else 0xffffffff
    uint32_t nls;               // light-sample iterations per bounce: ls*ls (Stage 7, one random light each),
                                // num_lights*ls*ls (Stage 6, every light in turn); 0 without lights
    uint32_t ls2;               // ls * ls
    uint32_t tile, tiles_x;
    uint32_t num_pixels;        // pixels in this batch (tiles * tile^2, some may be off-image)
    uint32_t num_samples;       // num_pixels * spp
    const uint32_t* tile_ids;   // tiles of this batch
    float aspect;

    uint32_t* pix_xy;           // per pixel: y << 16 | x, or 0xffffffff
    uint32_t* perms;            // [(5*depth+3)][num_pixels]
    float4* ray_od;             // [2i] origin xyz, time; [2i+1] direction xyz, - (one 32-byte sector per path)
    float4* hit01;              // [2i] t, shape, tri record, -; [2i+1] normal xyz, colour modifier
    float4* thr;                // throughput rgb, (numBounces | numDirac << 8)
    float4* res;                // radiance rgb
    float4* pos_wo;             // [2i] hit position xyz, time; [2i+1] outgoing xyz, material index
    float4* lit_tr;             // [2i] throughput at the bounce being lit; [2i+1] lightResult accumulator
    // The two rays of the current light sample, one 32-byte sector each, so that a traversal pass
    // fetches a ray with ONE scattered sector read (origin and direction in separate arrays were two):
    // [4i+0] shadow direction xyz, tMax; [4i+1] origin xyz, time;
    // [4i+2] probe direction xyz, brdf pdf (0 = none); [4i+3] origin xyz, time (again).
    // Written whole (64 bytes) by k_light_sample.
    float4* lrec;
    // ... and what k_resolve adds up afterwards, 32 bytes written whole by k_light_sample:
    // [2i+0] light-sample term rgb, valid flag (cleared by an occluded shadow ray);
    // [2i+1] partial BRDF-sample term rgb, light shape id
    float4* lterm;
    float4* mis_hit0;
    float4* mis_hit1;
    float4* xf_cache;           // per-sample transform cache: sc.anim_stride float4 per sample (rt_device.cuh), or NULL
    uint32_t qcap;              // capacity of one queue bin (= samples of the batch buffers)
    uint32_t* q_path[2];        // RT_QBINS bins each
    uint32_t* q_shade;          // RT_SBINS bins: hit paths waiting for k_shade
    uint32_t* q_lit;
    uint32_t* q_shadow;
    uint32_t* q_mis;
    uint32_t* q_meshq[2];       // split traversal: slots waiting for a mesh pass
    uint32_t* q_resume[2];      // ... for a top-level resume pass
    SplitBufs split;            // suspended-ray state
    uint32_t* ctl;              // CTL_WORDS queue counters
    uint32_t* bin_hint;         // largest count seen so far per shade bin [0, RT_SBINS) and light bin [RT_SBINS, +RT_LBINS): sizes the next batches' grids
    uint64_t* totals;           // 0 closest rays, 1 any rays, 2..7 work counters (RT_WORK_COUNTERS)
    float* image;               // width*height*3, or (packed != 0) this rank's tiles one after the other
    uint32_t packed;            // tile-packed output: pixel (tile slot k, row r, column q) at ((k * tile + r) * tile + q) * 3
    uint32_t tile_base;         // packed: slot of this batch's first tile in the rank's tile list
};

// Sample i's row of the per-sample transform cache (NULL when the scene has no animated transform)
__device__ __forceinline__ const float4* xf_row(const RenderCtx& c, uint32_t i)
{
    return c.xf_cache ? c.xf_cache + (size_t)i * c.sc.anim_stride : nullptr;
}

// Sample index -> (pixel of the batch, sample of the pixel); a shift and a mask for power-of-two sample counts
__device__ __forceinline__ void sample_split(const RenderCtx& c, uint32_t i, uint32_t& p, uint32_t& psi)
{
    if (c.spp_shift != 0xffffffffu)
    {
        p = i >> c.spp_shift;
        psi = i & (c.spp - 1u);
    }
    else
    {
        p = i / c.spp;
        psi = i % c.spp;
    }
}

// Where pixel p of the batch (image coordinates x, y) is written
__device__ __forceinline__ float* pixel_out(const RenderCtx& c, uint32_t p, uint32_t x, uint32_t y)
{
    if (c.packed)
        return c.image + ((size_t)c.tile_base * c.tile * c.tile + p) * 3;
    return c.image + ((size_t)y * c.width + x) * 3;
}

struct RenderBuffers
{
    size_t cap_samples, cap_pixels, cap_tiles;
    uint32_t cap_perm_slots;
    uint32_t cap_anim_stride;   // transform-cache row width the block was sized for (a later scene may need wider rows)
    void* block;                // one allocation for all per-sample state
    size_t block_bytes;
    RenderCtx ctx;              // pointers filled in
    uint32_t* d_tile_ids;
    float* d_image;             // for the host-output entry point
    size_t image_floats;
    size_t image_bytes, tile_bytes;     // real sizes of d_image / d_tile_ids (pool blocks may be larger than asked)
    cudaEvent_t ev[4];
    // Per-bin grid sizing (rt_launch_shade / rt_launch_light_sample): the kernels record the largest count of every bin
    // (bin_hint), a batch's end copies the table to pinned memory, and a later batch -- whichever first finds the copy
    // done -- sizes each bin's grid from it instead of from the batch
    uint32_t* d_bin_hint;
    uint32_t* h_bin_hint;
    uint32_t bin_hint_used[RT_SBINS + RT_LBINS];
    bool bin_hint_valid, bin_hint_pending;
    cudaEvent_t bin_hint_ev;
    std::vector<cudaEvent_t>* trace_events;   // start/stop pairs around traversal kernels (RT_RENDER_TIME_TRACE)
    size_t trace_events_used;
};

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------

// Append to bin `bin` of a binned queue (items[bin * cap + k]).  Opportunistic warp
// aggregation: the lanes that happen to call together and target the same bin share
// one atomicAdd; callers may be divergent.
__device__ __forceinline__ void bq_push(uint32_t* items, uint32_t* counts, uint32_t cap, uint32_t bin, uint32_t value)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t active = __activemask();
    const uint32_t peers = __match_any_sync(active, bin);
    const uint32_t leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader)
        base = atomicAdd(counts + bin, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    items[(size_t)bin * cap + base + __popc(peers & ((1u << lane) - 1))] = value;
}

// Read side of a binned queue: entry j of the concatenation of the bins
template <int NB>
struct BinQ
{
    const uint32_t* items;
    const uint32_t* counts;
    uint32_t cap;
    __device__ __forceinline__ uint32_t total() const
    {
        uint32_t n = 0;
        #pragma unroll
        for (int b = 0; b < NB; ++b) n += counts[b];
        return n;
    }
    __device__ __forceinline__ uint32_t at(uint32_t j) const
    {
        uint32_t b = 0;
        #pragma unroll
        for (int k = 0; k < NB - 1; ++k)
        {
            uint32_t cnt = counts[k];
            if (b == (uint32_t)k && j >= cnt) { j -= cnt; b = k + 1; }
        }
        return items[(size_t)b * cap + j];
    }
};

__device__ __forceinline__ uint32_t dir_octant(V3 d)
{
    return ((d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u)) % RT_QBINS;
}

__device__ __forceinline__ V3 xyz(float4 v) { return mk(v.x, v.y, v.z); }
__device__ __forceinline__ Color3 rgb(float4 v) { return mkc(v.x, v.y, v.z); }

#define RT_GRID_STRIDE(j, n) \
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x, j##_end = ((n) + 31u) & ~31u; j < j##_end; j += gridDim.x * blockDim.x)

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(RT_BLOCK)
k_pixel_setup(const __grid_constant__ RenderCtx c)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0)
        for (int k = 0; k < CTL_WORDS; ++k) c.ctl[k] = 0;
    if (p >= c.num_pixels)
        return;
    uint32_t per_tile = c.tile * c.tile;
    uint32_t tile = c.tile_ids[p / per_tile];
    uint32_t off = p % per_tile;
    uint32_t x = (tile % c.tiles_x) * c.tile + off % c.tile;
    uint32_t y = (tile / c.tiles_x) * c.tile + off / c.tile;
    if (x >= c.width || y >= c.height)
    {
        c.pix_xy[p] = 0xffffffffu;
        return;
    }
    ChunkGrid g = chunk_grid(c.width, c.height);
    if (!chunk_covers(g, x, y))
    {
        // never rendered by the reference: stays Image's default black
        c.pix_xy[p] = 0xffffffffu;
        if (c.image)
        {
            float* out = pixel_out(c, p, x, y);
            out[0] = out[1] = out[2] = 0.0f;
        }
        return;
    }
    c.pix_xy[p] = (y << 16) | x;
    uint32_t perm[5 * RT_MAX_DEPTH + 3];
    pixel_permutations(g, x, y, c.depth, perm);
    uint32_t slots = 5 * c.depth + 3;
    for (uint32_t s = 0; s < slots; ++s)
        c.perms[(size_t)s * c.num_pixels + p] = perm[s];
}

// RenderThread::run inner loop body up to makeRay (RaytraceMain.cpp:120-142) and
// PerspectiveCamera::makeRay (:224-267)
__device__ __forceinline__ void camera_ray(const RenderCtx& c, uint32_t p, uint32_t psi, uint32_t x, uint32_t y,
                                           V3& origin, V3& dir, float& time)
{
    uint32_t D5 = 5 * c.depth;
    uint32_t perm_time = c.perms[(size_t)(D5 + 0) * c.num_pixels + p];
    uint32_t perm_lens = c.perms[(size_t)(D5 + 1) * c.num_pixels + p];
    uint32_t perm_sub = c.perms[(size_t)(D5 + 2) * c.num_pixels + p];
    float pu, pv;
    cmj2d(psi, c.ps, c.ps, perm_sub, pu, pv);
    float xu = ((float)x + pu) / (float)c.width;
    float yu = 1.0f - ((float)y + pv) / (float)c.height;
    // (the lens sample is a pure function of (psi, permutation): a pinhole camera never looks at it)
    float lens_u = 0.0f, lens_v = 0.0f;
    if (c.cam.lens_radius > 0)
        cmj2d(psi, c.ps, c.ps, perm_lens, lens_u, lens_v);
    float time_u = cmj1d(psi, c.spp, perm_time);

    float xs = (xu - 0.5f) * c.aspect + 0.5f;
    float ys = yu;
    const RtCamera& cam = c.cam;
    V3 fwd = mk(cam.forward[0], cam.forward[1], cam.forward[2]);
    V3 right = mk(cam.right[0], cam.right[1], cam.right[2]);
    V3 up = mk(cam.up[0], cam.up[1], cam.up[2]);
    origin = mk(cam.origin[0], cam.origin[1], cam.origin[2]);
    dir = fwd + right * ((xs - 0.5f) * cam.tan_fov) + up * ((ys - 0.5f) * cam.tan_fov);
    dir = normalized3(dir);
    time = cam.shutter_open + (cam.shutter_close - cam.shutter_open) * time_u;
    if (cam.lens_radius > 0)
    {
        float hshift, vshift;
        uniform_disk(lens_u, lens_v, hshift, vshift);
        hshift *= cam.lens_radius;
        vshift *= cam.lens_radius;
        V3 local_dir = normalized3(mk((xs - 0.5f) * cam.tan_fov, (ys - 0.5f) * cam.tan_fov, 1.0f));
        float focus_t = (cam.focal_distance - 0.0f) / local_dir.z;
        V3 focus = origin + dir * focus_t;
        origin = origin + (right * hshift + up * vshift);
        dir = normalized3(focus - origin);
    }
}

__global__ void __launch_bounds__(RT_BLOCK)
k_raygen(const __grid_constant__ RenderCtx c)
{
    RT_GRID_STRIDE(i, c.num_samples)
    {
        if (i < c.num_samples)
        {
            uint32_t p, psi;
            sample_split(c, i, p, psi);
            uint32_t xy = c.pix_xy[p];
            if (xy != 0xffffffffu)
            {
                V3 o, d;
                float time;
                camera_ray(c, p, psi, xy & 0xffffu, xy >> 16, o, d, time);
                if (c.xf_cache)
                {
                    // every animated transform of the scene at this sample's time, once, all lanes together
                    float4* row = c.xf_cache + (size_t)i * c.sc.anim_stride;
                    for (uint32_t k = 0; k < c.sc.num_anim; ++k)
                    {
                        const uint4 a = __ldg(c.sc.anim + k);
                        xform_cache_store(row, a.y, a.z, xform_eval(c.sc, a.x, time));
                    }
                }
                c.ray_od[2 * (size_t)i] = make_float4(o.x, o.y, o.z, time);
                c.ray_od[2 * (size_t)i + 1] = make_float4(d.x, d.y, d.z, 0.0f);
                c.thr[i] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u));
                c.res[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                bq_push(c.q_path[0], c.ctl + CTL_PATH_A, c.qcap, dir_octant(d), i);
            }
        }
    }
}

// Camera rays only (rt_generate_camera_rays)
__global__ void __launch_bounds__(RT_BLOCK)
k_camera_rays(const __grid_constant__ RenderCtx c, uint32_t psi, RtRay* out)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= c.num_pixels)
        return;
    uint32_t xy = c.pix_xy[p];
    if (xy == 0xffffffffu)
        return;
    uint32_t x = xy & 0xffffu, y = xy >> 16;
    V3 o, d;
    float time;
    camera_ray(c, p, psi, x, y, o, d, time);
    RtRay r;
    r.origin[0] = o.x; r.origin[1] = o.y; r.origin[2] = o.z;
    r.direction[0] = d.x; r.direction[1] = d.y; r.direction[2] = d.z;
    r.tmax = RT_RAY_TMAX;
    r.time = time;
    out[(size_t)y * c.width + x] = r;
}

__device__ __forceinline__ void flush_work_counters(const WorkCount& wc, uint64_t* totals)
{
    uint32_t v[RT_WORK_COUNTERS] = { wc.node_pops, wc.tri_tests, wc.shape_tests, wc.xform_evals, wc.xform_keyed, wc.xform_pairs };
    for (int k = 0; k < RT_WORK_COUNTERS; ++k)
    {
        uint32_t x = v[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0 && x)
            atomicAdd(reinterpret_cast<unsigned long long*>(totals + 2 + k), (unsigned long long)x);
    }
}

// Queue adaptors for the traversal kernels (rt_wave.cuh, rt_split.cuh)
struct PathIO
{
    BinQ<RT_QBINS> queue;
    const float4* ray_od;
    float4* hit01;
    // hits are handed to the shading stage binned by shape; misses end the path here
    uint32_t* shade_items;
    uint32_t* shade_counts;
    const float4* xf_cache;
    uint32_t xf_stride;
    __device__ __forceinline__ const float4* xf_row(uint32_t tag) const { return xf_cache ? xf_cache + (size_t)tag * xf_stride : nullptr; }
    __device__ __forceinline__ uint32_t count() const { return queue.total(); }
    __device__ __forceinline__ uint32_t tag_at(uint32_t j) const { return queue.at(j); }
    // the two 16-byte records of a ray and how they decode (used by the prefetching top-level pass)
    __device__ __forceinline__ const float4* rec_a(uint32_t tag) const { return ray_od + 2 * (size_t)tag; }
    __device__ __forceinline__ const float4* rec_b(uint32_t tag) const { return ray_od + 2 * (size_t)tag + 1; }
    __device__ __forceinline__ void decode(float4 a, float4 b, V3& o, V3& d, float& tmax, float& time) const
    {
        o = xyz(a); d = xyz(b); tmax = RT_RAY_TMAX; time = a.w;
    }
    __device__ __forceinline__ void load_tag(uint32_t tag, V3& o, V3& d, float& tmax, float& time) const
    {
        decode(ray_od[2 * (size_t)tag], ray_od[2 * (size_t)tag + 1], o, d, tmax, time);
    }
    __device__ __forceinline__ bool load(uint32_t j, V3& o, V3& d, float& tmax, float& time, uint32_t& tag) const
    {
        tag = queue.at(j);
        load_tag(tag, o, d, tmax, time);
        return true;
    }
    __device__ __forceinline__ void store(uint32_t tag, const WaveResult& r) const
    {
        hit01[2 * (size_t)tag] = make_float4(r.t, __int_as_float(r.shape), __int_as_float(r.tri_rec), 0.0f);
        if (r.shape >= 0)
            bq_push(shade_items, shade_counts, queue.cap, min((uint32_t)r.shape, (uint32_t)(RT_SBINS - 1)), tag);
    }
};

struct MisIO
{
    BinQ<RT_QBINS> queue;
    const float4* lrec;          // probe direction at [4 * tag + 2], origin and time at [4 * tag + 3]
    float4* mis_hit0;
    const float4* xf_cache;
    uint32_t xf_stride;
    __device__ __forceinline__ const float4* xf_row(uint32_t tag) const { return xf_cache ? xf_cache + (size_t)tag * xf_stride : nullptr; }
    __device__ __forceinline__ uint32_t count() const { return queue.total(); }
    __device__ __forceinline__ uint32_t tag_at(uint32_t j) const { return queue.at(j); }
    // the two 16-byte records of a ray and how they decode (used by the prefetching top-level pass)
    __device__ __forceinline__ const float4* rec_a(uint32_t tag) const { return lrec + 4 * (size_t)tag + 3; }
    __device__ __forceinline__ const float4* rec_b(uint32_t tag) const { return lrec + 4 * (size_t)tag + 2; }
    __device__ __forceinline__ void decode(float4 a, float4 b, V3& o, V3& d, float& tmax, float& time) const
    {
        o = xyz(a); d = xyz(b); tmax = RT_RAY_TMAX; time = a.w;
    }
    __device__ __forceinline__ void load_tag(uint32_t tag, V3& o, V3& d, float& tmax, float& time) const
    {
        decode(lrec[4 * (size_t)tag + 3], lrec[4 * (size_t)tag + 2], o, d, tmax, time);
    }
    __device__ __forceinline__ bool load(uint32_t j, V3& o, V3& d, float& tmax, float& time, uint32_t& tag) const
    {
        tag = queue.at(j);
        load_tag(tag, o, d, tmax, time);
        return true;
    }
    __device__ __forceinline__ void store(uint32_t tag, const WaveResult& r) const
    {
        mis_hit0[tag] = make_float4(r.t, __int_as_float(r.shape), __int_as_float(r.tri_rec), 0.0f);
    }
};

struct ShadowIO
{
    BinQ<RT_QBINS> queue;
    const float4* lrec;          // shadow direction and tMax at [4 * tag + 0], origin and time at [4 * tag + 1]
    float4* lterm;               // [2 * tag].w = light sample still valid
    const float4* xf_cache;
    uint32_t xf_stride;
    __device__ __forceinline__ const float4* xf_row(uint32_t tag) const { return xf_cache ? xf_cache + (size_t)tag * xf_stride : nullptr; }
    __device__ __forceinline__ uint32_t count() const { return queue.total(); }
    __device__ __forceinline__ uint32_t tag_at(uint32_t j) const { return queue.at(j); }
    // the two 16-byte records of a ray and how they decode (used by the prefetching top-level pass)
    __device__ __forceinline__ const float4* rec_a(uint32_t tag) const { return lrec + 4 * (size_t)tag + 1; }
    __device__ __forceinline__ const float4* rec_b(uint32_t tag) const { return lrec + 4 * (size_t)tag; }
    __device__ __forceinline__ void decode(float4 a, float4 b, V3& o, V3& d, float& tmax, float& time) const
    {
        o = xyz(a); d = xyz(b); tmax = b.w; time = a.w;
    }
    __device__ __forceinline__ void load_tag(uint32_t tag, V3& o, V3& d, float& tmax, float& time) const
    {
        decode(lrec[4 * (size_t)tag + 1], lrec[4 * (size_t)tag], o, d, tmax, time);
    }
    __device__ __forceinline__ bool load(uint32_t j, V3& o, V3& d, float& tmax, float& time, uint32_t& tag) const
    {
        tag = queue.at(j);
        load_tag(tag, o, d, tmax, time);
        return true;
    }
    // an occluded light sample is cancelled in place (the record k_resolve reads anyway)
    __device__ __forceinline__ void store(uint32_t tag, const WaveResult& r) const
    {
        if (r.any_hit)
            reinterpret_cast<float*>(lterm + 2 * (size_t)tag)[3] = 0.0f;
    }
};

__host__ __device__ __forceinline__ PathIO make_path_io(const RenderCtx& c, int cur)
{
    PathIO io = { { c.q_path[cur], c.ctl + CTL_PATH(cur), c.qcap }, c.ray_od, c.hit01, c.q_shade, c.ctl + CTL_SHADE,
                  c.xf_cache, c.sc.anim_stride };
    return io;
}
__host__ __device__ __forceinline__ MisIO make_mis_io(const RenderCtx& c)
{
    MisIO io = { { c.q_mis, c.ctl + CTL_MIS, c.qcap }, c.lrec, c.mis_hit0, c.xf_cache, c.sc.anim_stride };
    return io;
}
__host__ __device__ __forceinline__ ShadowIO make_shadow_io(const RenderCtx& c)
{
    ShadowIO io = { { c.q_shadow, c.ctl + CTL_SHADOW, c.qcap }, c.lrec, c.lterm, c.xf_cache, c.sc.anim_stride };
    return io;
}

// Path segments: closest hit (RaytraceMain.cpp:293-298)
template <int CAP, bool COUNT>
__global__ void __launch_bounds__(RT_BLOCK)
k_trace_paths(const __grid_constant__ RenderCtx c, int cur)
{
    PathIO io = make_path_io(c, cur);
    const uint32_t n = io.count();
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(c.totals + 0), (unsigned long long)n);
    WorkCount wc = RT_WORK_ZERO;
    trace_wave<CAP, false, COUNT>(c.sc, io, n, c.ctl + CTL_CUR_PATH, wc);
    if (COUNT)
        flush_work_counters(wc, c.totals);
}

// ---- split traversal kernels (rt_split.cuh) ---------------------------------
// minimum resident blocks per SM asked of the split traversal kernels (1 = let the compiler choose)
#ifndef RT_TOP_MINBLOCKS
#define RT_TOP_MINBLOCKS 1
#endif
#ifndef RT_MESH_MINBLOCKS
#define RT_MESH_MINBLOCKS 8        /* 64-register cap: with box_test_plain the pass would take 65 and lose a block per SM (C5: 1887 vs 1947) */
#endif
template <bool ANY, bool COUNT, bool FRESH, class IO>
__global__ void __launch_bounds__(RT_BLOCK, RT_TOP_MINBLOCKS)
k_split_top(const __grid_constant__ DScene sc, const IO io, const SplitBufs sb, const SplitPass ps, uint64_t* totals, int count_slot)
{
    split_zero(ps);
    if (FRESH && count_slot >= 0 && blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(totals + count_slot), (unsigned long long)io.count());
    WorkCount wc = RT_WORK_ZERO;
    trace_top<ANY, COUNT, FRESH>(sc, io, sb, ps, wc);
    if (COUNT)
        flush_work_counters(wc, totals);
}

// Fresh top-level pass over the tabulated walk (scenes whose top level fits DTopStep tables)
#ifndef RT_STATIC_MINBLOCKS
#define RT_STATIC_MINBLOCKS 8      /* 64 registers: 3208 -> 3523 Mrays/s (frame) in same-session A/B; 9, 10, 12 blocks are slower again */
#endif
template <bool ANY, bool COUNT, class IO>
__global__ void __launch_bounds__(RT_BLOCK, RT_STATIC_MINBLOCKS)
k_split_top_static(const __grid_constant__ DScene sc, const IO io, const SplitBufs sb, const SplitPass ps, uint64_t* totals, int count_slot)
{
    // per lane and per depth of the walk, the range a node's children inherit: (walk depth + 2) levels, sized by
    // the launch (the tables allow RT_WALK_MAX_DEPTH, the Stage 7 scene needs 4: the shared memory a block does not
    // take stays L1 cache, and this pass lives on L1 hits of the top-level nodes and shape records)
    extern __shared__ float lane_ranges[];
    float* lane_t0 = lane_ranges;
    float* lane_t1 = lane_ranges + (sc.top_walk_levels) * RT_BLOCK;
#if RT_STATIC_PREFETCH
    __shared__ float4 stage_rec[2 * 2 * RT_BLOCK];       // [stage][record a / b][thread]
#else
    float4* stage_rec = NULL;
#endif
    split_zero(ps);
    if (count_slot >= 0 && blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(totals + count_slot), (unsigned long long)io.count());
    WorkCount wc = RT_WORK_ZERO;
    trace_top_static<ANY, COUNT>(sc, io, sb, ps, wc, lane_t0, lane_t1, stage_rec);
    if (COUNT)
        flush_work_counters(wc, totals);
}

template <int CAP, bool ANY, bool COUNT, class IO>
__global__ void __launch_bounds__(RT_BLOCK, RT_MESH_MINBLOCKS)
k_split_mesh(const __grid_constant__ DScene sc, const IO io, const SplitBufs sb, const SplitPass ps, uint64_t* totals)
{
    split_zero(ps);
    WorkCount wc = RT_WORK_ZERO;
#if RT_MESH_PAIR
    trace_mesh_pair<CAP, ANY, COUNT>(sc, io, sb, ps, wc);
#else
    trace_mesh<CAP, ANY, COUNT>(sc, io, sb, ps, wc);
#endif
    if (COUNT)
        flush_work_counters(wc, totals);
}

// Which light does light-sample iteration `lsi` of path sample i use, and with which sample index
__device__ __forceinline__ void light_choice(const RenderCtx& c, uint32_t bounce, uint32_t lsi, uint32_t p, uint32_t psi,
                                             uint32_t& idx, uint32_t& light_index)
{
    if (c.sc.stage6)
    {
        // every light in turn, ls*ls samples each (S6 RaytraceMain.cpp:274-384)
        light_index = lsi / c.ls2;
        idx = psi * c.ls2 + lsi % c.ls2;
    }
    else
    {
        // random light (:358-364)
        uint32_t n1 = c.ps * c.ls * c.ps * c.ls;
        uint32_t perm_sel = c.perms[(size_t)(5 * bounce + 1) * c.num_pixels + p];
        idx = psi * c.nls + lsi;
        float liu = cmj1d(idx, n1, perm_sel);
        light_index = (uint32_t)(liu * (float)c.sc.num_lights);
        if (light_index >= c.sc.num_lights)
            light_index = c.sc.num_lights - 1;
    }
}

// pathTrace, one bounce, everything that does not need further rays, for hit path sample i.
// ST / BR >= 0: shape type and BRDF of the hit are known at compile time (the shade queue is binned by
// hit shape, and a shape has one type and one material), so a launch per bin runs a kernel that holds
// only that shape's normal code and that BRDF's sampling.
template <int ST, int BR>
__device__ __forceinline__ void shade_one(const RenderCtx& c, int cur, uint32_t bounce, uint32_t i)
{
    float4 h0 = c.hit01[2 * (size_t)i];
    int shape = __float_as_int(h0.y);
    if (shape >= 0)
    {
        float4 ro = c.ray_od[2 * (size_t)i], rd = c.ray_od[2 * (size_t)i + 1], th = c.thr[i], rs = c.res[i];
        // Intersection::m_normal / m_colorModifier of the winning hit, computed
        // here where all 32 lanes are busy rather than in the traversal loop
        float4 h1;
        {
            ClosestHit h;
            h.t = h0.x; h.shape = shape; h.tri_rec = __float_as_int(h0.z);
            TRS set_trs = xform_eval(c.sc, c.sc.set_xform, ro.w);
            LocalRay r0;
            r0.o = to_local_point(set_trs, xyz(ro));
            r0.d = to_local_vector(set_trs, xyz(rd));
            r0.inv = r0.d; r0.neg = 0;
            V3 nrm;
            float cmod;
            hit_shading_inputs(c.sc, r0, ro.w, h, nrm, cmod, xf_row(c, i), ST);
            h1 = make_float4(nrm.x, nrm.y, nrm.z, cmod);
            c.hit01[2 * (size_t)i + 1] = h1;
        }
        uint32_t state = __float_as_uint(th.w);
        uint32_t nb = state & 0xffu, nd = (state >> 8) & 0xffu;
        Color3 thr = rgb(th), result = rgb(rs);
        DShape sh = load_shape(c.sc, (uint32_t)shape);
        RtMaterial mat = c.sc.materials[sh.material];
        if (BR >= 0) mat.brdf = (uint32_t)BR;      // every hit of the bin has this BRDF: the others compile away

        // Emission only when seen directly or through mirrors (:303-306)
        // (Stage 6: only when seen directly, S6 RaytraceMain.cpp:250)
        if (nb == 0 || (nb == nd && !c.sc.stage6))
            result = result + thr * mkc(mat.emittance[0], mat.emittance[1], mat.emittance[2]);

        if (mat.brdf != RT_BRDF_NONE)      // an Emitter ends the path (:320-323)
        {
            V3 o = xyz(ro), d = xyz(rd);
            V3 position = o + h0.x * d;
            V3 normal = xyz(h1);
            V3 outgoing = -d;
            float cm = h1.w;
            bool dirac = mat.brdf == RT_BRDF_MIRROR;
            if (dirac)
                nd++;
            uint32_t p, psi;
            sample_split(c, i, p, psi);
            if (!dirac && c.nls > 0)
            {
                bq_push(c.q_lit, c.ctl + CTL_LIT, c.qcap, 0, i);
                // ... and regrouped by (light of light sample 0, BRDF kind) for k_light_sample; the
                // bins live in this bounce's consumed path queue
                uint32_t idx0, light0;
                light_choice(c, bounce, 0, p, psi, idx0, light0);
                bq_push(c.q_path[cur], c.ctl + CTL_LITB, c.qcap, (light0 & 3u) | (mat.brdf == RT_BRDF_GLOSSY ? 4u : 0u), i);
                c.lit_tr[2 * (size_t)i] = make_float4(thr.r, thr.g, thr.b, 0.0f);
                c.lit_tr[2 * (size_t)i + 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                c.pos_wo[2 * (size_t)i] = make_float4(position.x, position.y, position.z, ro.w);
                c.pos_wo[2 * (size_t)i + 1] = make_float4(outgoing.x, outgoing.y, outgoing.z, __uint_as_float(sh.material));
            }

            // Next leg of the path (:451-477)
            uint32_t perm = c.perms[(size_t)(5 * bounce + 0) * c.num_pixels + p];
            float u, v;
            cmj2d(psi, c.ps, c.ps, perm, u, v);
            V3 incoming;
            float pdf = 0.0f;
            float value = brdf_sample(mat.brdf, mat.exponent, incoming, outgoing, normal, u, v, pdf);
            if (pdf > 0.0f)
            {
                Color3 mc = mkc(mat.color[0], mat.color[1], mat.color[2]);
                Color3 f = mkc(cm, cm, cm) * mc * value * (fabsf(dot3(-incoming, normal)) / (pdf * 1.0f));
                thr = thr * f;
                nb++;
                V3 nd3 = -incoming;
                c.ray_od[2 * (size_t)i] = make_float4(position.x, position.y, position.z, ro.w);
                c.ray_od[2 * (size_t)i + 1] = make_float4(nd3.x, nd3.y, nd3.z, 0.0f);
                if (nb < c.depth)
                    bq_push(c.q_path[cur ^ 1], c.ctl + CTL_PATH(cur ^ 1), c.qcap, dir_octant(nd3), i);
            }
        }
        c.thr[i] = make_float4(thr.r, thr.g, thr.b, __uint_as_float(nb | (nd << 8)));
        c.res[i] = make_float4(result.r, result.g, result.b, 0.0f);
    }
}

__device__ __forceinline__ void shade_prologue(const RenderCtx& c)
{
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        for (int b = 0; b < RT_QBINS; ++b) { c.ctl[CTL_SHADOW + b] = 0; c.ctl[CTL_MIS + b] = 0; }
        c.ctl[CTL_CUR_SHADOW] = 0;
        c.ctl[CTL_CUR_MIS] = 0;
        c.ctl[CTL_CUR_PATH] = 0;
    }
}

__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE_MINBLOCKS)
k_shade(const __grid_constant__ RenderCtx c, int cur, uint32_t bounce)
{
    const BinQ<RT_SBINS> shade = { c.q_shade, c.ctl + CTL_SHADE, c.qcap };
    const uint32_t n = shade.total();
    shade_prologue(c);
    RT_GRID_STRIDE(j, n)
    {
        if (j < n)
            shade_one<-1, -1>(c, cur, bounce, shade.at(j));
    }
}

// The same for ONE bin (= one shape) of the shade queue; `first`: this launch also resets the counters
template <int ST, int BR>
__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE_MINBLOCKS)
k_shade_bin(const __grid_constant__ RenderCtx c, int cur, uint32_t bounce, uint32_t bin, int first)
{
    const uint32_t n = c.ctl[CTL_SHADE + bin];
    const uint32_t* items = c.q_shade + (size_t)bin * c.qcap;
    if (blockIdx.x == 0 && threadIdx.x == 0 && c.bin_hint != NULL)
        atomicMax(c.bin_hint + bin, n);
    if (first)
        shade_prologue(c);
    RT_GRID_STRIDE(j, n)
    {
        if (j < n)
            shade_one<ST, BR>(c, cur, bounce, items[j]);
    }
}

// Regroup the lit paths by (chosen light, BRDF kind) so that the lanes of a k_light_sample
// warp run the same light-sampling and BRDF code (light samples after the first; k_shade does it
// for the first).  The bins live in this bounce's path queue, consumed before k_shade ran.
__global__ void __launch_bounds__(RT_BLOCK)
k_light_select(const __grid_constant__ RenderCtx c, int cur, uint32_t bounce, uint32_t lsi)
{
    const uint32_t n = c.ctl[CTL_LIT];
    RT_GRID_STRIDE(j, n)
    {
        if (j < n)
        {
            const uint32_t i = c.q_lit[j];
            uint32_t idx, light_index;
            uint32_t p, psi;
            sample_split(c, i, p, psi);
            light_choice(c, bounce, lsi, p, psi, idx, light_index);
            const RtMaterial& mat = c.sc.materials[__float_as_uint(c.pos_wo[2 * (size_t)i + 1].w)];
            uint32_t bin = (light_index & 3u) | (mat.brdf == RT_BRDF_GLOSSY ? 4u : 0u);
            bq_push(c.q_path[cur], c.ctl + CTL_LITB, c.qcap, bin, i);
        }
    }
}

// One light sample of the direct-lighting loop (:336-422) for lit path sample i: produces at most one
// shadow ray and one BRDF-MIS probe.  LT / BR >= 0: the light's shape type / the surface's BRDF are
// known at compile time (the lit paths are regrouped by (light, BRDF kind) anyway, see k_light_select), so
// one launch per bin runs a kernel that holds only that light's sampling and that BRDF's code -- the
// all-in-one kernel is 64 KB of SASS and spent a quarter of its stall samples waiting for instructions
// (profiles/README.md, round 2).
template <int LT, int BR>
__device__ __forceinline__ void light_sample_one(const RenderCtx& c, uint32_t bounce, uint32_t lsi, uint32_t i)
{
    uint32_t p, psi;
    sample_split(c, i, p, psi);
    float4 pt = c.pos_wo[2 * (size_t)i], wm = c.pos_wo[2 * (size_t)i + 1], h1 = c.hit01[2 * (size_t)i + 1];
    V3 position = xyz(pt), outgoing = xyz(wm), normal = xyz(h1);
    float time = pt.w, cm = h1.w;
    RtMaterial mat = c.sc.materials[__float_as_uint(wm.w)];
    if (BR >= 0) mat.brdf = (uint32_t)BR;          // the bin holds this BRDF only: the other branches compile away
    Color3 mc = mkc(mat.color[0], mat.color[1], mat.color[2]);

    uint32_t n1 = c.ps * c.ls * c.ps * c.ls, n2 = c.ps * c.ls;
    const uint32_t* pp = c.perms + (size_t)(5 * bounce) * c.num_pixels + p;
    uint32_t perm_elem = pp[(size_t)2 * c.num_pixels];
    uint32_t perm_light = pp[(size_t)3 * c.num_pixels];
    uint32_t perm_brdf = pp[(size_t)4 * c.num_pixels];

    uint32_t idx, light_index;
    light_choice(c, bounce, lsi, p, psi, idx, light_index);
    if (c.sc.stage6)
    {
        // The Stage 6 reference draws these samples from one serial, data-dependent Rng,
        // which has no parallel equivalent; the counter-based stream stands in (same
        // strata, decorrelated per light), so Stage 6 images match statistically, not bitwise.
        uint32_t salt = (light_index + 1u) * 0x9e3779b9u;
        salt ^= salt >> 15; salt *= 0x85ebca6bu; salt ^= salt >> 13;
        perm_elem ^= salt; perm_light ^= salt * 0x9e3779b9u; perm_brdf ^= salt * 0x85ebca6bu;
    }
    uint32_t light_shape = c.sc.lights[light_index];
    DShape lsh = load_shape(c.sc, light_shape);
    if (LT >= 0) lsh.type = (uint32_t)LT;          // every light of the bin is of this kind
    RtMaterial lmat = c.sc.materials[lsh.material];
    Color3 emitted = mkc(lmat.emittance[0], lmat.emittance[1], lmat.emittance[2]);

    float lsu, lsv;
    cmj2d(idx, n2, n2, perm_light, lsu, lsv);
    float leu = cmj1d(idx, n1, perm_elem);
    V3 lpos, lnrm;
    float lpdf;
    light_sample(c.sc, lsh, position, time, lsu, lsv, leu, lpos, lnrm, lpdf, xf_row(c, i));

    float4 shl = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4 shd = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (lpdf > 0.0f)
    {
        V3 li = position - lpos;
        float dist;
        li = normalized3(li, &dist);
        float bpdf = 0.0f;
        float bres = brdf_evaluate(mat.brdf, mat.exponent, li, outgoing, normal, bpdf);
        if (bres > 0.0f && bpdf > 0.0f)
        {
            V3 sd = -li;
            float mis = power_heuristic(lpdf, bpdf);
            Color3 L = emitted * mkc(cm, cm, cm) * mc * bres * fabsf(dot3(sd, normal)) * mis / (lpdf * 1.0f);
            shd = make_float4(sd.x, sd.y, sd.z, dist - RT_RAY_TMIN);
            shl = make_float4(L.r, L.g, L.b, 1.0f);
            bq_push(c.q_shadow, c.ctl + CTL_SHADOW, c.qcap, dir_octant(sd), i);
        }
    }

    // BRDF sample towards (hopefully) the same light (:410-422)
    float bsu, bsv;
    cmj2d(idx, n2, n2, perm_brdf, bsu, bsv);
    V3 bi;
    float bpdf = 0.0f;
    float bres = brdf_sample(mat.brdf, mat.exponent, bi, outgoing, normal, bsu, bsv, bpdf);
    float4 md = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float4 mp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (bpdf > 0.0f && bres > 0.0f)
    {
        V3 pd = -bi;
        Color3 P = emitted * mkc(cm, cm, cm) * mc * bres * fabsf(dot3(pd, normal));
        md = make_float4(pd.x, pd.y, pd.z, bpdf);
        mp = make_float4(P.r, P.g, P.b, __uint_as_float(light_shape));
        bq_push(c.q_mis, c.ctl + CTL_MIS, c.qcap, dir_octant(pd), i);
    }
    // the whole 64-byte ray record and the whole 32-byte term record, every time: full sectors,
    // no read-modify-write
    float4* rec = c.lrec + 4 * (size_t)i;
    rec[0] = shd; rec[1] = pt; rec[2] = md; rec[3] = pt;
    float4* term = c.lterm + 2 * (size_t)i;
    term[0] = shl; term[1] = mp;
}

__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE_MINBLOCKS)
k_light_sample(const __grid_constant__ RenderCtx c, int cur, uint32_t bounce, uint32_t lsi)
{
    const BinQ<RT_LBINS> lit = { c.q_path[cur], c.ctl + CTL_LITB, c.qcap };
    const uint32_t n = lit.total();
    RT_GRID_STRIDE(j, n)
    {
        if (j < n)
            light_sample_one<-1, -1>(c, bounce, lsi, lit.at(j));
    }
}

// The same for ONE bin of the regrouped queue
template <int LT, int BR>
__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE_MINBLOCKS)
k_light_sample_bin(const __grid_constant__ RenderCtx c, int cur, uint32_t bounce, uint32_t lsi, uint32_t bin)
{
    const uint32_t n = c.ctl[CTL_LITB + bin];
    const uint32_t* items = c.q_path[cur] + (size_t)bin * c.qcap;
    if (blockIdx.x == 0 && threadIdx.x == 0 && c.bin_hint != NULL)
        atomicMax(c.bin_hint + RT_SBINS + bin, n);
    RT_GRID_STRIDE(j, n)
    {
        if (j < n)
            light_sample_one<LT, BR>(c, bounce, lsi, items[j]);
    }
}

// Shadow rays: any hit (RaytraceMain.cpp:394-395)
template <int CAP, bool COUNT>
__global__ void __launch_bounds__(RT_BLOCK)
k_trace_shadow(const __grid_constant__ RenderCtx c)
{
    ShadowIO io = make_shadow_io(c);
    const uint32_t n = io.count();
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(c.totals + 1), (unsigned long long)n);
    WorkCount wc = RT_WORK_ZERO;
    trace_wave<CAP, true, COUNT>(c.sc, io, n, c.ctl + CTL_CUR_SHADOW, wc);
    if (COUNT)
        flush_work_counters(wc, c.totals);
}

// BSDF-sampled MIS probes: closest hit (RaytraceMain.cpp:422-423)
template <int CAP, bool COUNT>
__global__ void __launch_bounds__(RT_BLOCK)
k_trace_mis(const __grid_constant__ RenderCtx c)
{
    MisIO io = make_mis_io(c);
    const uint32_t n = io.count();
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(c.totals + 0), (unsigned long long)n);
    WorkCount wc = RT_WORK_ZERO;
    trace_wave<CAP, false, COUNT>(c.sc, io, n, c.ctl + CTL_CUR_MIS, wc);
    if (COUNT)
        flush_work_counters(wc, c.totals);
}

// Combine the two MIS samples of light sample `lsi` (:396-439) and, after the last
// one, fold the bounce's direct lighting into the path (:443-447)
__global__ void __launch_bounds__(RT_BLOCK, RT_SHADE_MINBLOCKS)
k_resolve(const __grid_constant__ RenderCtx c, uint32_t lsi)
{
    const uint32_t n = c.ctl[CTL_LIT];
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        for (int b = 0; b < RT_QBINS; ++b) { c.ctl[CTL_SHADOW + b] = 0; c.ctl[CTL_MIS + b] = 0; }   // refilled by the next light sample
        for (int b = 0; b < RT_LBINS; ++b) c.ctl[CTL_LITB + b] = 0;
        c.ctl[CTL_CUR_SHADOW] = 0;
        c.ctl[CTL_CUR_MIS] = 0;
    }
    RT_GRID_STRIDE(j, n)
    {
        if (j >= n)
            continue;
        uint32_t i = c.q_lit[j];
        Color3 lr = rgb(c.lit_tr[2 * (size_t)i + 1]);
        const float4* rec = c.lrec + 4 * (size_t)i;
        const float4* term = c.lterm + 2 * (size_t)i;
        float4 shl = term[0];
        if (shl.w != 0.0f)
            lr = lr + rgb(shl);
        float4 md = rec[2];
        if (md.w > 0.0f)
        {
            float4 mp = term[1], mh0 = c.mis_hit0[i];
            uint32_t light_shape = __float_as_uint(mp.w);
            if (__float_as_int(mh0.y) == (int)light_shape)
            {
                float4 pt = rec[3];
                DShape lsh = load_shape(c.sc, light_shape);
                // normal of the probe's hit (only needed once the probe found the light)
                ClosestHit h;
                h.t = mh0.x; h.shape = (int32_t)light_shape; h.tri_rec = __float_as_int(mh0.z);
                TRS set_trs = xform_eval(c.sc, c.sc.set_xform, pt.w);
                LocalRay r0;
                r0.o = to_local_point(set_trs, xyz(pt));
                r0.d = to_local_vector(set_trs, xyz(md));
                r0.inv = r0.d; r0.neg = 0;
                V3 hn;
                float hcm;
                hit_shading_inputs(c.sc, r0, pt.w, h, hn, hcm, xf_row(c, i));
                float lpdf = light_intersect_pdf(c.sc, lsh, xyz(pt), xyz(md), pt.w, mh0.x, hn, xf_row(c, i));
                if (lpdf > 0.0f)
                {
                    float mis = power_heuristic(md.w, lpdf);
                    lr = lr + rgb(mp) * mis / (md.w * 1.0f);
                }
            }
        }
        if (c.sc.stage6)
        {
            // after a light's last sample: average and add (S6 RaytraceMain.cpp:375-383)
            if ((lsi + 1) % c.ls2 == 0)
            {
                lr = lr / (float)c.ls2;
                float4 rs = c.res[i];
                Color3 result = rgb(rs) + rgb(c.lit_tr[2 * (size_t)i]) * lr;
                c.res[i] = make_float4(result.r, result.g, result.b, 0.0f);
                lr = mkc(0.0f, 0.0f, 0.0f);
            }
            c.lit_tr[2 * (size_t)i + 1] = make_float4(lr.r, lr.g, lr.b, 0.0f);
        }
        else if (lsi + 1 == c.nls)
        {
            float weight = (float)c.sc.num_lights / (float)c.nls;
            lr = lr * weight;
            float4 rs = c.res[i];
            Color3 result = rgb(rs) + rgb(c.lit_tr[2 * (size_t)i]) * lr;
            c.res[i] = make_float4(result.r, result.g, result.b, 0.0f);
        }
        else
        {
            c.lit_tr[2 * (size_t)i + 1] = make_float4(lr.r, lr.g, lr.b, 0.0f);
        }
    }
}

// Box filter: samples summed in ascending order, then divided (:145-156)
__global__ void __launch_bounds__(RT_BLOCK)
k_accumulate(const __grid_constant__ RenderCtx c)
{
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= c.num_pixels)
        return;
    uint32_t xy = c.pix_xy[p];
    if (xy == 0xffffffffu)
        return;
    Color3 sum = mkc(0.0f, 0.0f, 0.0f);
    const float4* r = c.res + (size_t)p * c.spp;
    for (uint32_t s = 0; s < c.spp; ++s)
        sum = sum + rgb(r[s]);
    sum = sum / (float)c.spp;
    uint32_t x = xy & 0xffffu, y = xy >> 16;
    float* out = pixel_out(c, p, x, y);
    out[0] = sum.r;
    out[1] = sum.g;
    out[2] = sum.b;
}

// displayImage (MainWindow.cpp:37-91): negative -> green, pow(c * 2^exposure, 1/gamma),
// NaN -> blue, clamp, truncate.  Output bytes B, G, R, A.
__global__ void k_tonemap(const float* __restrict__ rgb_in, size_t n, float exposure, float gamma_exp, uint8_t* __restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float r = rgb_in[3 * i], g = rgb_in[3 * i + 1], b = rgb_in[3 * i + 2];
    if (r < 0.0f || g < 0.0f || b < 0.0f)
    {
        r = 0.0f; g = 1.0f; b = 0.0f;
    }
    else
    {
        r = ref_powf(r * exposure, gamma_exp);
        g = ref_powf(g * exposure, gamma_exp);
        b = ref_powf(b * exposure, gamma_exp);
        if (r != r || g != g || b != b)
        {
            r = 0.0f; g = 0.0f; b = 1.0f;
        }
    }
    r = std_max(0.0f, std_min(1.0f, r));
    g = std_max(0.0f, std_min(1.0f, g));
    b = std_max(0.0f, std_min(1.0f, b));
    out[4 * i + 3] = 0xff;
    out[4 * i + 2] = (uint8_t)(r * 255.0f);
    out[4 * i + 1] = (uint8_t)(g * 255.0f);
    out[4 * i + 0] = (uint8_t)(b * 255.0f);
}

// Stage 1 (Rayito_Stage1/main.cpp:93-135, rayito.h:474-509): rays through pixel
// corners, closest one-sided plane in list order, kRayTMin = 1e-5, 8-bit truncation.
// Stage 1's Vector::normalize divides unconditionally (rayito.h:194).
__global__ void k_stage1(const RtStage1Plane* __restrict__ planes, uint32_t num_planes, const RtCamera cam,
                         uint32_t width, uint32_t height, uint8_t* __restrict__ out, float* __restrict__ out_f)
{
    uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t y = blockIdx.y;
    if (x >= width || y >= height)
        return;
    float yu = 1.0f - ((float)y / (float)(height - 1));
    float xu = (float)x / (float)(width - 1);
    V3 fwd = mk(cam.forward[0], cam.forward[1], cam.forward[2]);
    V3 right = mk(cam.right[0], cam.right[1], cam.right[2]);
    V3 up = mk(cam.up[0], cam.up[1], cam.up[2]);
    V3 o = mk(cam.origin[0], cam.origin[1], cam.origin[2]);
    V3 d = fwd + right * ((xu - 0.5f) * cam.tan_fov) + up * ((yu - 0.5f) * cam.tan_fov);
    float len = length3(d);
    d = mk(d.x / len, d.y / len, d.z / len);
    float best = RT_RAY_TMAX;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    for (uint32_t k = 0; k < num_planes; ++k)
    {
        V3 n = mk(planes[k].normal[0], planes[k].normal[1], planes[k].normal[2]);
        V3 p = mk(planes[k].position[0], planes[k].position[1], planes[k].position[2]);
        float n_dot_d = dot3(n, d);
        if (n_dot_d >= 0.0f)
            continue;
        float t = (dot3(p, n) - dot3(o, n)) / dot3(d, n);
        if (t >= best || t < 0.00001f)
            continue;
        best = t;
        r = planes[k].color[0]; g = planes[k].color[1]; b = planes[k].color[2];
    }
    if (out_f)
    {
        // pixelColor before clamp(): what the WRITE_PFM build streams out (main.cpp:122-123)
        float* pf = out_f + ((size_t)y * width + x) * 3;
        pf[0] = r; pf[1] = g; pf[2] = b;
    }
    r = std_max(0.0f, std_min(1.0f, r));
    g = std_max(0.0f, std_min(1.0f, g));
    b = std_max(0.0f, std_min(1.0f, b));
    uint8_t* px = out + ((size_t)y * width + x) * 3;
    px[0] = (uint8_t)(r * 255.0f);
    px[1] = (uint8_t)(g * 255.0f);
    px[2] = (uint8_t)(b * 255.0f);
}

inline int rt_stage1_impl(int device, const RtStage1Plane* planes, uint32_t num_planes, const RtCamera* cam,
                          uint32_t width, uint32_t height, uint8_t* rgb8, float* rgb = NULL)
{
    if (cam == NULL || rgb8 == NULL || (num_planes && planes == NULL) || width < 2 || height < 2)
        return rt_fail(RT_ERR_ARG, "bad argument (Stage 1 needs width, height >= 2)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return rt_fail(RT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    RT_CUDA(cudaSetDevice(device));
    RtStage1Plane* d_planes = NULL;
    uint8_t* d_out = NULL;
    float* d_out_f = NULL;
    size_t bytes = (size_t)width * height * 3;
    cudaError_t e = cudaMalloc((void**)&d_planes, sizeof(RtStage1Plane) * (num_planes ? num_planes : 1));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, bytes);
    if (e == cudaSuccess && rgb != NULL) e = cudaMalloc((void**)&d_out_f, bytes * sizeof(float));
    if (e == cudaSuccess && num_planes)
        e = cudaMemcpy(d_planes, planes, sizeof(RtStage1Plane) * num_planes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        dim3 grid((width + 127) / 128, height);
        k_stage1<<<grid, 128>>>(d_planes, num_planes, *cam, width, height, d_out, d_out_f);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(rgb8, d_out, bytes, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rgb != NULL) e = cudaMemcpy(rgb, d_out_f, bytes * sizeof(float), cudaMemcpyDeviceToHost);
    if (d_planes) cudaFree(d_planes);
    if (d_out) cudaFree(d_out);
    if (d_out_f) cudaFree(d_out_f);
    if (e != cudaSuccess) return rt_cuda_fail(e, "stage 1 render");
    return RT_OK;
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------

// Give a RenderBuffers' device memory and events back for good
inline void rt_render_free(int device, RenderBuffers* rb)
{
    if (rb == NULL)
        return;
    rt_detail::pool_free(device, rb->block, rb->block_bytes);
    rt_detail::pool_free(device, rb->d_tile_ids, rb->tile_bytes);
    rt_detail::pool_free(device, rb->d_image, rb->image_bytes);
    if (rb->d_bin_hint) cudaFree(rb->d_bin_hint);
    if (rb->h_bin_hint) cudaFreeHost(rb->h_bin_hint);
    if (rb->bin_hint_ev) cudaEventDestroy(rb->bin_hint_ev);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(rb->ev[i]);
    if (rb->trace_events)
    {
        for (size_t i = 0; i < rb->trace_events->size(); ++i) cudaEventDestroy((*rb->trace_events)[i]);
        delete rb->trace_events;
    }
    delete rb;
}

// Rayito::raytrace() builds a fresh device scene per call, so the wavefront state of the previous call -- its
// 88 GB block carved into arrays, its events -- used to be taken apart at rt_scene_destroy and put together again by
// the next call's first render: a cudaMemGetInfo, a pool search and a dozen event creations per call, 5 to 80 ms on
// the boxes of round 2 (RAYITO_B200_TIMING=1, "reserve").  One RenderBuffers per device is parked WHOLE instead and
// adopted by the next scene that renders on that device.  rt_release_cached_memory() frees it.
namespace rt_detail
{
inline std::mutex& parked_render_lock() { static std::mutex m; return m; }
inline std::vector<std::pair<int, RenderBuffers*> >& parked_render() { static std::vector<std::pair<int, RenderBuffers*> > v; return v; }

inline RenderBuffers* take_parked_render(int device)
{
    std::lock_guard<std::mutex> guard(parked_render_lock());
    std::vector<std::pair<int, RenderBuffers*> >& v = parked_render();
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i].first == device)
        {
            RenderBuffers* rb = v[i].second;
            v.erase(v.begin() + i);
            return rb;
        }
    return NULL;
}

inline void release_parked_render()
{
    std::vector<std::pair<int, RenderBuffers*> > all;
    {
        std::lock_guard<std::mutex> guard(parked_render_lock());
        all.swap(parked_render());
    }
    int current = 0;
    cudaGetDevice(&current);
    for (size_t i = 0; i < all.size(); ++i)
    {
        cudaSetDevice(all[i].first);
        rt_render_free(all[i].first, all[i].second);
    }
    cudaSetDevice(current);
}
} // namespace rt_detail

inline void rt_render_release(RtScene* s)
{
    RenderBuffers* rb = s->render;
    if (rb == NULL)
        return;
    s->render = NULL;
    RenderBuffers* old = NULL;
    {
        std::lock_guard<std::mutex> guard(rt_detail::parked_render_lock());
        std::vector<std::pair<int, RenderBuffers*> >& v = rt_detail::parked_render();
        size_t at = v.size();
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].first == s->device) at = i;
        if (at == v.size())
            v.push_back(std::make_pair(s->device, rb));
        else if (v[at].second->block_bytes <= rb->block_bytes)
        {
            old = v[at].second;         // keep the larger working set
            v[at].second = rb;
        }
        else
            old = rb;
    }
    rt_render_free(s->device, old);
}

namespace rt_detail
{

struct Carver
{
    char* base;
    size_t off;
    template <typename T> T* take(size_t count)
    {
        off = (off + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : NULL;
        off += count * sizeof(T);
        return p;
    }
};

inline size_t carve(RenderCtx& c, char* base, size_t samples, size_t pixels, uint32_t slots, uint32_t xf_stride)
{
    Carver k = { base, 0 };
    c.pix_xy = k.take<uint32_t>(pixels);
    c.perms = k.take<uint32_t>(pixels * slots);
    c.ray_od = k.take<float4>(samples * 2);
    c.hit01 = k.take<float4>(samples * 2);
    c.thr = k.take<float4>(samples);
    c.res = k.take<float4>(samples);
    c.pos_wo = k.take<float4>(samples * 2);
    c.lit_tr = k.take<float4>(samples * 2);
    c.lrec = k.take<float4>(samples * 4);
    c.lterm = k.take<float4>(samples * 2);
    c.mis_hit0 = k.take<float4>(samples);
    c.mis_hit1 = k.take<float4>(samples);
    c.xf_cache = xf_stride ? k.take<float4>(samples * xf_stride) : NULL;
    c.qcap = (uint32_t)samples;
    c.q_path[0] = k.take<uint32_t>(samples * RT_QBINS);
    c.q_path[1] = k.take<uint32_t>(samples * RT_QBINS);
    c.q_shade = k.take<uint32_t>(samples * RT_SBINS);
    c.q_lit = k.take<uint32_t>(samples);
    c.q_shadow = k.take<uint32_t>(samples * RT_QBINS);
    c.q_mis = k.take<uint32_t>(samples * RT_QBINS);
    c.q_meshq[0] = k.take<uint32_t>(samples);
    c.q_meshq[1] = k.take<uint32_t>(samples);
    c.q_resume[0] = k.take<uint32_t>(samples);
    c.q_resume[1] = k.take<uint32_t>(samples);
    c.split.rec = k.take<float4>(samples * RT_SPLIT_REC);
    c.split.stack = k.take<float4>(samples * (RT_SPLIT_TOPCAP - 5));
    c.ctl = k.take<uint32_t>(CTL_WORDS);
    c.totals = k.take<uint64_t>(8);
    return k.off + 256;
}

// Which rank renders tile (tx, ty): (tx + m * ty) mod world, a lattice.  Every rank gets tiles from all image
// regions (sky rows are ~1 ray/sample, mesh pixels 6+), and m is the multiplier whose lattice keeps a rank's own
// tiles FARTHEST APART (largest shortest vector among the tiles of one rank): m = 1, the diagonal interleave of
// round 1, puts a rank's tiles corner to corner along diagonals, so a rank's share of the expensive pixels followed
// whatever diagonal structure the image has (8 ranks on config C4: per-rank frame times spread by 3.4 %, session Q/S
// in profiles/README.md); for world = 8 the choice is m = 3 (shortest vector (2,2) instead of (1,-1)).
// RAYITO_B200_TILE_DEAL=diagonal keeps m = 1 (A/B runs).
inline uint32_t tile_lattice_multiplier(uint32_t world)
{
    static std::mutex lock;
    static std::vector<uint32_t> known(1, 0u);      // known[w]: multiplier of world size w (0 = not computed yet)
    const char* env = std::getenv("RAYITO_B200_TILE_DEAL");
    if (world < 3 || (env != NULL && env[0] == 'd'))
        return 1u;
    std::lock_guard<std::mutex> guard(lock);
    if (known.size() <= world)
        known.resize((size_t)world + 1, 0u);
    if (known[world] == 0)
    {
        const int w = (int)world;
        const int reach = w < 64 ? w : 64;
        long best_len = -1;
        uint32_t best_m = 1;
        for (int m = 1; m < w; ++m)
        {
            long shortest = (long)w * w;            // (w, 0) is always in the lattice
            for (int dy = 0; dy <= reach; ++dy)
                for (int dx = -reach; dx <= reach; ++dx)
                {
                    if ((dx == 0 && dy == 0) || ((dx + m * dy) % w + w) % w != 0)
                        continue;
                    long len = (long)dx * dx + (long)dy * dy;
                    if (len < shortest) shortest = len;
                }
            if (shortest > best_len)
            {
                best_len = shortest;
                best_m = (uint32_t)m;
            }
        }
        known[world] = best_m;
    }
    return known[world];
}

inline void rank_tiles(uint32_t tiles_x, uint32_t tiles_y, uint32_t rank, uint32_t world, std::vector<uint32_t>& out)
{
    out.clear();
    const uint32_t m = tile_lattice_multiplier(world);
    for (uint32_t ty = 0; ty < tiles_y; ++ty)
        for (uint32_t tx = 0; tx < tiles_x; ++tx)
            if ((tx + m * ty) % world == rank)
                out.push_back(ty * tiles_x + tx);
}

} // namespace rt_detail

struct RenderPlan
{
    uint32_t tile, tiles_x, tiles_y;
    uint32_t spp, slots;
    uint32_t tiles_per_batch;
    std::vector<uint32_t> tiles;
};

inline int rt_plan(const RtScene* s, const RtRenderParams* prm, RenderPlan& plan)
{
    if (prm->width == 0 || prm->height == 0 || prm->width > 65535 || prm->height > 65535)
        return rt_fail(RT_ERR_ARG, "image size must be 1..65535");
    if (prm->pixel_samples_hint == 0 || prm->pixel_samples_hint > 1024)
        return rt_fail(RT_ERR_ARG, "pixel_samples_hint must be 1..1024");
    if (prm->max_ray_depth > RT_MAX_DEPTH)
        return rt_fail(RT_ERR_ARG, "max_ray_depth above the supported 16");
    if (prm->world == 0 || prm->rank >= prm->world)
        return rt_fail(RT_ERR_ARG, "rank/world invalid");
    uint64_t n1 = (uint64_t)prm->pixel_samples_hint * prm->light_samples_hint;
    if (n1 * n1 > 0xffffffffull)
        return rt_fail(RT_ERR_ARG, "(pixel_samples*light_samples)^2 does not fit the sampler's 32-bit index");
    (void)s;
    plan.tile = prm->tile_size ? prm->tile_size : RT_DEFAULT_TILE;
    plan.tiles_x = (prm->width + plan.tile - 1) / plan.tile;
    plan.tiles_y = (prm->height + plan.tile - 1) / plan.tile;
    plan.spp = prm->pixel_samples_hint * prm->pixel_samples_hint;
    plan.slots = 5 * prm->max_ray_depth + 3;
    uint64_t batch = prm->max_batch_samples ? prm->max_batch_samples : RT_DEFAULT_BATCH;
    uint64_t per_tile = (uint64_t)plan.tile * plan.tile * plan.spp;
    if (per_tile >= (1ull << 31))
        return rt_fail(RT_ERR_ARG, "tile_size^2 * spp too large; use a smaller tile");
    plan.tiles_per_batch = (uint32_t)std::max<uint64_t>(1, batch / per_tile);
    rt_detail::rank_tiles(plan.tiles_x, plan.tiles_y, prm->rank, prm->world, plan.tiles);
    if (plan.tiles_per_batch > plan.tiles.size())
        plan.tiles_per_batch = (uint32_t)std::max<size_t>(1, plan.tiles.size());
    return RT_OK;
}

inline int rt_render_reserve(RtScene* s, RenderPlan& plan)
{
    size_t pixels = (size_t)plan.tiles_per_batch * plan.tile * plan.tile;
    size_t samples = pixels * plan.spp;
    RenderBuffers* rb = s->render;
    if (rb == NULL)
    {
        rb = rt_detail::take_parked_render(s->device);       // the previous scene's, whole (see rt_render_release)
        if (rb == NULL)
        {
            rb = new RenderBuffers();
            std::memset(rb, 0, sizeof(*rb));
            for (int i = 0; i < 4; ++i) cudaEventCreate(&rb->ev[i]);
            const size_t hint_bytes = (RT_SBINS + RT_LBINS) * sizeof(uint32_t);
            if (cudaMalloc((void**)&rb->d_bin_hint, hint_bytes) == cudaSuccess && cudaMemset(rb->d_bin_hint, 0, hint_bytes) == cudaSuccess &&
                cudaHostAlloc((void**)&rb->h_bin_hint, hint_bytes, cudaHostAllocDefault) == cudaSuccess &&
                cudaEventCreateWithFlags(&rb->bin_hint_ev, cudaEventDisableTiming) == cudaSuccess)
                std::memset(rb->h_bin_hint, 0, hint_bytes);
            else
            {
                cudaGetLastError();         // no hints: every bin's grid stays sized by the batch
                if (rb->d_bin_hint) cudaFree(rb->d_bin_hint);
                if (rb->h_bin_hint) cudaFreeHost(rb->h_bin_hint);
                rb->d_bin_hint = rb->h_bin_hint = NULL;
            }
        }
        s->render = rb;
    }
    if (samples > rb->cap_samples || pixels > rb->cap_pixels || plan.slots > rb->cap_perm_slots ||
        s->d.anim_stride > rb->cap_anim_stride)
    {
        rt_detail::pool_free(s->device, rb->block, rb->block_bytes);
        rb->block = NULL;
        rb->cap_samples = rb->cap_pixels = 0;
        RenderCtx probe;
        size_t bytes = rt_detail::carve(probe, NULL, samples, pixels, plan.slots, s->d.anim_stride);
        size_t got_bytes = 0;
        // leave room for the rest of the process: shrink the batch until it fits in
        // at most 60 % of the free device memory, and on allocation failure
        for (;;)
        {
            // parked blocks count as free: the pool hands them back or releases them on demand
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            free_b += rt_detail::pool_parked_bytes(s->device);
            cudaError_t e = bytes <= free_b / 10 * 6 ? rt_detail::pool_alloc(s->device, &rb->block, bytes, &got_bytes)
                                                     : cudaErrorMemoryAllocation;
            if (e == cudaSuccess)
                break;
            cudaGetLastError();
            rb->block = NULL;
            if (plan.tiles_per_batch <= 1)
                return rt_cuda_fail(cudaErrorMemoryAllocation, "wavefront state allocation");
            plan.tiles_per_batch = (plan.tiles_per_batch + 1) / 2;
            pixels = (size_t)plan.tiles_per_batch * plan.tile * plan.tile;
            samples = pixels * plan.spp;
            bytes = rt_detail::carve(probe, NULL, samples, pixels, plan.slots, s->d.anim_stride);
        }
        rb->block_bytes = got_bytes;
        rb->cap_samples = samples;
        rb->cap_pixels = pixels;
        rb->cap_perm_slots = plan.slots;
        rb->cap_anim_stride = s->d.anim_stride;
    }
    rt_detail::carve(rb->ctx, static_cast<char*>(rb->block), rb->cap_samples, rb->cap_pixels, rb->cap_perm_slots, s->d.anim_stride);
    if (plan.tiles.size() > rb->cap_tiles)
    {
        rt_detail::pool_free(s->device, rb->d_tile_ids, rb->tile_bytes);
        rb->d_tile_ids = NULL;
        rb->cap_tiles = 0;
        RT_CUDA(rt_detail::pool_alloc(s->device, (void**)&rb->d_tile_ids, plan.tiles.size() * sizeof(uint32_t), &rb->tile_bytes));
        rb->cap_tiles = plan.tiles.size();
    }
    return RT_OK;
}

// Event pair helpers: traversal kernels are timed on the launching stream
inline void rt_trace_mark(RenderBuffers* rb, bool timed, cudaStream_t st)
{
    if (!timed)
        return;
    if (rb->trace_events == NULL)
        rb->trace_events = new std::vector<cudaEvent_t>();
    if (rb->trace_events_used == rb->trace_events->size())
    {
        cudaEvent_t e;
        cudaEventCreate(&e);
        rb->trace_events->push_back(e);
    }
    cudaEventRecord((*rb->trace_events)[rb->trace_events_used++], st);
}

// One traversal stage in split mode: fresh top-level pass, then (mesh pass, resume
// pass) once per mesh shape of the scene.  Counts live on the device; passes with an
// empty queue exit at once.
template <bool ANY, bool COUNT, class IO>
static void rt_launch_split_stage(RtScene* s, const RenderCtx& c, const IO& io,
                                  uint32_t* fresh_cursor, int count_slot, unsigned grid, unsigned grid_mesh, cudaStream_t st, uint64_t& launches)
{
    cudaMemsetAsync(c.ctl + CTL_MESH_N, 0, 8 * sizeof(uint32_t), st);
    SplitPass p;
    p.in_queue = NULL;
    p.in_count = NULL;           // the fresh pass counts its IO's binned queue
    p.cursor = fresh_cursor;
    p.out_queue = c.q_meshq[0];
    p.out_count = c.ctl + CTL_MESH_N + 0;
    p.zero[0] = p.zero[1] = p.zero[2] = p.zero[3] = NULL;
    // The tabulated walk pays most for shadow rays (incoherent origins: the per-lane pass ran at 12
    // lanes and twice the instructions, 2.4 -> 1.7 ms per launch).  Closest-hit rays visit so few
    // leaves each that a warp-wide walk over the union of their leaves executes about as many
    // instructions as the per-lane pass; same-session A/B at 1920x1080: per-lane 3282, tabulated for
    // shadow rays only 3360, tabulated for all 3389 Mrays/s (profiles/README.md, v7).
#ifndef RT_STATIC_TOP_CLOSEST
#define RT_STATIC_TOP_CLOSEST 1
#endif
    if (c.sc.top_walk_steps > 0 && !s->dynamic_top && (ANY || RT_STATIC_TOP_CLOSEST))
        k_split_top_static<ANY, COUNT, IO><<<grid, RT_BLOCK, 2 * c.sc.top_walk_levels * RT_BLOCK * sizeof(float), st>>>(
            c.sc, io, c.split, p, c.totals, count_slot);
    else
        k_split_top<ANY, COUNT, true, IO><<<grid, RT_BLOCK, 0, st>>>(c.sc, io, c.split, p, c.totals, count_slot);
    launches += 1;
    const bool deep = s->mesh_stack_need > 32;
    for (uint32_t k = 0; k < s->num_mesh_shapes; ++k)
    {
        const int a = (int)(k & 1), b = (int)((k + 1) & 1);
        SplitPass m;
        m.in_queue = c.q_meshq[a];
        m.in_count = c.ctl + CTL_MESH_N + a;
        m.cursor = c.ctl + CTL_MESH_CUR + a;
        m.out_queue = c.q_resume[a];
        m.out_count = c.ctl + CTL_RES_N + a;
        m.zero[0] = c.ctl + CTL_MESH_N + b;
        m.zero[1] = c.ctl + CTL_MESH_CUR + b;
        m.zero[2] = m.zero[3] = NULL;
        if (deep) k_split_mesh<64, ANY, COUNT, IO><<<grid_mesh, RT_BLOCK, 0, st>>>(c.sc, io, c.split, m, c.totals);
        else      k_split_mesh<32, ANY, COUNT, IO><<<grid_mesh, RT_BLOCK, 0, st>>>(c.sc, io, c.split, m, c.totals);
        SplitPass r;
        r.in_queue = c.q_resume[a];
        r.in_count = c.ctl + CTL_RES_N + a;
        r.cursor = c.ctl + CTL_RES_CUR + a;
        r.out_queue = c.q_meshq[b];
        r.out_count = c.ctl + CTL_MESH_N + b;
        r.zero[0] = c.ctl + CTL_RES_N + b;
        r.zero[1] = c.ctl + CTL_RES_CUR + b;
        r.zero[2] = r.zero[3] = NULL;
        k_split_top<ANY, COUNT, false, IO><<<grid, RT_BLOCK, 0, st>>>(c.sc, io, c.split, r, c.totals, -1);
        launches += 2;
    }
}

// Small kernel that does the queue bookkeeping k_trace_paths does in unified mode
__global__ void k_stage_prologue(const __grid_constant__ RenderCtx c, int cur)
{
    // counters of the queues this bounce's path trace and shade kernels fill
    for (int b = 0; b < RT_QBINS; ++b) c.ctl[CTL_PATH(cur ^ 1) + b] = 0;
    for (int b = 0; b < RT_SBINS; ++b) c.ctl[CTL_SHADE + b] = 0;
    for (int b = 0; b < RT_LBINS; ++b) c.ctl[CTL_LITB + b] = 0;      // k_shade regroups the lit paths for light sample 0
    c.ctl[CTL_LIT] = 0;
}

// k_shade, one specialised launch per hit shape (= bin of the shade queue) when the scene has fewer shapes than bins
#ifndef RT_SHADE_SPECIALISE
#define RT_SHADE_SPECIALISE 1
#endif
template <int ST>
static void rt_launch_shade_bin(const RenderCtx& c, int cur, uint32_t bounce, uint32_t bin, int first, uint32_t brdf,
                                unsigned grid, cudaStream_t st)
{
    switch (brdf)
    {
    case RT_BRDF_LAMBERT: k_shade_bin<ST, RT_BRDF_LAMBERT><<<grid, RT_BLOCK, 0, st>>>(c, cur, bounce, bin, first); break;
    case RT_BRDF_GLOSSY:  k_shade_bin<ST, RT_BRDF_GLOSSY><<<grid, RT_BLOCK, 0, st>>>(c, cur, bounce, bin, first); break;
    case RT_BRDF_MIRROR:  k_shade_bin<ST, RT_BRDF_MIRROR><<<grid, RT_BLOCK, 0, st>>>(c, cur, bounce, bin, first); break;
    default:              k_shade_bin<ST, RT_BRDF_NONE><<<grid, RT_BLOCK, 0, st>>>(c, cur, bounce, bin, first); break;
    }
}
// Grid of ONE bin's launch.  `grid` is sized by the batch; most bins hold a small fraction of it (a frame of config C4: the
// floor plane's bin 26 M entries, the rectangle light's a few thousand), and 75 776 blocks take ~100 us just to start and
// find nothing.  Once a count of the bin is known from an earlier batch (RenderBuffers.bin_hint_used: the largest seen, with
// a quarter on top) the bin gets that many blocks, never fewer than eight per SM.  The kernels are grid-stride loops, so any
// size is correct; only the time changes.
static unsigned rt_bin_grid(const RtScene* s, unsigned hint_index, unsigned grid, int sms)
{
    const RenderBuffers* rb = s->render;
    if (!rb->bin_hint_valid)
        return grid;
    const uint64_t want = ((uint64_t)rb->bin_hint_used[hint_index] * 5 / 4 + RT_BLOCK - 1) / RT_BLOCK + 1;
    return (unsigned)std::min<uint64_t>(grid, std::max<uint64_t>(want, (uint64_t)sms * 8));
}

static void rt_launch_shade(RtScene* s, const RenderCtx& c, int cur, uint32_t bounce, unsigned grid_all, int sms, cudaStream_t st,
                            uint64_t& launches)
{
    unsigned grid = grid_all;
    const size_t ns = s->shape_types.size();
    if (!RT_SHADE_SPECIALISE || ns == 0 || ns > RT_SBINS - 1 || s->d.stage6)
    {
        k_shade<<<grid, RT_BLOCK, 0, st>>>(c, cur, bounce);
        return;
    }
    for (size_t b = 0; b < ns; ++b)
    {
        const int first = b == 0 ? 1 : 0;
        const uint32_t brdf = s->shape_brdfs[b];
        grid = rt_bin_grid(s, (unsigned)b, grid_all, sms);
        switch (s->shape_types[b])
        {
        case RT_SHAPE_PLANE:  rt_launch_shade_bin<RT_SHAPE_PLANE>(c, cur, bounce, (uint32_t)b, first, brdf, grid, st); break;
        case RT_SHAPE_SPHERE: rt_launch_shade_bin<RT_SHAPE_SPHERE>(c, cur, bounce, (uint32_t)b, first, brdf, grid, st); break;
        case RT_SHAPE_RECT:   rt_launch_shade_bin<RT_SHAPE_RECT>(c, cur, bounce, (uint32_t)b, first, brdf, grid, st); break;
        default:              rt_launch_shade_bin<RT_SHAPE_MESH>(c, cur, bounce, (uint32_t)b, first, brdf, grid, st); break;
        }
    }
    launches += ns - 1;         // the caller counts one launch for this stage
}

// k_light_sample, one specialised launch per (light, BRDF kind) bin when the scene allows it
#ifndef RT_LIGHT_SPECIALISE
#define RT_LIGHT_SPECIALISE 1
#endif
template <int LT>
static void rt_launch_light_bin(const RenderCtx& c, int cur, uint32_t bounce, uint32_t lsi, uint32_t bin, bool glossy,
                                unsigned grid, cudaStream_t st)
{
    if (glossy) k_light_sample_bin<LT, RT_BRDF_GLOSSY><<<grid, RT_BLOCK, 0, st>>>(c, cur, bounce, lsi, bin);
    else        k_light_sample_bin<LT, RT_BRDF_LAMBERT><<<grid, RT_BLOCK, 0, st>>>(c, cur, bounce, lsi, bin);
}
static void rt_launch_light_sample(RtScene* s, const RenderCtx& c, int cur, uint32_t bounce, uint32_t lsi, unsigned grid_all, int sms,
                                   cudaStream_t st, uint64_t& launches)
{
    unsigned grid = grid_all;
    // bin = (light index & 3) | (glossy ? 4 : 0); a bin's light kind is known when every light that maps to it is
    // of one kind (always, with at most four lights)
    const size_t nl = s->light_types.size();
    bool uniform = RT_LIGHT_SPECIALISE && nl > 0 && !s->d.stage6;
    int kind[4] = { -1, -1, -1, -1 };
    for (size_t l = 0; l < nl && uniform; ++l)
    {
        int& k = kind[l & 3];
        if (k >= 0 && k != (int)s->light_types[l]) uniform = false;
        k = (int)s->light_types[l];
    }
    if (!uniform)
    {
        k_light_sample<<<grid, RT_BLOCK, 0, st>>>(c, cur, bounce, lsi);
        return;
    }
    unsigned extra = 0;
    for (uint32_t bin = 0; bin < RT_LBINS; ++bin)
    {
        const int lt = kind[bin & 3];
        const bool glossy = (bin & 4u) != 0;
        if (lt < 0 || (glossy ? !s->has_glossy : !s->has_lambert))
            continue;               // nothing can land in this bin
        grid = rt_bin_grid(s, RT_SBINS + bin, grid_all, sms);
        if (lt == RT_SHAPE_RECT)        rt_launch_light_bin<RT_SHAPE_RECT>(c, cur, bounce, lsi, bin, glossy, grid, st);
        else if (lt == RT_SHAPE_SPHERE) rt_launch_light_bin<RT_SHAPE_SPHERE>(c, cur, bounce, lsi, bin, glossy, grid, st);
        else                            rt_launch_light_bin<RT_SHAPE_MESH>(c, cur, bounce, lsi, bin, glossy, grid, st);
        ++extra;
    }
    if (extra > 1)
        launches += extra - 1;      // the caller counts one launch for this stage
}

namespace rt_detail
{
struct RenderGrids { int sms, path, shadow, mis, split_top, split_mesh64, split_mesh32; };
}

template <bool COUNT>
static int rt_launch_batch(RtScene* s, const RenderCtx& c, cudaStream_t st, uint64_t& launches, bool timed, uint64_t& trace_launches, bool split)
{
    RenderBuffers* rb = s->render;
    const int cap = s->stack_cap;
    // Grid sizes of the persistent kernels: exactly as many blocks as can be resident.  The occupancy queries behind them
    // are host calls the GPU waits for at the start of every frame, so they are made once per device and kernel family.
    typedef rt_detail::RenderGrids Grids;
    static std::mutex grids_lock;                                 // (one pair per instantiation, i.e. per COUNT)
    static std::map<std::pair<int, int>, Grids> grids_known;      // (device, stack capacity class)
    Grids g;
    {
        const int cap_class = cap <= 32 ? 32 : (cap <= 64 ? 64 : 104);
        std::lock_guard<std::mutex> guard(grids_lock);
        auto it = grids_known.find(std::make_pair(s->device, cap_class));
        if (it == grids_known.end())
        {
            g.sms = 148;
            cudaDeviceGetAttribute(&g.sms, cudaDevAttrMultiProcessorCount, s->device);
            g.path = g.shadow = g.mis = g.split_top = g.split_mesh64 = g.split_mesh32 = 4;
            if (cap_class == 32)
            {
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.path, k_trace_paths<32, COUNT>, RT_BLOCK, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.shadow, k_trace_shadow<32, COUNT>, RT_BLOCK, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.mis, k_trace_mis<32, COUNT>, RT_BLOCK, 0);
            }
            else if (cap_class == 64)
            {
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.path, k_trace_paths<64, COUNT>, RT_BLOCK, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.shadow, k_trace_shadow<64, COUNT>, RT_BLOCK, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.mis, k_trace_mis<64, COUNT>, RT_BLOCK, 0);
            }
            else
            {
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.path, k_trace_paths<104, COUNT>, RT_BLOCK, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.shadow, k_trace_shadow<104, COUNT>, RT_BLOCK, 0);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.mis, k_trace_mis<104, COUNT>, RT_BLOCK, 0);
            }
            // split kernels are lighter; their persistent grid is sized from the heaviest of them
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.split_top, k_split_top<false, COUNT, true, PathIO>, RT_BLOCK, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.split_mesh64, k_split_mesh<64, false, COUNT, PathIO>, RT_BLOCK, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.split_mesh32, k_split_mesh<32, false, COUNT, PathIO>, RT_BLOCK, 0);
            cudaGetLastError();
            grids_known[std::make_pair(s->device, cap_class)] = g;
        }
        else
            g = it->second;
    }
    const int dev_sms = g.sms;
    // Grid-stride kernels: blocks per SM (RT_WIDE_*_PER_SM; the environment overrides them for A/B runs)
    static const int wide_shade_per_sm = std::getenv("RAYITO_B200_WIDE_SHADE") ? std::atoi(std::getenv("RAYITO_B200_WIDE_SHADE")) : RT_WIDE_SHADE_PER_SM;
    static const int wide_gen_per_sm = std::getenv("RAYITO_B200_WIDE_GEN") ? std::atoi(std::getenv("RAYITO_B200_WIDE_GEN")) : RT_WIDE_GEN_PER_SM;
    const uint64_t sample_blocks = ((uint64_t)c.num_samples + RT_BLOCK - 1) / RT_BLOCK;
    const unsigned wide = (unsigned)std::min<uint64_t>(sample_blocks, (uint64_t)dev_sms * std::max(wide_shade_per_sm, 1));
    const unsigned wide_gen = (unsigned)std::min<uint64_t>(sample_blocks, (uint64_t)dev_sms * std::max(wide_gen_per_sm, 1));
    const unsigned pix_blocks = (c.num_pixels + RT_BLOCK - 1) / RT_BLOCK;
    const unsigned tg_path = (unsigned)(dev_sms * std::max(g.path, 1));
    const unsigned tg_shadow = (unsigned)(dev_sms * std::max(g.shadow, 1));
    const unsigned tg_mis = (unsigned)(dev_sms * std::max(g.mis, 1));
#ifndef RT_SEPARATE_MESH_GRID
#define RT_SEPARATE_MESH_GRID 1      /* +0.5 % in same-session A/B */
#endif
    const unsigned tg_split = (unsigned)(dev_sms * std::max(std::min(g.split_top, std::min(g.split_mesh64, g.split_mesh32)), 1));
    const unsigned tg_mesh = RT_SEPARATE_MESH_GRID ? (unsigned)(dev_sms * std::max(s->mesh_stack_need > 32 ? g.split_mesh64 : g.split_mesh32, 1)) : tg_split;

    // bin counts of an earlier batch, if their copy has arrived (see rt_bin_grid)
    if (rb->bin_hint_pending && cudaEventQuery(rb->bin_hint_ev) == cudaSuccess)
    {
        std::memcpy(rb->bin_hint_used, rb->h_bin_hint, sizeof(rb->bin_hint_used));
        rb->bin_hint_valid = true;
        rb->bin_hint_pending = false;
    }
    cudaGetLastError();     // (cudaErrorNotReady from the query is not an error)

    k_pixel_setup<<<pix_blocks, RT_BLOCK, 0, st>>>(c);
    k_raygen<<<wide_gen, RT_BLOCK, 0, st>>>(c);
    launches += 2;
    int cur = 0;
    for (uint32_t b = 0; b < c.depth; ++b)
    {
        rt_trace_mark(rb, timed, st);
        if (split)
        {
            uint64_t before = launches;
            k_stage_prologue<<<1, 1, 0, st>>>(c, cur);
            PathIO io = make_path_io(c, cur);
            rt_launch_split_stage<false, COUNT>(s, c, io, c.ctl + CTL_CUR_PATH, 0, tg_split, tg_mesh, st, launches);
            launches += 1;
            trace_launches += launches - before;
            launches -= 1;      // the shared "+= 2" below accounts for trace + shade
        }
        else
        {
            k_stage_prologue<<<1, 1, 0, st>>>(c, cur);
            launches += 1;
            if (cap <= 32)      k_trace_paths<32, COUNT><<<tg_path, RT_BLOCK, 0, st>>>(c, cur);
            else if (cap <= 64) k_trace_paths<64, COUNT><<<tg_path, RT_BLOCK, 0, st>>>(c, cur);
            else                k_trace_paths<104, COUNT><<<tg_path, RT_BLOCK, 0, st>>>(c, cur);
            trace_launches += 1;
        }
        rt_trace_mark(rb, timed, st);
        rt_launch_shade(s, c, cur, b, wide, dev_sms, st, launches);
        launches += 2;
        for (uint32_t l = 0; l < c.nls; ++l)
        {
            if (l > 0)
            {
                // (light sample 0 was regrouped by k_shade itself)
                k_light_select<<<wide, RT_BLOCK, 0, st>>>(c, cur, b, l);
                launches += 1;
            }
            rt_launch_light_sample(s, c, cur, b, l, wide, dev_sms, st, launches);
            rt_trace_mark(rb, timed, st);
            if (split)
            {
                uint64_t before = launches;
                ShadowIO sio = make_shadow_io(c);
                rt_launch_split_stage<true, COUNT>(s, c, sio, c.ctl + CTL_CUR_SHADOW, 1, tg_split, tg_mesh, st, launches);
                MisIO mio = make_mis_io(c);
                rt_launch_split_stage<false, COUNT>(s, c, mio, c.ctl + CTL_CUR_MIS, 0, tg_split, tg_mesh, st, launches);
                trace_launches += launches - before;
                launches -= 2;  // the shared "+= 4" below accounts for two traces
            }
            else
            {
                if (cap <= 32)      { k_trace_shadow<32, COUNT><<<tg_shadow, RT_BLOCK, 0, st>>>(c);  k_trace_mis<32, COUNT><<<tg_mis, RT_BLOCK, 0, st>>>(c); }
                else if (cap <= 64) { k_trace_shadow<64, COUNT><<<tg_shadow, RT_BLOCK, 0, st>>>(c);  k_trace_mis<64, COUNT><<<tg_mis, RT_BLOCK, 0, st>>>(c); }
                else                { k_trace_shadow<104, COUNT><<<tg_shadow, RT_BLOCK, 0, st>>>(c); k_trace_mis<104, COUNT><<<tg_mis, RT_BLOCK, 0, st>>>(c); }
                trace_launches += 2;
            }
            rt_trace_mark(rb, timed, st);
            k_resolve<<<wide, RT_BLOCK, 0, st>>>(c, l);
            launches += 4;
        }
        cur ^= 1;
    }
    k_accumulate<<<pix_blocks, RT_BLOCK, 0, st>>>(c);
    launches += 1;
    if (rb->d_bin_hint != NULL && !rb->bin_hint_pending)
    {
        cudaMemcpyAsync(rb->h_bin_hint, rb->d_bin_hint, sizeof(rb->bin_hint_used), cudaMemcpyDeviceToHost, st);
        cudaEventRecord(rb->bin_hint_ev, st);
        rb->bin_hint_pending = true;
    }
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

inline int rt_fill_ctx(RtScene* s, const RtCamera* camera, const RtRenderParams* prm, const RenderPlan& plan, RenderCtx& c)
{
    c = s->render->ctx;
    c.bin_hint = s->render->d_bin_hint;
    c.sc = s->d;
    c.cam = *camera;
    c.width = prm->width;
    c.height = prm->height;
    c.ps = prm->pixel_samples_hint;
    c.ls = prm->light_samples_hint;
    c.depth = prm->max_ray_depth;
    c.spp = plan.spp;
    c.spp_shift = 0xffffffffu;
    if ((plan.spp & (plan.spp - 1u)) == 0u)
        for (uint32_t k = 0; k < 32; ++k)
            if ((1u << k) == plan.spp) c.spp_shift = k;
    // samplers.m_numLightSamples = lights.empty() ? 0 : ls*ls  (RaytraceMain.cpp:77)
    c.ls2 = prm->light_samples_hint * prm->light_samples_hint;
    c.nls = s->d.num_lights == 0 ? 0 : c.ls2;
    if (s->d.stage6)
        c.nls = s->d.num_lights * c.ls2;
    c.tile = plan.tile;
    c.tiles_x = plan.tiles_x;
    c.aspect = (float)prm->width / (float)prm->height;
    return RT_OK;
}

// packed_floats != 0: rgb_out is a DEVICE buffer of that many floats receiving this rank's tiles packed
// one after the other (rt_packed_floats) instead of a width*height*3 frame
inline int rt_render_impl(RtScene* s, const RtCamera* camera, const RtRenderParams* prm, float* rgb_out, bool out_on_device,
                          RtRenderStats* stats, cudaStream_t st, size_t packed_floats = 0)
{
    if (s == NULL || camera == NULL || prm == NULL || rgb_out == NULL)
        return rt_fail(RT_ERR_ARG, "null argument");
    RT_CUDA(cudaSetDevice(s->device));
    // RAYITO_B200_TIMING=1: host-clock phases of this call on stderr
    const bool host_timing = std::getenv("RAYITO_B200_TIMING") != NULL;
    std::chrono::steady_clock::time_point hc[5];
    hc[0] = std::chrono::steady_clock::now();
    RenderPlan plan;
    int rc = rt_plan(s, prm, plan);
    if (rc != RT_OK) return rc;
    hc[1] = std::chrono::steady_clock::now();
    rc = rt_render_reserve(s, plan);
    if (rc != RT_OK) return rc;
    RenderBuffers* rb = s->render;
    hc[2] = std::chrono::steady_clock::now();

    const size_t image_floats = (size_t)prm->width * prm->height * 3;
    float* d_image = rgb_out;
    if (!out_on_device)
    {
        if (rb->image_floats < image_floats)
        {
            rt_detail::pool_free(s->device, rb->d_image, rb->image_bytes);
            rb->d_image = NULL;
            rb->image_floats = 0;
            RT_CUDA(rt_detail::pool_alloc(s->device, (void**)&rb->d_image, image_floats * sizeof(float), &rb->image_bytes));
            rb->image_floats = image_floats;
        }
        d_image = rb->d_image;
        RT_CUDA(cudaMemsetAsync(d_image, 0, image_floats * sizeof(float), st));
    }

    RenderCtx c;
    rt_fill_ctx(s, camera, prm, plan, c);
    c.image = d_image;
    c.tile_ids = rb->d_tile_ids;
    c.packed = packed_floats != 0 ? 1u : 0u;
    c.tile_base = 0;
    if (packed_floats != 0 && (!out_on_device || packed_floats < plan.tiles.size() * (size_t)plan.tile * plan.tile * 3))
        return rt_fail(RT_ERR_ARG, "packed tile buffer must be a device buffer of at least rt_packed_floats() floats");

    RT_CUDA(cudaEventRecord(rb->ev[0], st));
    if (!plan.tiles.empty())
        RT_CUDA(cudaMemcpyAsync(rb->d_tile_ids, plan.tiles.data(), plan.tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    RT_CUDA(cudaMemsetAsync(c.totals, 0, 8 * sizeof(uint64_t), st));
    RT_CUDA(cudaEventRecord(rb->ev[1], st));

    const bool count = (prm->flags & RT_RENDER_COUNT_WORK) != 0;
    const bool timed = (prm->flags & RT_RENDER_TIME_TRACE) != 0;
    s->dynamic_top = (prm->flags & RT_RENDER_DYNAMIC_TOP) != 0;
    const bool split = (prm->flags & RT_RENDER_UNIFIED_TRAVERSAL) == 0 && s->num_mesh_shapes >= 1 &&
                       s->num_mesh_shapes <= RT_SPLIT_MAX_MESHES && s->top_stack_need <= RT_SPLIT_TOPCAP &&
                       s->mesh_stack_need <= 64;
    rb->trace_events_used = 0;
    uint64_t launches = 0, trace_launches = 0;
    uint64_t samples = 0;
    for (size_t t0 = 0; t0 < plan.tiles.size(); t0 += plan.tiles_per_batch)
    {
        size_t nt = std::min<size_t>(plan.tiles_per_batch, plan.tiles.size() - t0);
        c.tile_ids = rb->d_tile_ids + t0;
        c.tile_base = (uint32_t)t0;
        c.num_pixels = (uint32_t)(nt * plan.tile * plan.tile);
        c.num_samples = c.num_pixels * plan.spp;
        rc = count ? rt_launch_batch<true>(s, c, st, launches, timed, trace_launches, split)
                   : rt_launch_batch<false>(s, c, st, launches, timed, trace_launches, split);
        if (rc != RT_OK) return rc;
    }
    RT_CUDA(cudaEventRecord(rb->ev[2], st));

    uint64_t totals[8] = { 0 };
    RT_CUDA(cudaMemcpyAsync(totals, c.totals, sizeof(totals), cudaMemcpyDeviceToHost, st));
    if (!out_on_device)
    {
        if (prm->world == 1)
        {
            RT_CUDA(cudaMemcpyAsync(rgb_out, d_image, image_floats * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        else
        {
            // Only this rank's tiles are written back; other pixels stay untouched
            for (size_t k = 0; k < plan.tiles.size(); ++k)
            {
                uint32_t tx = plan.tiles[k] % plan.tiles_x, ty = plan.tiles[k] / plan.tiles_x;
                uint32_t x0 = tx * plan.tile, y0 = ty * plan.tile;
                uint32_t w = std::min(plan.tile, prm->width - x0), h = std::min(plan.tile, prm->height - y0);
                size_t off = ((size_t)y0 * prm->width + x0) * 3;
                RT_CUDA(cudaMemcpy2DAsync(rgb_out + off, (size_t)prm->width * 12, d_image + off, (size_t)prm->width * 12,
                                          (size_t)w * 12, h, cudaMemcpyDeviceToHost, st));
            }
        }
    }
    RT_CUDA(cudaEventRecord(rb->ev[3], st));
    hc[3] = std::chrono::steady_clock::now();
    RT_CUDA(cudaStreamSynchronize(st));
    hc[4] = std::chrono::steady_clock::now();
    if (host_timing)
    {
        double ms[4];
        for (int i = 0; i < 4; ++i)
            ms[i] = std::chrono::duration<double, std::milli>(hc[i + 1] - hc[i]).count();
        std::fprintf(stderr, "[rayito_b200] rt_render host clock: plan %.1f ms, reserve %.1f, enqueue %.1f, wait %.1f\n",
                     ms[0], ms[1], ms[2], ms[3]);
    }

    for (size_t k = 0; k < plan.tiles.size(); ++k)
    {
        uint32_t tx = plan.tiles[k] % plan.tiles_x, ty = plan.tiles[k] / plan.tiles_x;
        uint32_t x0 = tx * plan.tile, y0 = ty * plan.tile;
        uint32_t w = std::min(plan.tile, prm->width - x0), h = std::min(plan.tile, prm->height - y0);
        samples += (uint64_t)w * h * plan.spp;
    }
    if (stats != NULL)
    {
        std::memset(stats, 0, sizeof(*stats));
        stats->samples = samples;
        stats->closest_rays = totals[0];
        stats->any_rays = totals[1];
        stats->node_pops = totals[2];
        stats->tri_tests = totals[3];
        stats->shape_tests = totals[4];
        stats->xform_evals = totals[5];
        stats->xform_keyed = totals[6];
        stats->xform_pairs = totals[7];
        stats->kernel_launches = launches;
        cudaEventElapsedTime(&stats->upload_ms, rb->ev[0], rb->ev[1]);
        cudaEventElapsedTime(&stats->render_ms, rb->ev[1], rb->ev[2]);
        cudaEventElapsedTime(&stats->download_ms, rb->ev[2], rb->ev[3]);
        stats->trace_ms = 0.0f;
        stats->trace_launches = trace_launches;
        if (timed && rb->trace_events)
        {
            double total = 0.0;
            for (size_t k = 0; k + 1 < rb->trace_events_used; k += 2)
            {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, (*rb->trace_events)[k], (*rb->trace_events)[k + 1]);
                total += ms;
            }
            stats->trace_ms = (float)total;
        }
    }
    return RT_OK;
}

inline int rt_camera_rays_impl(RtScene* s, const RtCamera* camera, const RtRenderParams* prm, uint32_t psi, RtRay* rays)
{
    if (s == NULL || camera == NULL || prm == NULL || rays == NULL)
        return rt_fail(RT_ERR_ARG, "null argument");
    RT_CUDA(cudaSetDevice(s->device));
    RenderPlan plan;
    int rc = rt_plan(s, prm, plan);
    if (rc != RT_OK) return rc;
    if (psi >= plan.spp)
        return rt_fail(RT_ERR_ARG, "psi out of range");
    rc = rt_render_reserve(s, plan);
    if (rc != RT_OK) return rc;
    RenderBuffers* rb = s->render;
    size_t n = (size_t)prm->width * prm->height;
    rc = RT_OK;
    RtRay* d_rays = NULL;
    RT_CUDA(cudaMalloc((void**)&d_rays, n * sizeof(RtRay)));
    cudaMemset(d_rays, 0, n * sizeof(RtRay));
    RenderCtx c;
    rt_fill_ctx(s, camera, prm, plan, c);
    c.image = NULL;
    c.packed = 0;
    c.tile_base = 0;
    if (!plan.tiles.empty())
        cudaMemcpy(rb->d_tile_ids, plan.tiles.data(), plan.tiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    for (size_t t0 = 0; t0 < plan.tiles.size(); t0 += plan.tiles_per_batch)
    {
        size_t nt = std::min<size_t>(plan.tiles_per_batch, plan.tiles.size() - t0);
        c.tile_ids = rb->d_tile_ids + t0;
        c.num_pixels = (uint32_t)(nt * plan.tile * plan.tile);
        c.num_samples = c.num_pixels * plan.spp;
        unsigned blocks = (c.num_pixels + RT_BLOCK - 1) / RT_BLOCK;
        k_pixel_setup<<<blocks, RT_BLOCK>>>(c);
        k_camera_rays<<<blocks, RT_BLOCK>>>(c, psi, d_rays);
    }
    cudaError_t e = cudaMemcpy(rays, d_rays, n * sizeof(RtRay), cudaMemcpyDeviceToHost);
    cudaFree(d_rays);
    if (e != cudaSuccess) return rt_cuda_fail(e, "camera rays");
    return RT_OK;
}

// Same on device buffers (the frame rt_render_device / rt_render_multi left in HBM), enqueued on `st`
inline int rt_tonemap_device_impl(int device, const float* d_rgb, size_t num_pixels, float exposure_stops, float gamma,
                                  uint8_t* d_bgra, cudaStream_t st)
{
    if (d_rgb == NULL || d_bgra == NULL)
        return rt_fail(RT_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return rt_fail(RT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    RT_CUDA(cudaSetDevice(device));
    if (num_pixels == 0) return RT_OK;
    float gamma_exp = 1.0f / gamma;
    float exposure = std::pow(2.0f, exposure_stops);
    k_tonemap<<<(unsigned)((num_pixels + 255) / 256), 256, 0, st>>>(d_rgb, num_pixels, exposure, gamma_exp, d_bgra);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

inline int rt_tonemap_impl(int device, const float* rgb_in, size_t num_pixels, float exposure_stops, float gamma, uint8_t* bgra)
{
    if (rgb_in == NULL || bgra == NULL)
        return rt_fail(RT_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return rt_fail(RT_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    RT_CUDA(cudaSetDevice(device));
    if (num_pixels == 0) return RT_OK;
    float* d_in = NULL;
    uint8_t* d_out = NULL;
    RT_CUDA(cudaMalloc((void**)&d_in, num_pixels * 12));
    cudaError_t e = cudaMalloc((void**)&d_out, num_pixels * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d_in, rgb_in, num_pixels * 12, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        // gammaExponent = 1.0f / gamma; exposure = pow(2.0f, stops)  (MainWindow.cpp:43-45)
        float gamma_exp = 1.0f / gamma;
        float exposure = std::pow(2.0f, exposure_stops);
        k_tonemap<<<(unsigned)((num_pixels + 255) / 256), 256>>>(d_in, num_pixels, exposure, gamma_exp, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(bgra, d_out, num_pixels * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    if (e != cudaSuccess) return rt_cuda_fail(e, "tonemap");
    return RT_OK;
}

#endif // RAYITO_B200_RT_RENDER_CUH
