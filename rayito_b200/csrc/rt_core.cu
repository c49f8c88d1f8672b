// librayito_b200.so -- C ABI (include/rayito_b200.h) and the ray-batch kernels.
//
// Built for sm_100a only, with -fmad=false (see rt_device.cuh for the arithmetic
// contract).  There is deliberately no CPU implementation in this library: every
// compute entry point needs a CUDA device and fails loudly without one.
#include <cstdio>
#include <string>

#include "rt_scene.cuh"
#include "rt_build.cuh"
#include "rt_trace.cuh"
#include "rt_wave.cuh"
#include "rt_render.cuh"
#include "rt_comm.cuh"
#include "rt_stage23.cuh"

thread_local std::string g_rt_error;

int rt_fail(int code, const std::string& what)
{
    g_rt_error = what;
    return code;
}

int rt_cuda_fail(cudaError_t e, const char* where)
{
    g_rt_error = std::string("CUDA error at ") + where + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return RT_ERR_CUDA;
}

// ---------------------------------------------------------------------------
// Ray-batch kernels: one ray per thread.
// ---------------------------------------------------------------------------

__device__ __forceinline__ void load_ray(const RtRay* rays, size_t i, V3& o, V3& d, float& tmax, float& time)
{
    const float4* p = reinterpret_cast<const float4*>(rays + i);
    float4 a = __ldg(p);
    float4 b = __ldg(p + 1);
    o = mk(a.x, a.y, a.z);
    d = mk(a.w, b.x, b.y);
    tmax = b.z;
    time = b.w;
}

__device__ __forceinline__ void flush_work(const WorkCount& wc, uint64_t* work)
{
    // Warp-reduce, then one atomic per counter per warp
    uint32_t v[RT_WORK_COUNTERS] = { wc.node_pops, wc.tri_tests, wc.shape_tests, wc.xform_evals, wc.xform_keyed, wc.xform_pairs };
    for (int k = 0; k < RT_WORK_COUNTERS; ++k)
    {
        uint32_t x = v[k];
        for (int off = 16; off > 0; off >>= 1)
            x += __shfl_down_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0 && x)
            atomicAdd(reinterpret_cast<unsigned long long*>(work + k), (unsigned long long)x);
    }
}

// Ray-batch adaptors for the wave traversal
struct BatchIO
{
    const RtRay* rays;
    float4* raw;        // closest: (t, shape, triangle record, -)
    uint8_t* any;       // any hit: 0 / 1
    __device__ __forceinline__ bool load(uint32_t j, V3& o, V3& d, float& tmax, float& time, uint32_t& tag) const
    {
        tag = j;
        load_ray(rays, j, o, d, tmax, time);
        return true;
    }
    __device__ __forceinline__ const float4* xf_row(uint32_t) const { return nullptr; }     // ray batches: no per-sample cache
    __device__ __forceinline__ void store(uint32_t tag, const WaveResult& r) const
    {
        if (raw) raw[tag] = make_float4(r.t, __int_as_float(r.shape), __int_as_float(r.tri_rec), 0.0f);
        else any[tag] = r.any_hit ? 1 : 0;
    }
};

template <int CAP, bool ANY, bool COUNT>
__global__ void __launch_bounds__(128)
k_trace_batch(const __grid_constant__ DScene sc, const RtRay* __restrict__ rays, uint32_t n,
              float4* raw, uint8_t* any, uint32_t* cursor, uint64_t* work)
{
    WorkCount wc = RT_WORK_ZERO;
    BatchIO io = { rays, raw, any };
    trace_wave<CAP, ANY, COUNT>(sc, io, n, cursor, wc);
    if (COUNT)
        flush_work(wc, work);
}

// (t, shape, triangle record) -> RtHit / RtHitEx: face and fan-triangle ids, and for
// the extended record the shading normal and colour modifier (hit_shading_inputs)
template <bool EX>
__global__ void __launch_bounds__(128)
k_finalize_batch(const __grid_constant__ DScene sc, const RtRay* __restrict__ rays, uint32_t n,
                 const float4* raw, void* __restrict__ hits)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float4 r = raw[i];
    ClosestHit h;
    h.t = r.x;
    h.shape = __float_as_int(r.y);
    h.tri_rec = __float_as_int(r.z);
    int32_t face = -1, tri = -1;
    if (h.tri_rec >= 0)
    {
        face = (int32_t)__float_as_uint(__ldg(sc.tris + 3 * (size_t)h.tri_rec + 0).w);
        tri = (int32_t)__float_as_uint(__ldg(sc.tris + 3 * (size_t)h.tri_rec + 1).w);
    }
    float4 first = make_float4(h.t, __int_as_float(h.shape), __int_as_float(face), __int_as_float(tri));
    if (EX)
    {
        V3 o, d;
        float tmax, time;
        load_ray(rays, i, o, d, tmax, time);
        TRS set_trs = xform_eval(sc, sc.set_xform, time);
        LocalRay r0;
        r0.o = to_local_point(set_trs, o);
        r0.d = to_local_vector(set_trs, d);
        r0.inv = r0.d; r0.neg = 0;
        V3 nrm;
        float cm;
        hit_shading_inputs(sc, r0, time, h, nrm, cm);
        float4* out = reinterpret_cast<float4*>(static_cast<RtHitEx*>(hits) + i);
        out[0] = first;
        out[1] = make_float4(nrm.x, nrm.y, nrm.z, cm);
    }
    else
    {
        reinterpret_cast<float4*>(hits)[i] = first;     // may alias raw: same slot, read before write
    }
}

static unsigned persistent_blocks(const void* kernel, int device)
{
    int sms = 148, per_sm = 4;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0) != cudaSuccess || per_sm < 1)
        per_sm = 4;
    return (unsigned)(sms * per_sm);
}

template <bool ANY, bool COUNT>
static int launch_batch_trace(RtScene* s, const RtRay* d_rays, size_t n, float4* raw, uint8_t* any, uint64_t* d_work, cudaStream_t st)
{
    if (n == 0) return RT_OK;
    if (n >= 0xffffffffull) return rt_fail(RT_ERR_ARG, "ray batch too large (>= 2^32 rays)");
    // One work cursor per call, taken round-robin from the scene's 16 slots: calls enqueued on
    // different streams do not share a cursor (more than 16 calls in flight at once would).
    uint32_t* cursor = s->d_cursor + (s->cursor_next.fetch_add(1u) % RT_SCENE_CURSORS);
    RT_CUDA(cudaMemsetAsync(cursor, 0, sizeof(uint32_t), st));
    const uint32_t n32 = (uint32_t)n;
    if (s->stack_cap <= 32)
        k_trace_batch<32, ANY, COUNT><<<persistent_blocks((const void*)k_trace_batch<32, ANY, COUNT>, s->device), 128, 0, st>>>(s->d, d_rays, n32, raw, any, cursor, d_work);
    else if (s->stack_cap <= 64)
        k_trace_batch<64, ANY, COUNT><<<persistent_blocks((const void*)k_trace_batch<64, ANY, COUNT>, s->device), 128, 0, st>>>(s->d, d_rays, n32, raw, any, cursor, d_work);
    else
        k_trace_batch<104, ANY, COUNT><<<persistent_blocks((const void*)k_trace_batch<104, ANY, COUNT>, s->device), 128, 0, st>>>(s->d, d_rays, n32, raw, any, cursor, d_work);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

template <bool COUNT, bool EX>
static int launch_closest(RtScene* s, const RtRay* d_rays, size_t n, void* d_hits, float4* d_raw, uint64_t* d_work, cudaStream_t st)
{
    if (n == 0) return RT_OK;
    int rc = launch_batch_trace<false, COUNT>(s, d_rays, n, d_raw, NULL, d_work, st);
    if (rc != RT_OK) return rc;
    unsigned blocks = (unsigned)((n + 127) / 128);
    k_finalize_batch<EX><<<blocks, 128, 0, st>>>(s->d, d_rays, (uint32_t)n, d_raw, d_hits);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

template <bool COUNT>
static int launch_any(RtScene* s, const RtRay* d_rays, size_t n, uint8_t* d_hits, uint64_t* d_work, cudaStream_t st)
{
    return launch_batch_trace<true, COUNT>(s, d_rays, n, NULL, d_hits, d_work, st);
}

static int ensure_scratch(RtScene* s, size_t in_bytes, size_t out_bytes)
{
    if (in_bytes > s->scratch_in_bytes)
    {
        rt_detail::pool_free(s->device, s->scratch_in, s->scratch_in_bytes);
        s->scratch_in = NULL;
        s->scratch_in_bytes = 0;
        RT_CUDA(rt_detail::pool_alloc(s->device, &s->scratch_in, in_bytes, &in_bytes));
        s->scratch_in_bytes = in_bytes;
    }
    if (out_bytes > s->scratch_out_bytes)
    {
        rt_detail::pool_free(s->device, s->scratch_out, s->scratch_out_bytes);
        s->scratch_out = NULL;
        s->scratch_out_bytes = 0;
        RT_CUDA(rt_detail::pool_alloc(s->device, &s->scratch_out, out_bytes, &out_bytes));
        s->scratch_out_bytes = out_bytes;
    }
    return RT_OK;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C"
{

const char* rt_last_error_string(void) { return g_rt_error.c_str(); }

int rt_abi_version(void) { return RT_ABI_VERSION; }

int rt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rt_scene_create(const RtSceneDesc* desc, int device, RtScene** out_scene)
{
    return rt_scene_build(desc, device, 0u, out_scene);
}

int rt_scene_create_ex(const RtSceneDesc* desc, int device, uint32_t flags, RtScene** out_scene)
{
    return rt_scene_build(desc, device, flags, out_scene);
}

int rt_scene_mesh_nodes(RtScene* s, uint32_t mesh, RtBvhNode* nodes, uint32_t capacity, uint32_t* depth, float* build_ms)
{
    if (s == NULL || mesh >= s->mesh_faces.size() || (capacity && nodes == NULL))
        return rt_fail(RT_ERR_ARG, "null argument or mesh index out of range");
    RT_CUDA(cudaSetDevice(s->device));
    if (depth) *depth = s->mesh_depth < 0 ? 0u : (uint32_t)s->mesh_depth;      // (deepest of all meshes)
    if (build_ms) *build_ms = s->bvh_build_ms;
    const uint32_t faces = s->mesh_faces[mesh];
    const uint32_t count = std::min<uint32_t>(capacity, faces ? 2 * faces - 1 : 0u);
    if (count == 0)
        return RT_OK;
    std::vector<DNode> dn(count);
    std::vector<uint32_t> fft((size_t)faces + 1);
    RT_CUDA(cudaMemcpy(dn.data(), s->d.mesh_nodes + s->mesh_first_node[mesh], (size_t)count * sizeof(DNode), cudaMemcpyDeviceToHost));
    RT_CUDA(cudaMemcpy(fft.data(), s->d.face_first_tri + s->mesh_first_face[mesh], fft.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < count; ++i)
    {
        RtBvhNode& n = nodes[i];
        n.bbox_min[0] = dn[i].q0.x; n.bbox_min[1] = dn[i].q0.y; n.bbox_min[2] = dn[i].q0.z;
        n.bbox_max[0] = dn[i].q0.w; n.bbox_max[1] = dn[i].q1.x; n.bbox_max[2] = dn[i].q1.y;
        uint32_t word, flags;
        std::memcpy(&word, &dn[i].q1.z, 4);
        std::memcpy(&flags, &dn[i].q1.w, 4);
        if (flags & RT_NODE_LEAF)
        {
            // leaves are stored as (first triangle record, count): back to the face that owns the record
            const uint32_t* at = std::upper_bound(fft.data(), fft.data() + fft.size(), word);
            word = (uint32_t)(at - fft.data()) - 1u;
            flags = RT_NODE_LEAF;
        }
        n.first_child_or_prim = word;
        n.flags = flags;
    }
    return RT_OK;
}

int rt_scene_destroy(RtScene* s)
{
    if (s == NULL) return RT_OK;
    cudaSetDevice(s->device);
    // The *_device entry points do not synchronise and may still be running on a caller's
    // (non-blocking) stream: nothing of this scene may go back to the pool, where the next
    // rt_scene_create would overwrite it, before the device is idle.
    cudaDeviceSynchronize();
    rt_render_release(s);
    rt_detail::pool_free(s->device, s->scratch_in, s->scratch_in_bytes);
    rt_detail::pool_free(s->device, s->scratch_out, s->scratch_out_bytes);
    rt_detail::pool_free(s->device, s->d_work, s->work_alloc);
    rt_detail::pool_free(s->device, s->d_cursor, s->cursor_alloc);
    rt_detail::pool_free(s->device, s->arena, s->arena_alloc);
    delete s;
    return RT_OK;
}

int rt_trace_closest_device(RtScene* s, const RtRay* d_rays, size_t n, RtHit* d_hits, uint64_t* d_work, void* stream)
{
    if (s == NULL || (n && (d_rays == NULL || d_hits == NULL))) return rt_fail(RT_ERR_ARG, "null argument");
    RT_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the raw (t, shape, record) result is written into d_hits itself and converted in place
    float4* raw = reinterpret_cast<float4*>(d_hits);
    return d_work ? launch_closest<true, false>(s, d_rays, n, d_hits, raw, d_work, st)
                  : launch_closest<false, false>(s, d_rays, n, d_hits, raw, NULL, st);
}

int rt_trace_any_device(RtScene* s, const RtRay* d_rays, size_t n, uint8_t* d_hits, uint64_t* d_work, void* stream)
{
    if (s == NULL || (n && (d_rays == NULL || d_hits == NULL))) return rt_fail(RT_ERR_ARG, "null argument");
    RT_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return d_work ? launch_any<true>(s, d_rays, n, d_hits, d_work, st)
                  : launch_any<false>(s, d_rays, n, d_hits, NULL, st);
}

static int trace_closest_host(RtScene* s, const RtRay* rays, size_t n, void* hits, bool ex)
{
    if (s == NULL || (n && (rays == NULL || hits == NULL))) return rt_fail(RT_ERR_ARG, "null argument");
    if (n == 0) return RT_OK;
    RT_CUDA(cudaSetDevice(s->device));
    size_t out_bytes = n * (ex ? sizeof(RtHitEx) : sizeof(RtHit));
    // scratch_out holds the output records followed by the raw (t, shape, record) results
    size_t raw_off = (out_bytes + 255) & ~(size_t)255;
    int rc = ensure_scratch(s, n * sizeof(RtRay), raw_off + n * sizeof(float4));
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaMemcpy(s->scratch_in, rays, n * sizeof(RtRay), cudaMemcpyHostToDevice));
    float4* raw = reinterpret_cast<float4*>(static_cast<char*>(s->scratch_out) + raw_off);
    rc = ex ? launch_closest<false, true>(s, static_cast<const RtRay*>(s->scratch_in), n, s->scratch_out, raw, NULL, 0)
            : launch_closest<false, false>(s, static_cast<const RtRay*>(s->scratch_in), n, s->scratch_out, raw, NULL, 0);
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaMemcpy(hits, s->scratch_out, out_bytes, cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_trace_closest(RtScene* s, const RtRay* rays, size_t n, RtHit* hits)
{
    return trace_closest_host(s, rays, n, hits, false);
}

int rt_trace_closest_ex(RtScene* s, const RtRay* rays, size_t n, RtHitEx* hits)
{
    return trace_closest_host(s, rays, n, hits, true);
}

int rt_trace_any(RtScene* s, const RtRay* rays, size_t n, uint8_t* hits)
{
    if (s == NULL || (n && (rays == NULL || hits == NULL))) return rt_fail(RT_ERR_ARG, "null argument");
    if (n == 0) return RT_OK;
    RT_CUDA(cudaSetDevice(s->device));
    int rc = ensure_scratch(s, n * sizeof(RtRay), n);
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaMemcpy(s->scratch_in, rays, n * sizeof(RtRay), cudaMemcpyHostToDevice));
    rc = launch_any<false>(s, static_cast<const RtRay*>(s->scratch_in), n, static_cast<uint8_t*>(s->scratch_out), NULL, 0);
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaMemcpy(hits, s->scratch_out, n, cudaMemcpyDeviceToHost));
    return RT_OK;
}

// Host-buffer traces that also return the work counters (tests pin them to the oracle's)
static int trace_counted(RtScene* s, const RtRay* rays, size_t n, void* hits, bool any, uint64_t* work)
{
    if (s == NULL || work == NULL || (n && (rays == NULL || hits == NULL))) return rt_fail(RT_ERR_ARG, "null argument");
    std::memset(work, 0, RT_WORK_COUNTERS * sizeof(uint64_t));
    if (n == 0) return RT_OK;
    RT_CUDA(cudaSetDevice(s->device));
    int rc = ensure_scratch(s, n * sizeof(RtRay), any ? n : n * sizeof(RtHit));
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaMemcpy(s->scratch_in, rays, n * sizeof(RtRay), cudaMemcpyHostToDevice));
    RT_CUDA(cudaMemset(s->d_work, 0, RT_WORK_COUNTERS * sizeof(uint64_t)));
    const RtRay* d_rays = static_cast<const RtRay*>(s->scratch_in);
    rc = any ? launch_any<true>(s, d_rays, n, static_cast<uint8_t*>(s->scratch_out), s->d_work, 0)
             : launch_closest<true, false>(s, d_rays, n, s->scratch_out, static_cast<float4*>(s->scratch_out), s->d_work, 0);
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaMemcpy(hits, s->scratch_out, any ? n : n * sizeof(RtHit), cudaMemcpyDeviceToHost));
    RT_CUDA(cudaMemcpy(work, s->d_work, RT_WORK_COUNTERS * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_trace_closest_counted(RtScene* s, const RtRay* rays, size_t n, RtHit* hits, uint64_t* work)
{
    return trace_counted(s, rays, n, hits, false, work);
}

int rt_trace_any_counted(RtScene* s, const RtRay* rays, size_t n, uint8_t* hits, uint64_t* work)
{
    return trace_counted(s, rays, n, hits, true, work);
}

int rt_render(RtScene* s, const RtCamera* camera, const RtRenderParams* params, float* rgb, RtRenderStats* stats)
{
    return rt_render_impl(s, camera, params, rgb, false, stats, 0);
}

int rt_render_device(RtScene* s, const RtCamera* camera, const RtRenderParams* params, float* d_rgb,
                     RtRenderStats* stats, void* stream)
{
    return rt_render_impl(s, camera, params, d_rgb, true, stats, static_cast<cudaStream_t>(stream));
}

// ---- multi-GPU tile assembly (rt_comm.cuh) ----------------------------------------
int rt_comm_unique_id(uint8_t* id)
{
    if (id == NULL) return rt_fail(RT_ERR_ARG, "null argument");
    rt_detail::NcclApi* nccl = rt_detail::nccl_api();
    if (nccl->handle == NULL) return rt_fail(RT_ERR_COMM, nccl->why);
    static_assert(sizeof(ncclUniqueId) == RT_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId uid;
    RT_NCCL(nccl->GetUniqueId(&uid));
    std::memcpy(id, &uid, sizeof(uid));
    return RT_OK;
}

int rt_comm_create(const uint8_t* id, int rank, int world, int device, RtComm** out)
{
    if (id == NULL || out == NULL || world < 1 || rank < 0 || rank >= world) return rt_fail(RT_ERR_ARG, "bad argument");
    *out = NULL;
    RT_CUDA(cudaSetDevice(device));
    ncclComm_t comm = NULL;
    if (world > 1)
    {
        rt_detail::NcclApi* nccl = rt_detail::nccl_api();
        if (nccl->handle == NULL) return rt_fail(RT_ERR_COMM, nccl->why);
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof(uid));
        RT_NCCL(nccl->CommInitRank(&comm, world, uid, rank));
    }
    return rt_comm_wrap(comm, true, rank, world, device, out);
}

int rt_comm_from_nccl(void* nccl_comm, int device, RtComm** out)
{
    if (nccl_comm == NULL || out == NULL) return rt_fail(RT_ERR_ARG, "null argument");
    *out = NULL;
    rt_detail::NcclApi* nccl = rt_detail::nccl_api();
    if (nccl->handle == NULL) return rt_fail(RT_ERR_COMM, nccl->why);
    int rank = 0, world = 0;
    RT_NCCL(nccl->CommUserRank(static_cast<ncclComm_t>(nccl_comm), &rank));
    RT_NCCL(nccl->CommCount(static_cast<ncclComm_t>(nccl_comm), &world));
    return rt_comm_wrap(static_cast<ncclComm_t>(nccl_comm), false, rank, world, device, out);
}

int rt_comm_destroy(RtComm* c)
{
    if (c == NULL) return RT_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    rt_detail::pool_free(c->device, c->d_packed, c->packed_bytes);
    rt_detail::pool_free(c->device, c->d_tiles, c->tiles_bytes);
    rt_detail::pool_free(c->device, c->d_frame, c->frame_bytes);
    for (int i = 0; i < 3; ++i) cudaEventDestroy(c->ev[i]);
    if (c->owned && c->comm != NULL)
        rt_detail::nccl_api()->CommDestroy(c->comm);
    delete c;
    return RT_OK;
}

int rt_comm_rank(const RtComm* c, int* rank, int* world, int* device)
{
    if (c == NULL) return rt_fail(RT_ERR_ARG, "null argument");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (device) *device = c->device;
    return RT_OK;
}

int rt_render_multi_host(RtScene* s, const RtCamera* camera, const RtRenderParams* params, RtComm* comm, int root,
                         float* rgb, RtRenderStats* stats, float* assemble_ms)
{
    return rt_render_multi_host_impl(s, camera, params, comm, root, rgb, stats, assemble_ms);
}

size_t rt_packed_floats(uint32_t width, uint32_t height, uint32_t tile_size, uint32_t world, uint32_t rank)
{
    if (width == 0 || height == 0 || world == 0 || rank >= world) return 0;
    return rt_detail::packed_floats(width, height, tile_size, world, rank);
}

int rt_render_tiles_packed(RtScene* s, const RtCamera* camera, const RtRenderParams* params, float* d_packed,
                           size_t packed_floats, RtRenderStats* stats, void* stream)
{
    if (packed_floats == 0) return rt_fail(RT_ERR_ARG, "packed_floats must be rt_packed_floats() of this rank");
    return rt_render_impl(s, camera, params, d_packed, true, stats, static_cast<cudaStream_t>(stream), packed_floats);
}

int rt_unpack_tiles(int device, const float* d_packed, uint32_t width, uint32_t height, uint32_t tile_size,
                    uint32_t world, uint32_t rank, float* d_rgb, void* stream)
{
    if (d_packed == NULL || d_rgb == NULL || width == 0 || height == 0 || world == 0 || rank >= world)
        return rt_fail(RT_ERR_ARG, "bad argument");
    RT_CUDA(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t tile = tile_size ? tile_size : RT_DEFAULT_TILE;
    const uint32_t tx = (width + tile - 1) / tile, ty = (height + tile - 1) / tile;
    std::vector<uint32_t> mine;
    rt_detail::rank_tiles(tx, ty, rank, world, mine);
    if (mine.empty()) return RT_OK;
    uint32_t* d_ids = NULL;
    size_t got = 0;
    RT_CUDA(rt_detail::pool_alloc(device, (void**)&d_ids, mine.size() * sizeof(uint32_t), &got));
    cudaError_t e = cudaMemcpyAsync(d_ids, mine.data(), mine.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    int rc = e == cudaSuccess ? rt_unpack_impl(device, d_packed, d_ids, (uint32_t)mine.size(), tile, tx, width, height, d_rgb, st)
                              : rt_cuda_fail(e, "tile id upload");
    cudaStreamSynchronize(st);          // the id list goes back to the pool
    rt_detail::pool_free(device, d_ids, got);
    return rc;
}

int rt_render_multi(RtScene* s, const RtCamera* camera, const RtRenderParams* params, RtComm* comm, int root,
                    float* d_rgb, RtRenderStats* stats, float* assemble_ms, void* stream)
{
    return rt_render_multi_impl(s, camera, params, comm, root, d_rgb, stats, assemble_ms, static_cast<cudaStream_t>(stream));
}

int rt_generate_camera_rays(RtScene* s, const RtCamera* camera, const RtRenderParams* params, uint32_t psi, RtRay* rays)
{
    return rt_camera_rays_impl(s, camera, params, psi, rays);
}

int rt_stage1_render(int device, const RtStage1Plane* planes, uint32_t num_planes, const RtCamera* camera,
                     uint32_t width, uint32_t height, uint8_t* rgb8)
{
    return rt_stage1_impl(device, planes, num_planes, camera, width, height, rgb8);
}

int rt_stage1_render_float(int device, const RtStage1Plane* planes, uint32_t num_planes, const RtCamera* camera,
                            uint32_t width, uint32_t height, float* rgb, uint8_t* rgb8)
{
    if (rgb == NULL) return rt_fail(RT_ERR_ARG, "null argument");
    return rt_stage1_impl(device, planes, num_planes, camera, width, height, rgb8, rgb);
}

int rt_tonemap_bgra8_device(int device, const float* d_rgb, size_t num_pixels, float exposure_stops, float gamma,
                            uint8_t* d_bgra, void* stream)
{
    return rt_tonemap_device_impl(device, d_rgb, num_pixels, exposure_stops, gamma, d_bgra, static_cast<cudaStream_t>(stream));
}

void rt_release_cached_memory(void)
{
    rt_detail::s23_release();                      // the Stage 2/3 working set goes back to the pool first
    rt_detail::release_parked_render();            // ... and the wavefront state parked by the last scene destroyed
    rt_detail::pool_release_all();
    rt_detail::validate_scratch().release();       // the calling thread's BVH validation scratch
}

int rt_stage23_render(int device, const RtS23Scene* scene, const RtCamera* camera, const RtS23Params* params,
                      float* rgb, uint8_t* rgb8, RtRenderStats* stats)
{
    return rt_detail::rt_stage23_impl(device, scene, camera, params, rgb, rgb8, stats);
}

int rt_tile_owners(uint32_t width, uint32_t height, uint32_t tile_size, uint32_t world,
                   uint32_t* owners, uint32_t* tiles_x, uint32_t* tiles_y, uint32_t* tile_size_used)
{
    if (width == 0 || height == 0 || world == 0)
        return rt_fail(RT_ERR_ARG, "width, height and world must be positive");
    uint32_t tile = tile_size ? tile_size : RT_DEFAULT_TILE;
    uint32_t tx = (width + tile - 1) / tile, ty = (height + tile - 1) / tile;
    if (tiles_x) *tiles_x = tx;
    if (tiles_y) *tiles_y = ty;
    if (tile_size_used) *tile_size_used = tile;
    if (owners)
    {
        std::vector<uint32_t> mine;
        for (uint32_t r = 0; r < world; ++r)
        {
            rt_detail::rank_tiles(tx, ty, r, world, mine);
            for (size_t k = 0; k < mine.size(); ++k) owners[mine[k]] = r;
        }
    }
    return RT_OK;
}

int rt_sample_permutations(uint32_t width, uint32_t height, uint32_t depth, uint32_t x, uint32_t y, uint32_t* out)
{
    if (out == NULL || width == 0 || height == 0 || x >= width || y >= height || depth > RT_MAX_DEPTH)
        return rt_fail(RT_ERR_ARG, "bad argument");
    ChunkGrid g = chunk_grid(width, height);
    if (!chunk_covers(g, x, y))
        return rt_fail(RT_ERR_ARG, "pixel lies outside every reference chunk (image narrower than 4 pixels): it is never rendered");
    pixel_permutations(g, x, y, depth, out);
    return RT_OK;
}

float rt_cmj_sample1d(uint32_t index, uint32_t samples, uint32_t permutation)
{
    return cmj_sample1d(index, samples, permutation);
}

void rt_cmj_sample2d(uint32_t index, uint32_t x_samples, uint32_t y_samples, uint32_t permutation, float* u, float* v)
{
    cmj_sample2d(index, x_samples, y_samples, permutation, *u, *v);
}

int rt_libm_eval(int kind, const float* x, const float* y, size_t n, float* out)
{
    if (x == NULL || out == NULL || (kind == 2 && y == NULL) || kind < 0 || kind > 4)
        return rt_fail(RT_ERR_ARG, "bad argument");
    for (size_t i = 0; i < n; ++i)
    {
        if (kind >= 3)
        {
            float sn, cs;
            rtm_sincosf(x[i], sn, cs);
            out[i] = kind == 3 ? sn : cs;
        }
        else
            out[i] = kind == 0 ? rtm_sinf(x[i]) : kind == 1 ? rtm_cosf(x[i]) : rtm_powf(x[i], y[i]);
    }
    return RT_OK;
}

int rt_tonemap_bgra8(int device, const float* rgb, size_t num_pixels, float exposure_stops, float gamma, uint8_t* bgra)
{
    return rt_tonemap_impl(device, rgb, num_pixels, exposure_stops, gamma, bgra);
}

} // extern "C"
