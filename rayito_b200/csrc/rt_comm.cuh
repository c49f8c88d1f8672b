// Multi-GPU tile assembly: the one collective of the render path.
//
// The reference shards a frame over 16 CPU threads by image chunks that all write into one
// Image (Rayito_Stage7_QT/RaytraceMain.cpp:504-568).  Here a frame is sharded over GPUs by
// screen tiles (rt_tile_owners) and there is no shared Image: every rank renders ITS tiles into
// a packed buffer (tile after tile, no holes), the packed buffers travel to the root rank with
// one grouped ncclSend / ncclRecv over NVLink, and one kernel on the root scatters them into the
// frame.  Nothing else crosses GPUs: 12 bytes per pixel in total, each pixel exactly once (a
// sum-reduce of mostly-zero frames would move world x that).
//
// NCCL is resolved at run time (dlopen): the library loads and renders on one GPU without it,
// and inside a process that already holds a libnccl (PyTorch) the same copy is used.
#ifndef RAYITO_B200_RT_COMM_CUH
#define RAYITO_B200_RT_COMM_CUH

#include <dlfcn.h>
#include <nccl.h>

#include "rt_render.cuh"

namespace rt_detail
{

struct NcclApi
{
    void* handle;
    std::string why;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int*);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int*);
};

inline NcclApi* nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        std::memset(static_cast<void*>(&api.handle), 0, sizeof(api.handle));
        const char* env = std::getenv("RAYITO_B200_NCCL_LIB");
        // a copy the process already holds (PyTorch's) first, so that one NCCL serves everybody
        void* h = env ? dlopen(env, RTLD_NOW | RTLD_LOCAL) : dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
        if (h == NULL && env == NULL) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (h == NULL && env == NULL) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (h == NULL)
        {
            const char* e = dlerror();
            api.why = std::string("NCCL library not found (libnccl.so.2; set RAYITO_B200_NCCL_LIB): ") + (e ? e : "");
            return;
        }
        bool ok = true;
#define RT_NCCL_SYM(field, name)                                                   \
        *reinterpret_cast<void**>(&api.field) = dlsym(h, name);                    \
        if (api.field == NULL) { ok = false; api.why = std::string("NCCL symbol missing: ") + name; }
        RT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        RT_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        RT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        RT_NCCL_SYM(CommCount, "ncclCommCount")
        RT_NCCL_SYM(CommUserRank, "ncclCommUserRank")
        RT_NCCL_SYM(GroupStart, "ncclGroupStart")
        RT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        RT_NCCL_SYM(Send, "ncclSend")
        RT_NCCL_SYM(Recv, "ncclRecv")
        RT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        RT_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef RT_NCCL_SYM
        if (ok) api.handle = h;
    });
    return &api;
}

inline size_t packed_floats(uint32_t width, uint32_t height, uint32_t tile_size, uint32_t world, uint32_t rank)
{
    uint32_t tile = tile_size ? tile_size : RT_DEFAULT_TILE;
    uint32_t tx = (width + tile - 1) / tile, ty = (height + tile - 1) / tile;
    std::vector<uint32_t> mine;
    rank_tiles(tx, ty, rank, world, mine);
    return mine.size() * (size_t)tile * tile * 3;
}

} // namespace rt_detail

#define RT_NCCL(call)                                                                       \
    do {                                                                                    \
        ncclResult_t r__ = (call);                                                          \
        if (r__ != ncclSuccess)                                                             \
            return rt_fail(RT_ERR_COMM, std::string("NCCL error at " #call ": ") + rt_detail::nccl_api()->GetErrorString(r__)); \
    } while (0)

struct RtComm
{
    ncclComm_t comm;
    bool owned;                 // created by rt_comm_create (destroyed with the handle) or borrowed from the application
    int rank, world, device;
    float* d_packed;            // this rank's packed tiles (non-root) / everybody else's (root)
    size_t packed_bytes;        // real size of the pool block
    uint32_t* d_tiles;          // root: tile ids in packed order, all other ranks one after the other
    size_t tiles_bytes;
    float* d_frame;             // root, host-output entry point: the assembled frame before its download
    size_t frame_bytes;
    cudaEvent_t ev[3];
};

// Scatter packed tiles into the frame.  One thread per packed pixel.
__global__ void __launch_bounds__(256)
k_unpack_tiles(const float* __restrict__ packed, const uint32_t* __restrict__ tile_ids, uint32_t num_tiles,
               uint32_t tile, uint32_t tiles_x, uint32_t width, uint32_t height, float* __restrict__ rgb)
{
    const uint32_t per_tile = tile * tile;
    const size_t n = (size_t)num_tiles * per_tile;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        uint32_t slot = (uint32_t)(i / per_tile), off = (uint32_t)(i % per_tile);
        uint32_t id = tile_ids[slot];
        uint32_t x = (id % tiles_x) * tile + off % tile;
        uint32_t y = (id / tiles_x) * tile + off / tile;
        if (x >= width || y >= height)
            continue;
        const float* src = packed + i * 3;
        float* dst = rgb + ((size_t)y * width + x) * 3;
        dst[0] = src[0];
        dst[1] = src[1];
        dst[2] = src[2];
    }
}

inline int rt_unpack_impl(int device, const float* d_packed, const uint32_t* d_tile_ids, uint32_t num_tiles, uint32_t tile,
                          uint32_t tiles_x, uint32_t width, uint32_t height, float* d_rgb, cudaStream_t st)
{
    if (num_tiles == 0)
        return RT_OK;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    size_t n = (size_t)num_tiles * tile * tile;
    unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, (size_t)sms * 16);
    k_unpack_tiles<<<blocks, 256, 0, st>>>(d_packed, d_tile_ids, num_tiles, tile, tiles_x, width, height, d_rgb);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

inline int rt_comm_wrap(ncclComm_t comm, bool owned, int rank, int world, int device, RtComm** out)
{
    RtComm* c = new RtComm();
    c->comm = comm;
    c->owned = owned;
    c->rank = rank;
    c->world = world;
    c->device = device;
    c->d_packed = NULL;
    c->packed_bytes = 0;
    c->d_tiles = NULL;
    c->tiles_bytes = 0;
    c->d_frame = NULL;
    c->frame_bytes = 0;
    for (int i = 0; i < 3; ++i) cudaEventCreate(&c->ev[i]);
    *out = c;
    return RT_OK;
}

// rt_render_multi: see include/rayito_b200.h
inline int rt_render_multi_impl(RtScene* s, const RtCamera* camera, const RtRenderParams* prm, RtComm* comm, int root,
                                float* d_rgb, RtRenderStats* stats, float* assemble_ms, cudaStream_t st)
{
    if (s == NULL || camera == NULL || prm == NULL || comm == NULL)
        return rt_fail(RT_ERR_ARG, "null argument");
    if ((int)prm->world != comm->world || (int)prm->rank != comm->rank)
        return rt_fail(RT_ERR_ARG, "RtRenderParams.rank / world do not match the communicator");
    if (root < 0 || root >= comm->world)
        return rt_fail(RT_ERR_ARG, "root rank out of range");
    if (comm->device != s->device)
        return rt_fail(RT_ERR_ARG, "scene and communicator live on different devices");
    const bool is_root = comm->rank == root;
    if (is_root && d_rgb == NULL)
        return rt_fail(RT_ERR_ARG, "the root rank needs a device frame buffer");
    RT_CUDA(cudaSetDevice(s->device));
    rt_detail::NcclApi* nccl = rt_detail::nccl_api();
    if (comm->world > 1 && nccl->handle == NULL)
        return rt_fail(RT_ERR_COMM, nccl->why);

    const uint32_t tile = prm->tile_size ? prm->tile_size : RT_DEFAULT_TILE;
    const uint32_t tiles_x = (prm->width + tile - 1) / tile, tiles_y = (prm->height + tile - 1) / tile;
    const size_t per_tile = (size_t)tile * tile * 3;
    // packed layout of every rank (tile ids ascending per rank); the root lists all the others back to back
    std::vector<std::vector<uint32_t> > tiles((size_t)comm->world);
    for (int r = 0; r < comm->world; ++r)
        rt_detail::rank_tiles(tiles_x, tiles_y, (uint32_t)r, (uint32_t)comm->world, tiles[(size_t)r]);
    size_t need_floats = 0;
    std::vector<uint32_t> others;
    if (is_root)
    {
        for (int r = 0; r < comm->world; ++r)
            if (r != root) others.insert(others.end(), tiles[(size_t)r].begin(), tiles[(size_t)r].end());
        need_floats = others.size() * per_tile;
    }
    else
        need_floats = tiles[(size_t)comm->rank].size() * per_tile;
    if (need_floats * sizeof(float) > comm->packed_bytes)
    {
        rt_detail::pool_free(comm->device, comm->d_packed, comm->packed_bytes);
        comm->d_packed = NULL;
        comm->packed_bytes = 0;
        RT_CUDA(rt_detail::pool_alloc(comm->device, (void**)&comm->d_packed, need_floats * sizeof(float), &comm->packed_bytes));
    }
    if (is_root && others.size() * sizeof(uint32_t) > comm->tiles_bytes)
    {
        rt_detail::pool_free(comm->device, comm->d_tiles, comm->tiles_bytes);
        comm->d_tiles = NULL;
        comm->tiles_bytes = 0;
        RT_CUDA(rt_detail::pool_alloc(comm->device, (void**)&comm->d_tiles, others.size() * sizeof(uint32_t), &comm->tiles_bytes));
    }
    if (is_root && !others.empty())
        RT_CUDA(cudaMemcpyAsync(comm->d_tiles, others.data(), others.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));

    // 1. render: the root straight into the frame, everybody else into its packed buffer
    int rc;
    if (is_root)
        rc = rt_render_impl(s, camera, prm, d_rgb, true, stats, st);
    else
        rc = rt_render_impl(s, camera, prm, comm->d_packed, true, stats, st, need_floats ? need_floats : 1);
    if (rc != RT_OK)
        return rc;

    // 2. the collective: packed tiles to the root, one grouped send / receive
    RT_CUDA(cudaEventRecord(comm->ev[0], st));
    if (comm->world > 1)
    {
        RT_NCCL(nccl->GroupStart());
        if (is_root)
        {
            size_t off = 0;
            for (int r = 0; r < comm->world; ++r)
            {
                if (r == root) continue;
                size_t cnt = tiles[(size_t)r].size() * per_tile;
                if (cnt) RT_NCCL(nccl->Recv(comm->d_packed + off, cnt, ncclFloat32, r, comm->comm, st));
                off += cnt;
            }
        }
        else if (need_floats)
            RT_NCCL(nccl->Send(comm->d_packed, need_floats, ncclFloat32, root, comm->comm, st));
        RT_NCCL(nccl->GroupEnd());
    }
    RT_CUDA(cudaEventRecord(comm->ev[1], st));
    // 3. the root scatters the received tiles into the frame
    if (is_root)
    {
        rc = rt_unpack_impl(comm->device, comm->d_packed, comm->d_tiles, (uint32_t)others.size(), tile, tiles_x,
                            prm->width, prm->height, d_rgb, st);
        if (rc != RT_OK) return rc;
    }
    RT_CUDA(cudaEventRecord(comm->ev[2], st));
    RT_CUDA(cudaStreamSynchronize(st));
    if (assemble_ms)
        cudaEventElapsedTime(assemble_ms, comm->ev[0], comm->ev[2]);
    return RT_OK;
}

// The same with the assembled frame delivered to HOST memory on the root (what raytrace() returns)
inline int rt_render_multi_host_impl(RtScene* s, const RtCamera* camera, const RtRenderParams* prm, RtComm* comm, int root,
                                     float* rgb_host, RtRenderStats* stats, float* assemble_ms)
{
    if (comm == NULL || prm == NULL)
        return rt_fail(RT_ERR_ARG, "null argument");
    const bool is_root = comm->rank == root;
    if (is_root && rgb_host == NULL)
        return rt_fail(RT_ERR_ARG, "the root rank needs a host frame buffer");
    const size_t bytes = (size_t)prm->width * prm->height * 3 * sizeof(float);
    if (is_root && bytes > comm->frame_bytes)
    {
        RT_CUDA(cudaSetDevice(comm->device));
        rt_detail::pool_free(comm->device, comm->d_frame, comm->frame_bytes);
        comm->d_frame = NULL;
        comm->frame_bytes = 0;
        RT_CUDA(rt_detail::pool_alloc(comm->device, (void**)&comm->d_frame, bytes, &comm->frame_bytes));
    }
    int rc = rt_render_multi_impl(s, camera, prm, comm, root, is_root ? comm->d_frame : NULL, stats, assemble_ms, 0);
    if (rc != RT_OK)
        return rc;
    if (is_root)
    {
        cudaEvent_t e0 = comm->ev[0], e1 = comm->ev[1];
        RT_CUDA(cudaEventRecord(e0, 0));
        RT_CUDA(cudaMemcpy(rgb_host, comm->d_frame, bytes, cudaMemcpyDeviceToHost));
        RT_CUDA(cudaEventRecord(e1, 0));
        RT_CUDA(cudaEventSynchronize(e1));
        if (stats) cudaEventElapsedTime(&stats->download_ms, e0, e1);
    }
    return RT_OK;
}

#endif // RAYITO_B200_RT_COMM_CUH
