// Device-memory pool behind the scene / render entry points.
//
// Rayito::raytrace() builds a fresh device scene per call (as the reference rebuilds its
// scene per render), and cudaMalloc / cudaFree of the scene arena, the image and the ~11 GB
// wavefront state showed up as 50-430 ms of "destroy" time in the end-to-end numbers
// (RAYITO_B200_TIMING=1).  Freed blocks are parked here per device and handed back to the
// next request of a similar size, so steady-state calls allocate nothing.
// rt_release_cached_memory() (or process exit) returns everything to the driver.
#ifndef RAYITO_B200_RT_POOL_CUH
#define RAYITO_B200_RT_POOL_CUH

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace rt_detail
{

struct PoolEntry
{
    void* ptr;
    size_t bytes;
    int device;
};

inline std::mutex& pool_lock() { static std::mutex m; return m; }
inline std::vector<PoolEntry>& pool_entries() { static std::vector<PoolEntry> v; return v; }

// cudaMalloc, or a parked block of at least `bytes` (and not absurdly larger).  *got receives the
// real size of the block, which must be passed back to pool_free.
inline cudaError_t pool_alloc(int device, void** out, size_t bytes, size_t* got)
{
    if (bytes == 0)
        bytes = 16;
    {
        std::lock_guard<std::mutex> guard(pool_lock());
        std::vector<PoolEntry>& v = pool_entries();
        size_t best = v.size();
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].device == device && v[i].bytes >= bytes && v[i].bytes <= 2 * bytes + (1u << 20) &&
                (best == v.size() || v[i].bytes < v[best].bytes))
                best = i;
        if (best != v.size())
        {
            *out = v[best].ptr;
            *got = v[best].bytes;
            v.erase(v.begin() + best);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess)
    {
        // make room: give the parked blocks of this device back and try once more
        cudaGetLastError();
        std::vector<void*> victims;
        {
            std::lock_guard<std::mutex> guard(pool_lock());
            std::vector<PoolEntry>& v = pool_entries();
            for (size_t i = v.size(); i > 0; --i)
                if (v[i - 1].device == device)
                {
                    victims.push_back(v[i - 1].ptr);
                    v.erase(v.begin() + (i - 1));
                }
        }
        for (size_t i = 0; i < victims.size(); ++i) cudaFree(victims[i]);
        e = cudaMalloc(out, bytes);
    }
    *got = bytes;
    return e;
}

inline void pool_free(int device, void* ptr, size_t bytes)
{
    if (ptr == NULL)
        return;
    void* evict = NULL;
    {
        std::lock_guard<std::mutex> guard(pool_lock());
        std::vector<PoolEntry>& v = pool_entries();
        PoolEntry e = { ptr, bytes, device };
        v.push_back(e);
        size_t count = 0, smallest = v.size();
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].device == device)
            {
                ++count;
                if (smallest == v.size() || v[i].bytes < v[smallest].bytes) smallest = i;
            }
        if (count > 24)
        {
            evict = v[smallest].ptr;
            v.erase(v.begin() + smallest);
        }
    }
    if (evict) cudaFree(evict);
}

inline size_t pool_parked_bytes(int device)
{
    std::lock_guard<std::mutex> guard(pool_lock());
    size_t total = 0;
    const std::vector<PoolEntry>& v = pool_entries();
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i].device == device) total += v[i].bytes;
    return total;
}

// Pinned host staging block for the scene upload, cached between rt_scene_create() calls:
// the flattened scene is converted straight into it by the host worker threads and leaves
// with one cudaMemcpy at pinned-memory speed.  A pageable block would cost a page fault per
// 4 KB on every call (660 MB at 10 M triangles) and a slower, driver-staged copy.
// Hold stage_lock() from stage_acquire() until the copy has finished.
struct StageBlock
{
    void* ptr;
    size_t bytes;
    bool pinned;
};
inline std::mutex& stage_lock() { static std::mutex m; return m; }
inline StageBlock& stage_block() { static StageBlock b = { NULL, 0, false }; return b; }

inline void stage_release_locked()
{
    StageBlock& b = stage_block();
    if (b.ptr != NULL)
    {
        if (b.pinned) cudaFreeHost(b.ptr); else std::free(b.ptr);
    }
    b.ptr = NULL; b.bytes = 0; b.pinned = false;
}

// At least `bytes` of host memory (pinned when the driver grants it); NULL when out of memory
inline void* stage_acquire_locked(size_t bytes)
{
    StageBlock& b = stage_block();
    if (b.ptr != NULL && b.bytes >= bytes && b.bytes <= 4 * bytes + ((size_t)64 << 20))
        return b.ptr;
    stage_release_locked();
    size_t want = (bytes + ((size_t)4 << 20) - 1) & ~(((size_t)4 << 20) - 1);
    void* p = NULL;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) == cudaSuccess)
        b.pinned = true;
    else
    {
        cudaGetLastError();
        p = std::malloc(want);
        b.pinned = false;
    }
    b.ptr = p;
    b.bytes = p ? want : 0;
    return p;
}

inline void pool_release_all()
{
    {
        std::lock_guard<std::mutex> guard(stage_lock());
        stage_release_locked();
    }
    std::vector<PoolEntry> all;
    {
        std::lock_guard<std::mutex> guard(pool_lock());
        all.swap(pool_entries());
    }
    int current = 0;
    cudaGetDevice(&current);
    for (size_t i = 0; i < all.size(); ++i)
    {
        cudaSetDevice(all[i].device);
        cudaFree(all[i].ptr);
    }
    cudaSetDevice(current);
}

} // namespace rt_detail

#endif // RAYITO_B200_RT_POOL_CUH
