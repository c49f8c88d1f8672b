// Host-side worker threads for the once-per-raytrace() preparation of large scenes
// (BVH build, mesh bounds, flattening).  The reference prepares on one thread
// (Rayito_Stage7_QT/RScene.h:186-205, RMesh.h:89-129, RAccel.h:262-374); at 10 M
// triangles that is seconds of wall time in front of a render that takes about as
// long on the GPU, so the host side is spread over the cores -- under one rule:
// every result is bit-identical to the serial order of the reference.  Work is cut
// into index-ordered chunks whose partial results are combined in chunk order, and
// the BVH is split into subtrees whose node numbers are known before they are built.
#ifndef RAYITO_B200_PARALLEL_HPP
#define RAYITO_B200_PARALLEL_HPP

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <system_error>
#include <thread>
#include <vector>

namespace rayito_b200
{

// Worker threads for host preparation: RAYITO_B200_HOST_THREADS, else the hardware
// concurrency, capped at 32.  1 reproduces the single-threaded code path exactly
// (same results either way; the tests compare them).
inline unsigned hostThreads()
{
    const char* env = std::getenv("RAYITO_B200_HOST_THREADS");
    if (env != NULL)
    {
        long n = std::strtol(env, NULL, 10);
        if (n >= 1)
            return n > 256 ? 256u : (unsigned)n;
    }
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    return n > 32 ? 32u : n;
}

// Number of index-ordered chunks parallelChunks() will use for n items
inline unsigned chunkCount(size_t n, size_t grain)
{
    unsigned threads = hostThreads();
    size_t byGrain = grain ? (n + grain - 1) / grain : 1;
    if (byGrain < 1) byGrain = 1;
    return (unsigned)(byGrain < threads ? byGrain : threads);
}

// Calls body(chunk, begin, end) for `chunks` contiguous index ranges covering [0, n),
// chunk c before chunk c+1 in index order, each on its own thread (chunk 0 on the caller).
template <typename Body>
inline void parallelChunks(size_t n, unsigned chunks, Body body)
{
    if (chunks <= 1 || n == 0)
    {
        body(0u, (size_t)0, n);
        return;
    }
    // An exception on a worker (std::bad_alloc, say) is carried back to the caller
    std::vector<std::exception_ptr> failed(chunks);
    std::exception_ptr* fail = &failed[0];
    std::vector<std::thread> workers;
    workers.reserve(chunks - 1);
    for (unsigned c = 1; c < chunks; ++c)
    {
        size_t b = n * c / chunks, e = n * (c + 1) / chunks;
        try
        {
            workers.push_back(std::thread([body, fail, c, b, e]() {
                try { body(c, b, e); } catch (...) { fail[c] = std::current_exception(); }
            }));
        }
        catch (const std::system_error&)
        {
            // no thread to be had: this chunk runs here
            try { body(c, b, e); } catch (...) { fail[c] = std::current_exception(); }
        }
    }
    try { body(0u, (size_t)0, n / chunks); } catch (...) { fail[0] = std::current_exception(); }
    for (size_t i = 0; i < workers.size(); ++i)
        workers[i].join();
    for (unsigned c = 0; c < chunks; ++c)
        if (failed[c])
            std::rethrow_exception(failed[c]);
}

// A bag of independent jobs processed by `threads` workers; a job may add further jobs.
// run(job, bag) is called once per job; the call returns when the bag has drained.
template <typename Job>
class JobBag
{
public:
    JobBag() : m_running(0) { }

    void add(const Job& job)
    {
        {
            std::lock_guard<std::mutex> lock(m_mutex);
            m_jobs.push_back(job);
        }
        m_wake.notify_one();
    }

    template <typename Run>
    void drain(unsigned threads, Run run)
    {
        if (threads <= 1)
        {
            work(run);
            rethrow();
            return;
        }
        std::vector<std::thread> workers;
        workers.reserve(threads - 1);
        for (unsigned t = 1; t < threads; ++t)
        {
            try { workers.push_back(std::thread([this, run]() { this->work(run); })); }
            catch (const std::system_error&) { break; }     // fewer workers; the caller drains the rest
        }
        work(run);
        for (size_t i = 0; i < workers.size(); ++i)
            workers[i].join();
        rethrow();
    }

private:
    template <typename Run>
    void work(Run run)
    {
        std::unique_lock<std::mutex> lock(m_mutex);
        for (;;)
        {
            while (m_jobs.empty() && m_running != 0)
                m_wake.wait(lock);
            if (m_jobs.empty())
                break;              // nothing queued, nobody running: drained
            Job job = m_jobs.back();
            m_jobs.pop_back();
            ++m_running;
            lock.unlock();
            std::exception_ptr failure;
            try { run(job, *this); } catch (...) { failure = std::current_exception(); }
            lock.lock();
            if (failure)
            {
                // first failure wins; the queued jobs are dropped so that everybody drains
                if (!m_failure) m_failure = failure;
                m_jobs.clear();
            }
            if (--m_running == 0 && m_jobs.empty())
                m_wake.notify_all();
        }
    }

    void rethrow()
    {
        if (m_failure)
        {
            std::exception_ptr failure = m_failure;
            m_failure = std::exception_ptr();
            std::rethrow_exception(failure);
        }
    }

    std::mutex m_mutex;
    std::condition_variable m_wake;
    std::vector<Job> m_jobs;
    unsigned m_running;
    std::exception_ptr m_failure;
};

// Per-thread scratch block that survives between calls: a 10 M-triangle prepare() needs
// 140 MB of build items, and handing that back to the OS after every raytrace() means
// ~35 000 page faults to get it again on the next one.  releaseHostCaches() frees it.
class ScratchBlock
{
public:
    ScratchBlock() : m_data(NULL), m_bytes(0) { }
    ~ScratchBlock() { std::free(m_data); }
    void* get(size_t bytes)
    {
        if (bytes > m_bytes)
        {
            std::free(m_data);
            m_data = std::malloc(bytes);
            m_bytes = m_data ? bytes : 0;
        }
        return m_data;
    }
    void release() { std::free(m_data); m_data = NULL; m_bytes = 0; }
private:
    ScratchBlock(const ScratchBlock&);
    ScratchBlock& operator=(const ScratchBlock&);
    void* m_data;
    size_t m_bytes;
};

inline ScratchBlock& buildScratch()
{
    static thread_local ScratchBlock block;
    return block;
}

} // namespace rayito_b200

#endif // RAYITO_B200_PARALLEL_HPP
