// parallelPartition must reproduce std::partition's element order exactly (the BVH build's
// "empty side => cut the range in half" fallback depends on it, RAccel.h:337-352).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rayito_b200/parallel.hpp"

struct Item { unsigned id; float key; };
struct Above
{
    float where;
    bool operator()(const Item& it) const { return where < it.key; }
};

static unsigned rng_state = 12345u;
static unsigned rng() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main()
{
    const size_t sizes[] = { 0, 1, 2, 3, 17, 1000, 4097, 65536, 300001, 1500000 };
    const float cuts[] = { -1.0f, 0.0f, 0.1f, 0.5f, 0.999f, 2.0f };      // -1: all true, 2: all false
    int checked = 0;
    for (size_t si = 0; si < sizeof(sizes) / sizeof(sizes[0]); ++si)
        for (size_t ci = 0; ci < sizeof(cuts) / sizeof(cuts[0]); ++ci)
            for (unsigned chunks = 1; chunks <= 13; chunks += 3)
            {
                const size_t n = sizes[si];
                std::vector<Item> a(n), b;
                for (size_t i = 0; i < n; ++i)
                {
                    a[i].id = (unsigned)i;
                    // many equal keys, runs of equal outcomes, and some noise
                    a[i].key = (rng() % 7 == 0) ? cuts[ci] : (float)(rng() % 1000) / 1000.0f;
                }
                b = a;
                Above pred = { cuts[ci] };
                Item* base = a.empty() ? NULL : &a[0];
                Item* want = std::partition(base, base + n, pred);
                std::vector<unsigned> scratch(n + 1);
                Item* bbase = b.empty() ? NULL : &b[0];
                Item* got = rayito_b200::parallelPartition(bbase, n, pred, &scratch[0], chunks);
                if ((want - base) != (got - bbase))
                {
                    std::printf("FAIL split n=%zu cut=%g chunks=%u: %td vs %td\n", n, cuts[ci], chunks, want - base, got - bbase);
                    return 1;
                }
                for (size_t i = 0; i < n; ++i)
                    if (a[i].id != b[i].id)
                    {
                        std::printf("FAIL order n=%zu cut=%g chunks=%u at %zu\n", n, cuts[ci], chunks, i);
                        return 1;
                    }
                ++checked;
            }
    std::printf("parallel_partition_test: %d cases identical to std::partition\n", checked);
    return 0;
}
