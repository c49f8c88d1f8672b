"""Seeded ray batches shared by the parity tests."""
import numpy as np

from rayito_b200.capi import RAY_DTYPE


def random_rays(n, seed, center=(0.0, 0.0, 0.0), radius=9.0, target_radius=2.5, shadow_fraction=0.0):
    """Origins on a sphere around the scene, directions through a ball near its
    centre, times uniform in [0,1].  A fraction gets a finite tMax (shadow-like)."""
    rng = np.random.RandomState(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    origins = np.asarray(center, np.float64) + radius * d
    t = rng.normal(size=(n, 3))
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    targets = np.asarray(center, np.float64) + target_radius * rng.uniform(size=(n, 1)) ** (1 / 3) * t
    dirs = targets - origins
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rays = np.zeros(n, RAY_DTYPE)
    rays["origin"] = origins.astype(np.float32)
    rays["direction"] = dirs.astype(np.float32)
    rays["tmax"] = np.float32(1.0e30)
    rays["time"] = rng.uniform(size=n).astype(np.float32)
    k = int(n * shadow_fraction)
    if k:
        rays["tmax"][:k] = rng.uniform(2.0, 14.0, size=k).astype(np.float32)
    return rays


def axis_parallel_rays(n, seed):
    """Rays with exact zero (and negative-zero) direction components, some starting
    exactly on coordinates that bound boxes in the scenes: they exercise 0*inf = NaN
    in the slab test and the -0.0 -> +0.0 canonicalisation of identity transforms."""
    rng = np.random.RandomState(seed)
    rays = np.zeros(n, RAY_DTYPE)
    axes = rng.randint(0, 3, size=n)
    signs = rng.choice([-1.0, 1.0], size=n)
    zeros = rng.choice([0.0, -0.0], size=(n, 3))
    dirs = zeros.astype(np.float32)
    dirs[np.arange(n), axes] = signs
    grid = np.array([-3.0, -2.0, -1.5, -1.0, -0.5, 0.0, 0.2, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0], np.float32)
    origins = grid[rng.randint(0, len(grid), size=(n, 3))].astype(np.float32)
    jitter = rng.uniform(-3, 3, size=(n, 3)).astype(np.float32)
    use_grid = rng.uniform(size=(n, 3)) < 0.6
    origins = np.where(use_grid, origins, jitter).astype(np.float32)
    origins[np.arange(n), axes] = (-8.0 * signs).astype(np.float32)
    rays["origin"] = origins
    rays["direction"] = dirs
    rays["tmax"] = np.float32(1.0e30)
    rays["time"] = rng.choice([0.0, 0.25, 0.5, 1.0, 0.33, 0.67], size=n).astype(np.float32)
    return rays


def deep_scene_rays(n, seed, shadow_fraction=0.25):
    """Rays for the wedge-mesh scenes: a quarter aimed at the whole wedge, a quarter at its tip
    (where the face BVH is deepest), a quarter at the sphere chain / lights, and a quarter GRAZING
    the wedge from its tip outwards at shutter time 0 (no rotation yet): such a ray pierces the
    box of every row, nearest rows last, so the far children pile up on the traversal stack --
    dozens of live entries, which is what the deep-stack kernels exist for."""
    q = n // 4
    tip = (-0.5, -0.5, 0.0)
    a = random_rays(q, seed, center=(0.3, -0.5, 0.0), radius=7.0, target_radius=1.6, shadow_fraction=shadow_fraction)
    b = random_rays(q, seed + 1, center=tip, radius=5.0, target_radius=0.03, shadow_fraction=shadow_fraction)
    c = random_rays(q, seed + 2, center=(0.0, 0.0, 0.0), radius=9.0, target_radius=4.0,
                    shadow_fraction=shadow_fraction)
    m = n - 3 * q
    rng = np.random.RandomState(seed + 3)
    g = np.zeros(m, RAY_DTYPE)
    spread = 10.0 ** rng.uniform(-14.0, -0.7, size=(m, 1))
    offs = rng.normal(size=(m, 3)) * spread * 0.2
    offs[rng.uniform(size=m) < 0.6, 1:] = 0.0         # origin exactly on the wedge's axis: only the slope leaves it
    offs[:, 0] = -rng.uniform(0.0, 0.3, size=m)
    g["origin"] = (np.asarray(tip) + offs).astype(np.float32)
    d = np.concatenate([np.ones((m, 1)), rng.normal(size=(m, 2)) * spread], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    g["direction"] = d.astype(np.float32)
    g["tmax"] = np.float32(1.0e30)
    g["time"] = np.where(rng.uniform(size=m) < 0.7, 0.0, rng.uniform(size=m)).astype(np.float32)
    # ... and a fifth of them along the chain of halving spheres of the "deep both" scene, from its
    # small end outwards (same effect on the TOP-level stack)
    chain = rng.uniform(size=m) < 0.2
    axis = np.array([6.0, 1.5, 0.0]) / np.linalg.norm([6.0, 1.5, 0.0])
    cd = axis + rng.normal(size=(m, 3)) * spread * 0.3
    cd /= np.linalg.norm(cd, axis=1, keepdims=True)
    co = np.array([-3.0, -1.0, -1.0]) - rng.uniform(0.0, 0.5, size=(m, 1)) * axis
    g["origin"][chain] = co[chain].astype(np.float32)
    g["direction"][chain] = cd[chain].astype(np.float32)
    k = int(m * shadow_fraction)
    g["tmax"][:k] = rng.uniform(0.5, 4.0, size=k).astype(np.float32)
    return np.concatenate([a, b, c, g])


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)
