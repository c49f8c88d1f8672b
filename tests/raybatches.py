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


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)
