"""Shared helpers for the GPU parity tests (test infrastructure)."""
import numpy as np

from oracle import oracle as vo


def latent_data(n, dim, latent=8, noise=0.2, seed=0, normalize=False):
    rng = np.random.default_rng(seed)
    z = rng.normal(size=(n, latent)).astype(np.float32) @ rng.normal(size=(latent, dim)).astype(np.float32)
    x = (z + noise * rng.normal(size=(n, dim))).astype(np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x)


def queries_near(x, nq, jitter=0.1, seed=1):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, x.shape[0], nq)
    return np.ascontiguousarray((x[idx] + jitter * rng.normal(size=(nq, x.shape[1]))).astype(np.float32))


def build_oracle(metric, x, M=16, ef_c=100):
    g = vo.Hnsw(metric, x.shape[1], M=M, ef_construction=ef_c)
    g.insert_many(x)
    return g


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32),
                          np.ascontiguousarray(b, np.float32).view(np.uint32))
