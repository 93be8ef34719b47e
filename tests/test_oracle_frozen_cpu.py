"""The oracle's search-only forms used by the large-config CPU arms (CSR adjacency, f16 / packed-bit storage) must
answer exactly like its reference-shaped form (per-node lists, f32 lanes): SURVEY finding 0.5 defines parity for the
f16 and packed-Hamming configs that way."""
import numpy as np

from oracle import oracle as vo


def _same(a, b):
    return all(np.array_equal(u, v) for u, v in zip(a, b))


def test_frozen_f32_f16_and_packed_bits_equal_the_reference_form(tmp_path):
    rng = np.random.default_rng(0)
    n, dim = 2500, 48
    x = rng.normal(size=(n, dim)).astype(np.float32)
    g = vo.Hnsw(vo.COSINE, dim, M=8, ef_construction=50)
    g.insert_many(x)
    L = g.export_graph()
    q = rng.normal(size=(40, dim)).astype(np.float32)
    want = g.search_batch(q, 10, 64)
    assert _same(want, vo.frozen(vo.COSINE, x, L, g.M, g.M0, g.entry_point, g.max_layer).search_batch(q, 10, 64, threads=3))
    # f16 storage == the f32 form over the rounded values (half::f16 round-to-nearest-even, half_precision.rs:97)
    xh = x.astype(np.float16)
    want16 = vo.Hnsw.from_arrays(vo.COSINE, xh.astype(np.float32), L, g.M, g.M0, g.entry_point, g.max_layer).search_batch(q, 10, 64)
    assert _same(want16, vo.frozen(vo.COSINE, xh, L, g.M, g.M0, g.entry_point, g.max_layer).search_batch(q, 10, 64))
    # files: format v1 graph, vectors from the file or borrowed
    g.dump(str(tmp_path))
    assert _same(want, vo.open_index(str(tmp_path), vo.COSINE).search_batch(q, 10, 64, threads=2))
    assert _same(want16, vo.open_index(str(tmp_path), vo.COSINE, vectors=xh).search_batch(q, 10, 64))
    # packed bits + popcount (simd_explicit.rs:308-360) == f32 lanes thresholded at > 0.5 (simd_explicit.rs:256-287)
    xb = (rng.random((n, 128)) > 0.5).astype(np.float32)
    gb = vo.Hnsw(vo.HAMMING, 128, M=8, ef_construction=50)
    gb.insert_many(xb)
    qb = (rng.random((40, 128)) > 0.5).astype(np.float32)
    packed = np.packbits(xb.astype(np.uint8), axis=1, bitorder="little").view(np.uint64)
    fb = vo.frozen(vo.HAMMING, packed, gb.export_graph(), gb.M, gb.M0, gb.entry_point, gb.max_layer)
    assert _same(gb.search_batch(qb, 10, 64), fb.search_batch(qb, 10, 64, threads=4))
