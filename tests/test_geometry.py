"""mpifft4py_b200/_geometry.py: the pieces every shape / slice / mesh of the three classes is assembled from.
(The assembled results are compared with the unmodified reference in tests/test_host_api.py.)"""
import numpy as np

from mpifft4py_b200 import _geometry as G


def test_blocks_tile_an_axis_and_stretch_with_the_pad_factor():
    n, parts = 48, 4
    w = n // parts
    got = [G.block(w, i) for i in range(parts)]
    assert [s.start for s in got] == [0, 12, 24, 36] and got[-1].stop == n and all(s.step == 1 for s in got)
    pad = [G.block(w, i, 1.5) for i in range(parts)]
    assert pad[0] == slice(0, 18, 1) and pad[-1].stop == int(1.5 * n)
    assert [p.stop for p in pad[:-1]] == [p.start for p in pad[1:]]  # padded blocks tile the padded axis
    assert G.whole(10, 1.5) == slice(0, 15, 1) and G.whole(10, 1, None) == slice(0, 10)
    assert G.padded((4, 6, 8), 1.5) == (6, 9, 12)


def test_frequency_vectors_and_nyquist_removal():
    assert np.array_equal(G.frequencies(8), [0, 1, 2, 3, -4, -3, -2, -1])
    assert np.array_equal(G.frequencies(8, half=True), [0, 1, 2, 3, 4])
    ks = [G.frequencies(8), G.frequencies(6), G.frequencies(5)]
    G.drop_nyquist(ks, (8, 6, 5))
    assert ks[0][4] == 0 and ks[1][3] == 0 and np.array_equal(ks[2], np.fft.fftfreq(5, 1. / 5))  # odd axis untouched


def test_physical_meshes_sparse_and_dense_agree_up_to_rounding():
    N, L = np.array([8, 4, 6]), np.array([2 * np.pi, 1.0, 3.0])
    sl = (G.block(4, 1), G.whole(4), G.whole(6))
    sparse = G.sparse_physical_mesh(sl, N, L, np.float64, (4, 4, 6))
    dense = G.dense_physical_mesh(sl, N, L, np.float64)
    assert dense.shape == (3, 4, 4, 6) and all(x.shape == (4, 4, 6) for x in sparse)
    assert all(0 in x.strides for x in sparse)  # broadcast views, not copies
    for i in range(3):
        assert np.allclose(sparse[i], dense[i], rtol=1e-15, atol=1e-15)
    assert sparse[0][0, 0, 0] == 4 * L[0] / N[0]


def test_mask_keeps_the_two_thirds_band():
    N = np.array([12, 12, 12])
    K = G.sparse_spectral_mesh([G.frequencies(12), G.frequencies(12), G.frequencies(12, half=True)])
    m = G.two_thirds_mask(K, N)
    assert m.dtype == np.uint8 and m.shape == (12, 12, 7)
    kmax = 2. / 3. * (12 // 2 + 1)
    assert m[4, 0, 0] == 1 and m[5, 0, 0] == 0 and 4 < kmax < 5
    assert m[-4, -4, 4] == 1 and m[0, 0, 5] == 0
    assert int(m.sum()) == 9 * 9 * 5


def test_balanced_grid_is_compute_dims():
    assert [G.balanced_grid(P) for P in (1, 2, 4, 8, 16, 32, 12)] == [(1, 1), (2, 1), (2, 2), (4, 2), (4, 4), (8, 4), (4, 3)]
