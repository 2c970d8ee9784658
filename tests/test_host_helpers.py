"""Host-side logic of the patch / reconstruction helpers (no GPU): grid shapes, patch-corner order, RNG draw order, CSR
conversion, the rounding rule of the network reconstruction -- against the reference's own expressions, sklearn and the
oracle's numpy restatements."""
import numpy as np
import pytest

from onmf_ontf_ndl_b200 import _host, patches, reconstruct
from oracle import onmf_oracle as O


@pytest.mark.parametrize("H,W,k,res", [(36, 40, 5, 2), (36, 40, 5, 1), (10, 10, 10, 1), (11, 10, 10, 3), (512, 512, 10, 4)])
def test_grid_shape_is_the_reference_loop(H, W, k, res):
    # image_reconstruction.py:375-376: for i in range(0, H - k, res): for j in range(0, W - k, res)
    ny, nx = reconstruct.grid_shape(H, W, k, res)
    assert (ny, nx) == (len(range(0, H - k, res)), len(range(0, W - k, res)))
    co = patches.all_patch_coords((H, W), k, stride=res, include_last=False)
    want = [(i, j) for i in range(0, H - k, res) for j in range(0, W - k, res)]
    assert co.tolist() == [list(c) for c in want] and co.dtype == np.int32


@pytest.mark.parametrize("shape,k", [((17, 23), 4), ((9, 9), 9), ((12, 7, 3), 5)])
def test_all_patch_coords_is_sklearn_order(shape, k):
    from sklearn.feature_extraction.image import extract_patches_2d
    rng = np.random.default_rng(0)
    img = rng.random(shape)
    ref = extract_patches_2d(img, (k, k))
    co = patches.all_patch_coords(shape, k)
    assert len(co) == len(ref)
    if img.ndim == 2:
        X = O.gather_patches_gray(img, co, k)
    else:
        X = O.gather_patches_color_tensor(img, co, k).reshape(k * k * shape[2], -1)
    assert np.array_equal(X, ref.reshape(len(ref), -1).T)
    with pytest.raises(ValueError):
        patches.all_patch_coords((3, 9), 4)


def test_sample_patch_coords_draw_order():
    # image_reconstruction.py:184-186: a = np.random.choice(x[0] - k); b = np.random.choice(x[1] - k), per patch
    np.random.seed(5)
    got = patches.sample_patch_coords((30, 41), 7, 50)
    np.random.seed(5)
    want = []
    for _ in range(50):
        a = np.random.choice(30 - 7)
        b = np.random.choice(41 - 7)
        want.append([a, b])
    assert got.tolist() == want
    assert got[:, 0].max() < 30 - 7 and got[:, 1].max() < 41 - 7          # the last valid offset is never drawn, like the reference


def test_graph_to_csr_networkx_and_scipy():
    import networkx as nx
    import scipy.sparse as sp
    G = nx.Graph()
    G.add_nodes_from(["a", "b", "c", "d", "iso"])
    G.add_edges_from([("a", "c"), ("c", "b"), ("d", "a"), ("c", "c")])
    rowptr, colidx, nodes = patches.graph_to_csr(G)
    assert nodes == ["a", "b", "c", "d", "iso"] and rowptr.dtype == np.int64 and colidx.dtype == np.int32
    nb = [colidx[rowptr[i]:rowptr[i + 1]].tolist() for i in range(5)]
    assert nb == [[2, 3], [2], [0, 1, 2], [0], []]                        # sorted, both directions, self-loop once, isolated node
    for u in G.nodes():
        for v in G.nodes():
            i, j = nodes.index(u), nodes.index(v)
            assert (j in nb[i]) == G.has_edge(u, v)
    M = sp.csr_matrix(np.array([[0, 1, 0], [1, 0, 1], [0, 1, 1]], dtype=float))
    rp, ci, nd = patches.graph_to_csr(M)
    assert rp.tolist() == [0, 1, 3, 5] and ci.tolist() == [1, 0, 2, 1, 2] and nd == [0, 1, 2]


def test_patches_to_tensor_is_the_reference_tensor():
    rng = np.random.default_rng(1)
    img = rng.random((14, 15, 3))
    co = np.array([[0, 0], [3, 4], [8, 9]], dtype=np.int32)
    T = O.gather_patches_color_tensor(img, co, 6)                          # reference: (k*k, 3, N)
    X = T.reshape(6 * 6 * 3, 3)                                            # HWC data matrix = mode-2 joint unfolding
    assert np.array_equal(patches.patches_to_tensor(X, 3), T)


def test_simple_graph_edges_rounding_rule():
    # network_reconstruction_nx.py:499-507: an undirected edge when the directed pair's mean weight rounds to > 0
    pairs = np.array([[0, 1], [1, 0], [2, 3], [3, 3], [4, 5]], dtype=object)
    w = np.array([0.49, 0.51, 0.5, 1.7, -0.2])
    assert reconstruct.simple_graph_edges(pairs, w) == {frozenset((0, 1)), frozenset((3,))}      # round(0.5) = 0 (banker's), like numpy


def test_precision_names_and_errors():
    import torch
    assert _host.torch_dtype(None) == torch.float32 and _host.torch_dtype("fp64") == torch.float64
    assert _host.torch_dtype(np.float32) == torch.float32 and _host.torch_dtype(torch.float64) == torch.float64
    with pytest.raises(ValueError):
        _host.torch_dtype("bf16")
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            _host.device()                                                  # no CPU path: loud, not a silent fallback
