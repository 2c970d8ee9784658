"""The oracle (oracle/onmf_oracle.py numpy restatement, oracle/lars_oracle.c C restatement) against
(a) the golden fixtures written by the UNMODIFIED reference (oracle/make_golden.py) and
(b) scikit-learn 1.9.0, the dependency the reference's sparse coder calls (src/ontf.py:79-86).
CPU only."""
import os
import warnings

import numpy as np
import pytest

from oracle import c_oracle
from oracle import onmf_oracle as O

warnings.filterwarnings("ignore")

CASES = ["cfg1_renoir_gray", "cfg1_renoir_gray_epoch2", "cfg1_alphaNone_beta_full", "cfg1_alpha0",
         "cfg2_renoir_color_tensor", "cfg3_binary_motif", "cfg4_ising_pm1"]


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def step_inputs(g, i):
    W = g["W0"] if i == 0 else g["W_%d" % (i - 1)]
    if i == 0:
        A = g["A0"] if "A0" in g.files else np.zeros((W.shape[1], W.shape[1]))
        B = g["B0"] if "B0" in g.files else np.zeros((W.shape[1], W.shape[0]))
    else:
        A, B = g["A_%d" % (i - 1)], g["B_%d" % (i - 1)]
    return g["X"][:, g["idx"][i]], W, A, B


@pytest.mark.parametrize("name", CASES)
def test_codes_match_reference_every_step(golden_dir, name):
    """per-minibatch codes of the restated LARS (C) == reference lasso_lars codes, every recorded step."""
    g = load(golden_dir, name)
    for i in range(int(g["n_steps"])):
        Xb, W, _, _ = step_inputs(g, i)
        H = c_oracle.sparse_code(Xb, W, float(g["alpha"]))
        assert rel(H, g["H_%d" % i]) < 1e-9, (name, i)


@pytest.mark.parametrize("name", ["cfg1_renoir_gray", "cfg3_binary_motif", "cfg1_alpha0"])
def test_numpy_restatement_matches_reference(golden_dir, name):
    g = load(golden_dir, name)
    i = int(g["n_steps"]) - 1
    Xb, W, _, _ = step_inputs(g, i)
    Xb = Xb[:, :60]
    H = O.sparse_code_lars(Xb, W, float(g["alpha"]))
    assert rel(H, g["H_%d" % i][:, :60]) < 1e-9


@pytest.mark.parametrize("name", CASES)
def test_full_loop_matches_reference(golden_dir, name):
    """W, A, B after the whole recorded run (C oracle for the codes, numpy for the rest)."""
    g = load(golden_dir, name)
    beta = float(g["beta"]) if "beta" in g.files else None
    hist = float(g["history_in"]) if "history_in" in g.files else 0.0
    W = g["W0"]
    k, d = W.shape[1], W.shape[0]
    A = g["A0"] if "A0" in g.files else np.zeros((k, k))
    B = g["B0"] if "B0" in g.files else np.zeros((k, d))
    for i in range(int(g["n_steps"])):
        Xb = g["X"][:, g["idx"][i]]
        t = hist + i + 1
        assert t == float(g["t_%d" % i])
        H = c_oracle.sparse_code(Xb, W, float(g["alpha"]))
        A1, B1 = O.aggregate(A, B, H, Xb, t, beta)
        W = O.update_dict(W, A, B)          # OLD aggregates (src/ontf.py:151)
        A, B = A1, B1
        assert rel(W, g["W_%d" % i]) < 1e-9 and rel(A, g["A_%d" % i]) < 1e-9 and rel(B, g["B_%d" % i]) < 1e-9
    assert rel(W, g["W_final"]) < 1e-9
    assert float(g["history_out"]) == hist + int(g["n_steps"]) + 1


def test_cfg5_codes(golden_dir):
    g = load(golden_dir, "cfg5_synthetic")
    X = np.random.RandomState(int(g["x_seed"])).rand(1024, 160)
    W0 = np.random.RandomState(int(g["w0_seed"])).rand(1024, 256)
    H = c_oracle.sparse_code(X[:, g["idx"][0]], W0, 1.0)
    assert rel(H, g["H_0"]) < 1e-9


def test_c_step_equals_numpy_step(golden_dir):
    g = load(golden_dir, "cfg4_ising_pm1")
    Xb, W, A, B = step_inputs(g, 2)
    Xb = Xb[:, :40]
    H1, A1, B1, W1 = c_oracle.step(Xb, A, B, W, 3.0, 1.0)
    H2, A2, B2, W2 = O.step(Xb, A, B, W, 3.0, 1.0, coder="lars")
    assert rel(H1, H2) < 1e-10 and rel(A1, A2) < 1e-10 and rel(B1, B2) < 1e-10 and rel(W1, W2) < 1e-12


@pytest.mark.parametrize("kind,alpha", [("rand", 1.0), ("norm", 1.0), ("norm", 0.0), ("pm1", 1.0), ("norm", 0.05)])
def test_restated_lars_vs_sklearn(kind, alpha):
    """the restated solver against the dependency itself (sklearn lasso_lars, positive) on fresh inputs,
    including the alpha=0 regime and +-1 data (X may be negative, ising_reconstruction.py:114)."""
    rng = np.random.default_rng(7)
    d, k, n = 64, 20, 40
    W = rng.random((d, k))
    if kind != "rand":
        W /= np.linalg.norm(W, axis=0)
    X = rng.choice([-1.0, 1.0], size=(d, n)) if kind == "pm1" else rng.random((d, n))
    X[:, 3] = 0.0
    Href = O.sparse_code_sklearn(X, W, alpha)
    assert rel(c_oracle.sparse_code(X, W, alpha), Href) < 1e-9
    assert rel(O.sparse_code_lars(X, W, alpha), Href) < 1e-9
    assert np.all(Href[:, 3] == 0.0)


def test_update_dict_invariants_and_c_port():
    rng = np.random.default_rng(3)
    W = rng.random((50, 12))
    H = rng.random((12, 30))
    A, B = H @ H.T, H @ rng.random((30, 50))
    W1 = O.update_dict(W, A, B)
    assert W1.min() >= 0.0 and np.all(np.linalg.norm(W1, axis=0) <= 1.0 + 1e-12)   # SURVEY §4 invariants
    assert rel(c_oracle.update_dict(W, A, B), W1) < 1e-13
    # with zero aggregates the sweep only clamps / shrinks into the unit ball
    W2 = O.update_dict(W, np.zeros((12, 12)), np.zeros((12, 50)))
    assert rel(W2, W / np.maximum(1.0, np.linalg.norm(W, axis=0))) < 1e-14


def test_pgd_coder_golden(golden_dir):
    g = load(golden_dir, "pgd_coder")
    H = O.update_code_within_radius(g["X"], g["W"], g["H0"], r=None, alpha=1, sub_iter=10, stopping_diff=0.01)
    assert rel(H, g["H"]) < 1e-12
    Hr = O.update_code_within_radius(g["X"], g["W"], g["H0"], r=0.5, alpha=0.3, sub_iter=3, stopping_diff=0.01)
    assert rel(Hr, g["H_radius"]) < 1e-10
    H1 = O.update_code_within_radius(g["X"][:, :1], g["W"], g["H0"][:, :1], r=None, alpha=1, sub_iter=10,
                                     stopping_diff=0.01)
    assert rel(H1, g["H_single"]) < 1e-12


def test_gather_and_matricize_golden(golden_dir):
    g = load(golden_dir, "cfg1_renoir_gray")
    X = O.gather_patches_gray(g["img"], g["coords"], int(g["patch"]))
    assert np.array_equal(X, g["X"])
    g2 = load(golden_dir, "cfg2_renoir_color_tensor")
    T = O.gather_patches_color_tensor(g2["img"], g2["coords"], int(g2["patch"]))
    assert np.array_equal(T, g2["T"])
    Xm = O.matricize(T, int(g2["mode"]), bool(g2["joint"]))
    assert np.array_equal(Xm, g2["X"])            # (300 x N): feature f = (row*10 + col)*3 + channel
    assert Xm[(2 * 10 + 3) * 3 + 1, 5] == T[2 * 10 + 3, 1, 5]


def test_reconstruction_loop_golden(golden_dir):
    """oracle restatement of the per-patch reconstruction loop (image_reconstruction.py:358-406) against the fixture made
    with the reference's own update_code_within_radius."""
    g = load(golden_dir, "reconstruct_color")
    rec, cnt, codes = O.reconstruct_image_loop(g["img"], g["W"], int(g["patch"]), int(g["stride"]), 1, 10, 0.01, g["H0"])
    assert rel(codes, g["codes"]) < 1e-12 and rel(rec, g["recons"]) < 1e-12 and np.array_equal(cnt, g["count"])
    assert cnt.max() == 9 and cnt[-1, -1] == 0          # the loop never reaches the last row/column (range(0, H-k, res))


def test_full_cfg5_every_step(golden_dir):
    """the benchmarked shape (d=1024, k=256) on the learned dictionaries of a 10-step reference run: the restated LARS (C)
    reproduces the reference's codes at every step (192 columns per step; the aggregation uses the reference's full code
    matrix so that all ten dictionaries are exactly the reference's), and the restated A/B recursion + dictionary
    update reproduce the final state."""
    g = load(golden_dir, "full_cfg5")
    ntr, k = int(g["n_train"]), int(g["k"])
    X = np.random.RandomState(int(g["x_seed"])).rand(1024, ntr + int(g["n_holdout"]))
    W = np.random.RandomState(int(g["seed"])).rand(1024, k)
    assert abs(W.sum() - float(g["W0_checksum"])) < 1e-9
    A, B = np.zeros((k, k)), np.zeros((k, 1024))
    for i in range(int(g["iters"])):
        Xb = X[:, g["idx"][i]]
        Href = g["H_%d" % i].T                       # fixture stores the reference's H1 (n x r)
        H = c_oracle.sparse_code(Xb[:, :192], W, 1.0)
        assert rel(H, Href[:, :192]) < 1e-9, i
        A1, B1 = O.aggregate(A, B, Href, Xb, float(i + 1))
        W = O.update_dict(W, A, B)
        A, B = A1, B1
    assert rel(W, g["W"]) < 1e-9 and rel(A, g["A"]) < 1e-9 and rel(B, g["B"]) < 1e-9
    assert float(g["history"]) == int(g["iters"]) + 1


def test_network_reconstruction_golden(golden_dir):
    """oracle restatement of the per-step reconstruction loop (network_reconstruction_nx.py:464-491) against the weights /
    overlap counts the UNMODIFIED driver method produced (oracle/make_golden.py, fixture network_recons)."""
    g = load(golden_dir, "network_recons")
    adj = {}
    for a, b in g["graph_edges"].tolist():
        adj.setdefault(a, set()).add(b)
        adj.setdefault(b, set()).add(a)
    W = g["W"]
    w, c = O.reconstruct_network_loop(adj, W, g["embs"], 0.0, coder=lambda p: c_oracle.sparse_code(p, W, 0.0))
    pairs = [tuple(p) for p in g["pairs"].tolist()]
    assert sorted(w.keys()) == pairs
    assert max(abs(w[p] - x) for p, x in zip(pairs, g["weight"].tolist())) < 1e-12
    assert all(c[p] == x for p, x in zip(pairs, g["count"].tolist()))
    simple = {frozenset(p) for p, x in w.items() if np.round(x) > 0}
    assert simple == {frozenset(e) for e in g["simple_edges"].tolist()}
