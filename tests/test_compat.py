"""The reference's drivers import `src.onmf`, `utils.onmf`, `utils.ontf`; compat/ provides those paths."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_driver_import_paths_and_constructor_forms():
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        for m in ("src", "src.onmf", "src.ontf", "utils", "utils.onmf", "utils.ontf"):
            sys.modules.pop(m, None)
        from src.onmf import Online_NMF, update_code_within_radius            # image_reconstruction.py:1
        from utils.onmf import Online_NMF as NMF2                              # ising_reconstruction.py:1
        from utils.ontf import Online_NTF                                      # image_reconstruction_tensor.py:1
        import onmf_ontf_ndl_b200 as pkg
        assert Online_NMF is pkg.Online_NMF and NMF2 is pkg.Online_NMF and Online_NTF is pkg.Online_NTF
        assert update_code_within_radius is pkg.update_code_within_radius
        X = np.random.rand(12, 30)
        a = Online_NMF(X, 5, 7, 3)                                            # image_reconstruction.py:349 positional form
        assert (a.n_components, a.iterations, a.batch_size) == (5, 7, 3) and a.code.shape == (5, 30)
        b = Online_NMF(X)                                                      # ising_reconstruction.py:194
        assert b.n_components == 100 and b.history == 0 and b.subsample is False
        c = Online_NMF(X, n_components=4, iterations=3, batch_size=2, ini_dict=None, ini_A=None, ini_B=None, ini_C=None,
                       history=0, alpha=None, beta=0.5)                        # ising_reconstruction.py:116-126
        assert c.beta == 0.5 and c._alpha() == 0
        t = Online_NTF(np.random.rand(9, 3, 20), 6, iterations=4, sub_iterations=2, learn_joint_dict=True, mode=2,
                       batch_size=5)                                           # image_reconstruction_tensor.py:234-239
        assert t.code.shape == (3, 6) and t._alpha() == 2 and t.subsample is True
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        for m in ("src", "src.onmf", "src.ontf", "utils", "utils.onmf", "utils.ontf"):
            sys.modules.pop(m, None)
