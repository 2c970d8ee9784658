"""Import shim: `from src.onmf import Online_NMF, update_code_within_radius` (reference
image_reconstruction.py:1) / `from utils.onmf import Online_NMF` (ising_reconstruction.py:1)."""
from onmf_ontf_ndl_b200.onmf import DEBUG, Online_NMF, update_code_within_radius  # noqa: F401
