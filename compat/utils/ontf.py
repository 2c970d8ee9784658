"""Import shim: `from utils.ontf import Online_NTF` (reference image_reconstruction_tensor.py:1)."""
from onmf_ontf_ndl_b200.ontf import DEBUG, Online_NTF  # noqa: F401
