"""The C-ABI library loads (no GPU needed) and exports every symbol include/onmf_b200.h declares;
host-side entry points that need no device behave."""
import ctypes
import os
import re

import pytest

from onmf_ontf_ndl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "onmf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(onmf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 17
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libonmf_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "python binding misses %s" % s
    assert sorted(_lib.SIGNATURES) == syms


def test_identification():
    lib = _lib.load()
    assert lib.onmf_version() == 100
    assert lib.onmf_built_arch() == 100        # sm_100a only


def test_workspace_queries_are_host_only():
    import torch
    assert _lib.lasso_lars_workspace(torch.float32, 25, 1000) >= 64 + 8 * 1000
    assert _lib.lasso_lars_workspace(torch.float64, 256, 4096) >= _lib.lasso_lars_workspace(torch.float32, 256, 4096) > 0
    assert _lib.lasso_lars_workspace(torch.float32, 1000, 10) == 0           # > 512 atoms: not instantiated
    assert _lib.surrogate_workspace(torch.float32, 262144, 256, 1024) > 0


def test_argument_errors_are_reported_not_crashes():
    lib = _lib.load()
    rc = lib.onmf_lasso_lars(0, None, None, 10, 25, 100, 1.0, 1000, None, None, 0, None, None)
    assert rc == -1 and b"null" in lib.onmf_last_error()
    rc = lib.onmf_update_dict(7, None, None, None, 10, 5, None, None)
    assert rc == -1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from onmf_ontf_ndl_b200 import Online_NTF
    m = Online_NTF(np.random.rand(10, 20, 1), n_components=3, iterations=2, batch_size=5)
    with pytest.raises(_lib.OnmfKernelError):
        m.train_dict_single()
    with pytest.raises(_lib.OnmfKernelError):
        m.joint_sparse_code_tensor(np.random.rand(10, 4), np.random.rand(10, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "onmf_ontf_ndl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in txt.replace("# oracle", ""), fn
            assert "sklearn" not in txt.split('"""')[-1], fn


def test_bench_reference_arm_prints_contract_line():
    """bench.py --impl reference: the CPU arm (oracle port of the reference's numpy + sklearn step) on a tiny bounded
    sample, two worker processes; one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "0", "--cpu-cols", "64", "--cpu-procs", "2"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "samples/s" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == 2 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_step_buffers_struct_layout_matches_header(tmp_path):
    """ctypes mirror of onmf_step_buffers vs the C definition: same size and field offsets (compiled with gcc)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fields = [f[0] for f in _lib.StepBuffers._fields_]
    src = tmp_path / "layout.c"
    body = "\n".join('  printf("%s %%zu\\n", offsetof(onmf_step_buffers, %s));' % (f, f) for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "onmf_b200.h"\nint main(void) {\n'
                   '  printf("sizeof %zu\\n", sizeof(onmf_step_buffers));\n' + body + "\n  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(_lib.StepBuffers)
    for f in fields:
        assert int(out[f]) == getattr(_lib.StepBuffers, f).offset, f
