"""No-GPU checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every
symbol include/macvo_b200.h declares (no compute call is made)."""
import os
import re
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from macvo_b200 import build, ops
    build.build(verbose=False)
    return ops.load_library()


def _header_symbols():
    text = open(os.path.join(REPO, "include", "macvo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(macvo_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from macvo_b200 import ops
    declared = _header_symbols()
    assert len(declared) >= 13
    assert sorted(ops.EXPORTS) == declared, "ops.EXPORTS must bind exactly what the header declares"
    for name in declared:
        assert hasattr(lib, name), name


def test_version_string(lib):
    from macvo_b200 import ops
    assert ops.version().startswith("macvo_b200") and "sm_100a" in ops.version()


def test_host_only_queries(lib):
    """workspace-size functions are pure host arithmetic"""
    assert lib.macvo_corr_workspace_bytes(2, 256, 4800, 0) == 0
    assert lib.macvo_corr_workspace_bytes(2, 256, 4800, 1) == 4 * 2 * 4800 * 256 * 2
    assert lib.macvo_corr_workspace_bytes(2, 256, 4800, 2) == 2 * 2 * 4800 * 256 * 2
    assert lib.macvo_select_workspace_bytes(480, 640) >= 480 * 640


def test_sass_is_blackwell_native():
    """the tensor-core kernel must contain tcgen05 / TMA SASS (UTCHMMA, UTMALDG, UTCCP, LDTM)"""
    from macvo_b200 import build
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UTCCP", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from macvo_b200 import ops
    with pytest.raises(ops.MacvoB200Error):
        ops.corr_build(torch.zeros(1, 64, 4, 4), torch.zeros(1, 64, 4, 4))
    with pytest.raises(ops.MacvoB200Error):
        ops.corr_lookup(torch.zeros(16, 1, 4, 4), torch.zeros(1, 2, 4, 4))
