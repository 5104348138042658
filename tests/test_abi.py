"""The C-ABI library loads without a GPU and exports every symbol include/scann_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scann_b200.h")


def _production_header():
    """The header without comments and without the development-probe block (#ifdef SCANN_DEV_PROBES)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"#ifdef SCANN_DEV_PROBES.*?#endif", "", src, flags=re.S)


def header_symbols():
    src = _production_header()
    return sorted(set(re.findall(r"\b(scann_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    from scann_b200 import build
    build.build()
    from scann_b200 import _abi
    assert os.path.exists(_abi.LIB_PATH)
    assert _abi.lib.scann_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    from scann_b200 import _abi
    syms = header_symbols()
    assert len(syms) >= 20
    raw = ctypes.CDLL(_abi.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in the header but not exported"
        assert s in _abi.PROTOTYPES, f"{s} has no ctypes prototype"
    for s in _abi.PROTOTYPES:
        assert s in syms, f"{s} bound in _abi.py but not declared in the header"
    # the production library carries no development probes (timestamp stores, tcgen05 self-tests)
    if "SCANN_DEV_PROBES" not in os.environ.get("SCANN_NVCC_DEFS", ""):
        for s in _abi.DEV_PROTOTYPES:
            assert not hasattr(raw, s), f"{s} is a development probe and must not be in the production library"


def test_prototype_arity_matches_header():
    from scann_b200 import _abi
    src = _production_header()
    for name, (_, args) in _abi.PROTOTYPES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, ctypes prototype {len(args)}"


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product path must raise, never fall back to the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scann_b200 import _abi
    assert _abi.lib.scann_device_sm_count() == -1
    with pytest.raises(_abi.ScannAbiError):
        _abi.require_gpu()
    from scann_b200.configs import get_config
    from scann_b200.model import create_model
    with pytest.raises(Exception):
        create_model(get_config("qm9"))


def test_allreduce_entry_points_refuse_to_run_without_a_communicator():
    """scann_allreduce_* (NCCL bound at run time): without scann_allreduce_init the sum must fail loudly, and the library
    itself must not depend on NCCL at load time."""
    from scann_b200 import _abi
    assert _abi.lib.scann_allreduce_world() == 0
    assert _abi.lib.scann_allreduce_sum(None, 16, None) != 0
    assert "no communicator" in _abi.last_error()
    assert _abi.lib.scann_allreduce_destroy() == 0
    import subprocess
    deps = subprocess.run(["ldd", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in deps.lower()
