"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
and the ctypes struct layout matches (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

from gnn_fraud_detection_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gnnfd_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnnfd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("gnnfd_csr_build", "gnnfd_project_fwd", "gnnfd_gat_fwd", "gnnfd_gat_bwd_dst", "gnnfd_gat_bwd_src",
                 "gnnfd_project_bwd", "gnnfd_last_error"):
        assert must in syms
    assert set(syms) == set(_abi.SIGNATURES), "ctypes binding and header disagree on the symbol set"


def test_library_loads_and_exports_every_declared_symbol():
    assert os.path.exists(_abi.LIB_PATH), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    raw = C.CDLL(_abi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} is declared in include/gnnfd_b200.h but not exported"
    lib = _abi.lib()
    assert lib.gnnfd_abi_version() == _abi.ABI_VERSION
    assert lib.gnnfd_sizeof_graph() == C.sizeof(_abi.Graph)
    assert lib.gnnfd_sizeof_hub_plan() == C.sizeof(_abi.HubPlan)


def test_host_side_argument_errors_need_no_gpu():
    lib = _abi.lib()
    nb = C.c_size_t()
    assert lib.gnnfd_csr_workspace_bytes(10, 20, _abi.ADD_SELF_LOOPS, C.byref(nb)) == 0 and nb.value > 0
    # E + N beyond int32 is refused with a message, not truncated
    assert lib.gnnfd_csr_workspace_bytes(2 ** 30, 2 ** 31, _abi.ADD_SELF_LOOPS, C.byref(nb)) == -4
    assert b"int32" in lib.gnnfd_last_error()
    with pytest.raises(_abi.GnnfdError):
        _abi.check(lib.gnnfd_csr_workspace_bytes(-1, 0, 0, C.byref(nb)))
    g = _abi.Graph()
    assert lib.gnnfd_gat_fwd_workspace_bytes(C.byref(g), 8, 64, C.byref(nb)) == 0
    # unsupported head geometry fails loudly instead of falling back
    g.n_dst = 0
    rc = lib.gnnfd_gat_fwd(C.byref(g), 16, _abi.F32, 16, 16, None, 3, 17, 0.2, 0, 0, None, 0.0, 0, 16, 16, 16, None, 0, None)
    assert rc == -5 and b"not built" in lib.gnnfd_last_error()
