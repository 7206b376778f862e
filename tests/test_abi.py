"""CPU: the built shared library loads without a GPU and exports every function include/mtt_b200.h declares;
the ctypes binding (lib.py) covers exactly that set. No compute is called."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "mtt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import mtt_b200  # noqa: F401
    from mtt_b200 import lib

    assert os.path.exists(lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    cdll = ctypes.CDLL(lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/mtt_b200.h but not exported"
    assert set(names) == set(lib.SYMBOLS), set(names) ^ set(lib.SYMBOLS)


def test_library_refuses_to_run_without_sm100():
    import mtt_b200  # noqa: F401
    from mtt_b200 import lib
    import torch

    l = lib.load()
    assert l.mtt_version() >= 100
    if not torch.cuda.is_available():
        assert l.mtt_device_check() != 0
        assert b"CUDA" in l.mtt_last_error() or b"device" in l.mtt_last_error()
