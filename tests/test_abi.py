"""The C-ABI library loads (no GPU needed) and exports exactly the symbols include/marl_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "marl_b200.h")).read()
    src = re.sub(r"#if 0.*?#endif /\* MQ_PENDING \*/", "", src, flags=re.S)      # not yet enabled sections
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mq_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from dqn_marl_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.SO_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in marl_b200.h but not exported by libmarl_b200.so"
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"
    assert _lib.load().mq_abi_version() == 3


def test_floor_field_errors_are_reported():
    import numpy as np
    import pytest
    from dqn_marl_b200 import _lib
    with pytest.raises(_lib.MqError) as ei:
        _lib.floor_field(4, 4, np.zeros((6, 6), np.uint8), np.array([[9, 9]], np.int32), np.zeros((6, 6)))
    assert "outside the grid" in str(ei.value)


def test_product_never_imports_the_oracle():
    """The product package must not reference oracle/ (a CPU fallback would void the parity claims)."""
    pkg = os.path.join(ROOT, "dqn_marl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|env_oracle|keyed_draws|ref_harness|replay_oracle|qnet_oracle|floor_field_py)", txt, flags=re.M), f
                assert "liborc" not in txt, f
