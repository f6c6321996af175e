"""CPU: the C-ABI library loads, exports every symbol include/xb200.h declares, and the ctypes binding has the
same arity as the header.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_protos():
    h = open(os.path.join(REPO, "include", "xb200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    protos = re.findall(r"\n(?:int|const char\*) (xb_\w+)\(([^;]*?)\);", h, re.S)
    return {name: (0 if args.strip() == "void" else len(args.split(","))) for name, args in protos}


def test_library_builds_loads_and_exports_every_declared_symbol():
    from xuanpolicy_b200 import _lib
    from xuanpolicy_b200.csrc import build
    build.build()
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    protos = _header_protos()
    assert len(protos) >= 16
    for name in protos:
        assert hasattr(raw, name), "missing export %s" % name
    assert lib.xb_version() == 100
    assert lib.xb_error_string(-1).decode().startswith("xb200: bad argument")


def test_binding_arity_matches_header():
    from xuanpolicy_b200 import _lib
    protos = _header_protos()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in protos, name
        assert protos[name] == len(argtypes), (name, protos[name], len(argtypes))
    assert set(protos) - set(_lib.SIGNATURES) == {"xb_error_string"}


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: constructing the drop-ins on a CPU device raises."""
    import torch
    import xuanpolicy_b200 as xb
    with pytest.raises(RuntimeError):
        xb.DummyVecEnv_Gym(xb.make_env_fns("CartPole-v1", 1, 4), device="cpu")
    obs_space, act_space = xb.make_spaces("CartPole-v1")
    with pytest.raises(RuntimeError):
        xb.DummyOnPolicyBuffer(obs_space, act_space, {"old_logp": ()}, 4, 8, device="cpu")
    from xuanpolicy_b200 import ops
    with pytest.raises(xb.XB200Error):
        ops.sincos_f64(torch.zeros(4, dtype=torch.float64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "xuanpolicy_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
