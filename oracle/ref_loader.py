"""Import the UNMODIFIED reference (`/root/reference`) behind stub gym/gymnasium/mpi4py modules.

ORACLE / TEST INFRASTRUCTURE.  Only usable in the build container: /root/reference does not exist on
the GPU box, so nothing marked `-m gpu`, `smoke()` or `bench.py` calls this.  It is used by
oracle/make_goldens.py (to generate tests/golden/*.npz) and by the CPU tests that pin oracle/ref_port.py
against the live reference when the tree is present.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("XB200_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "xuance"))


def load(trig="libm"):
    """Returns the imported `xuance` package (reference code, stub third-party deps)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (_REPO, _STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import gym_restated
    gym_restated.DEFAULT_TRIG = trig
    import xuance  # noqa: F401
    return xuance
