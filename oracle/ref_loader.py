"""Import the UNMODIFIED reference behind stub gym/gymnasium/mpi4py modules.

ORACLE / TEST INFRASTRUCTURE.  Two places hold the reference:
  * /root/reference           the source tree — build container only.  Used by oracle/make_goldens*.py (to generate
                              tests/golden/*.npz) and by the CPU tests that pin oracle/ref_port.py against the live
                              reference (tests/test_oracle_vs_reference.py);
  * oracle/_ref/              the same package pip-installed by oracle/build_ref.py (git-ignored; it travels to the GPU
                              box with the snapshot).  Used there only by `bench.py --impl reference` / `cpu_baseline`.
Nothing marked `-m gpu` and not smoke() reads either.  The product never imports this.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")
_REPO = os.path.dirname(_HERE)
_CANDIDATES = [os.environ.get("XB200_REFERENCE_ROOT", "/root/reference"), os.path.join(_HERE, "_ref")]


def reference_root():
    for root in _CANDIDATES:
        if os.path.isdir(os.path.join(root, "xuance", "torch")):
            return root
    return None


REFERENCE_ROOT = reference_root() or _CANDIDATES[0]


def available():
    return reference_root() is not None


def source_tree_available():
    """True only where the reference SOURCE tree is present (golden generation, live-reference pinning tests)."""
    return os.path.isdir(os.path.join(_CANDIDATES[0], "xuance", "torch"))


def load(trig="libm"):
    """Returns the imported `xuance` package (reference code, stub third-party deps)."""
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not present (neither %s nor %s)" % tuple(_CANDIDATES))
    for p in (_REPO, _STUBS, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import gym_restated
    gym_restated.DEFAULT_TRIG = trig
    import xuance  # noqa: F401
    return xuance
