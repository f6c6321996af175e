"""Minimal stand-in for the (absent) third-party `gym==0.26.2` package.

TEST INFRASTRUCTURE ONLY.  It exists so that the *unmodified* reference tree
(/root/reference, importable only in the build container) can be imported and
run to generate golden vectors (oracle/make_goldens.py).  `make()` hands out
the restated classic-control environments from oracle/gym_restated.py, which
is a restatement of gym 0.26.2's published equations.
Nothing under xuanpolicy_b200/ may import this.
"""
from . import spaces
from .spaces import Space
from .core import Env, Wrapper
from . import error, utils


def make(env_id, render_mode=None, **kwargs):
    from oracle import gym_restated
    return gym_restated.make(env_id, render_mode=render_mode, **kwargs)
