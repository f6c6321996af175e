"""Build and drive the LIVE reference PPO agent (unmodified xuance code) on the restated classic-control physics.

ORACLE / TEST INFRASTRUCTURE (golden generation, live-reference pinning tests, bench.py's CPU arms).
`build_runner` goes through the reference's own entry point `xuance.get_runner` (xuance/common/common_tools.py:85-164):
YAML config (xuance/configs/ppo/classic_control/*.yaml) + overrides -> Runner_DRL (runner_drl.py:15-74) -> make_envs
(environment/__init__.py:36-90) -> DummyVecEnv_Gym of Gym_Env -> Basic_MLP / *_AC_Policy -> Adam + LinearLR ->
PPOCLIP_Agent.  Logs and model directories land in a scratch directory.
"""
import contextlib
import io
import os
import tempfile
from argparse import Namespace

from . import ref_loader


@contextlib.contextmanager
def _scratch_cwd():
    old = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="xb200_refrun_")
    os.chdir(tmp)
    try:
        yield tmp
    finally:
        os.chdir(old)


def build_runner(env_id, trig="libm", quiet=True, method="ppo", **overrides):
    """The reference's Runner_DRL for `method` ("ppo", "pg", "ppg", ...) on `env_id`; `overrides` are parser_args (parallels, n_steps, seed,
    representation_hidden_size, use_obsnorm, ...).  runner.agent is the live PPOCLIP_Agent, runner.envs its
    DummyVecEnv_Gym (already reset, runner_basic.py:12)."""
    ref_loader.load(trig=trig)
    import xuance
    args = dict(device="cpu", render=False, test_mode=False, logger="tensorboard", running_steps=10 ** 9)
    args.update(overrides)
    out = io.StringIO()
    with _scratch_cwd():
        with (contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext()):
            runner = xuance.get_runner(method=method, env="classic_control", env_id=env_id, parser_args=Namespace(**args))
    return runner


def silence_progress():
    """PPOCLIP_Agent.train wraps its loop in tqdm (ppoclip_agent.py:61); route it to a null stream for timing runs."""
    os.environ.setdefault("TQDM_DISABLE", "1")
