"""Recipe: install the UNMODIFIED reference (/root/reference) into the git-ignored oracle/_ref/.

ORACLE / TEST INFRASTRUCTURE.  Build container only (the GPU box has no /root/reference; it receives the built
oracle/_ref/ with the snapshot, like our own .so files).  Nothing is copied into the tracked tree.

    python -m oracle.build_ref [--force]

pip-installs the reference from a scratch copy under /tmp (the build writes egg-info into its source tree and
/root/reference is read-only) with `--no-index --no-build-isolation --no-deps --target oracle/_ref`.  Its setup.py
declares `setup_requires=['pytest-runner']`, which is not in the offline wheelhouse and is not needed to install: the
recipe satisfies it with an empty dist-info on PYTHONPATH (the reference's own files are not touched).  The third-party
packages the reference imports but this image lacks (gym 0.26.2, gymnasium, mpi4py) stay the ~100-line stubs of
oracle/stubs/ with the restated classic-control physics of oracle/gym_restated.py — see ref_loader.py.

`bench.py --impl reference` imports `xuance` from oracle/_ref/ (cpu_baseline.kind = "reference": the live
PPOCLIP_Agent.train over DummyVecEnv_Gym / DummyOnPolicyBuffer / PPOCLIP_Learner); without it, it falls back to the
restatement oracle/ref_port.py (kind = "port").
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("XB200_REFERENCE_ROOT", "/root/reference")


def installed():
    return os.path.isdir(os.path.join(DEST, "xuance", "torch"))


def build(force=False):
    """Returns DEST if the reference is installed there (now or before), else None (no source tree here)."""
    if installed() and not force:
        return DEST
    if not os.path.isdir(os.path.join(SRC, "xuance")):
        return None
    tmp = tempfile.mkdtemp(prefix="xb200_ref_")
    try:
        work = os.path.join(tmp, "src")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "docs", "*.gif", "*.png", "*.jpg"))
        fake = os.path.join(tmp, "site", "pytest_runner-6.0.1.dist-info")
        os.makedirs(fake)
        with open(os.path.join(fake, "METADATA"), "w") as f:
            f.write("Metadata-Version: 2.1\nName: pytest-runner\nVersion: 6.0.1\n")
        env = dict(os.environ)
        env["PYTHONPATH"] = os.path.join(tmp, "site") + os.pathsep + env.get("PYTHONPATH", "")
        if os.path.isdir(DEST):
            shutil.rmtree(DEST)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DEST, work]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        if r.returncode != 0 or not installed():
            raise RuntimeError("installing the reference into oracle/_ref failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-2000:]))
        return DEST
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
