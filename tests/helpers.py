"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "meta" in d:
        d["meta"] = json.loads(str(d["meta"]))
    return d


def replay_buffer_protocol(buf, g):
    """Drives a DummyOnPolicyBuffer-shaped object with the agent's store/finish_path protocol
    (ppoclip_agent.py:68-100) using the golden's recorded rollout."""
    T, N = g["rew"].shape
    for t in range(T):
        buf.store(g["obs"][t], g["act"][t], g["rew"][t], g["val"][t], g["term"][t], {"old_logp": g["logp"][t]})
        if buf.full:
            for i in range(N):
                buf.finish_path(0.0 if g["term"][t, i] else g["boot"][t, i], i)
            break
        for i in range(N):
            if g["term"][t, i] or g["trunc"][t, i]:
                buf.finish_path(0.0 if g["term"][t, i] else g["boot"][t, i], i)


def gae_close(a, b, rtol=1e-5):
    """|a-b| <= rtol*max(|b|, rms(b)) (advantages cross zero)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = np.maximum(np.abs(b), np.sqrt(np.mean(b * b)))
    return bool(np.all(np.abs(a - b) <= rtol * scale)), float(np.max(np.abs(a - b) / scale))


def rel_close(a, b, rtol):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    denom = max(float(np.sqrt(np.mean(b * b))), 1e-30)
    err = float(np.max(np.abs(a - b)) / denom)
    return err <= rtol, err
