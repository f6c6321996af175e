"""%globaltimer stamps of the rollout's two kernels (-DXB_STEP_TS build): where a vector step's time goes.

    python tools/profile/step_timeline.py build      (build container: compiles tools/profile/_dbg/libxb200_sts.so)
    XB200_LIB=tools/profile/_dbg/libxb200_sts.so python tools/profile/step_timeline.py run   (B200)

Tags: rollout step kernel 100 entry, 101 after the dependency wait, 104 env + store done (CTA 0), 102 last CTA enters the
statistics tail, 103 tail done; forward kernel (CTA 0) 200 entry, 201 set-up done, 202 after the dependency wait, 203 done.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
DBG = os.path.join(HERE, "_dbg")


def build():
    from xuanpolicy_b200.csrc import build as kbuild
    print(kbuild.build(out=os.path.join(DBG, "libxb200_sts.so"), obj_dir=os.path.join(DBG, "obj_sts"), extra=["-DXB_STEP_TS"]))


def run():
    import ctypes as C
    import numpy as np
    import torch
    from xuanpolicy_b200 import _lib
    from xuanpolicy_b200.configs import build_ppo
    lib = _lib.load()
    agent = build_ppo("Pendulum-v1", parallels=4096, n_steps=128, n_epoch=1, n_minibatch=8, use_obsnorm=True, use_rewnorm=True,
                      shuffle="device", seed=1)
    with torch.cuda.device(agent.device):
        agent._capture()
    ts = torch.zeros(1 + 2 * 30000, dtype=torch.int64, device="cuda")
    lib.xb_debug_set_step_ts.argtypes = [C.c_void_p]
    for _ in range(3):
        agent._rollout_graph.replay()
    torch.cuda.synchronize()
    # the graph captured a NULL stamp buffer: re-capture with the buffer set
    lib.xb_debug_set_step_ts(C.c_void_p(ts.data_ptr()))
    agent._rollout_graph = None
    with torch.cuda.device(agent.device):
        agent._capture()
    ts.zero_()
    torch.cuda.synchronize()
    agent._rollout_graph.replay()
    torch.cuda.synchronize()
    t = ts.cpu().numpy()
    n = int(t[0])
    ev = sorted((int(t[2 + 2 * i]), int(t[1 + 2 * i])) for i in range(min(n, 30000)))
    t0 = ev[0][0]
    print("events:", n)
    # print steps 40..43
    starts = [i for i, (_, tag) in enumerate(ev) if tag == 200]
    lo, hi = starts[40], starts[44]
    base = ev[lo][0]
    for tm, tag in ev[lo:hi]:
        print("%8.2f us  %d" % ((tm - base) / 1e3, tag))
    # mean deltas over all steps
    import collections
    d = collections.defaultdict(list)
    last = {}
    for tm, tag in ev:
        last[tag] = tm
        if tag == 202 and 103 in last: d["tail end(103) -> fwd released(202)"].append(tm - last[103])
        if tag == 202 and 104 in last: d["cta0 env done(104) -> fwd released(202)"].append(tm - last[104])
        if tag == 203 and 202 in last: d["fwd main (202->203)"].append(tm - last[202])
        if tag == 101 and 203 in last: d["fwd done(203) -> step released(101)"].append(tm - last[203])
        if tag == 104 and 101 in last: d["step main cta0 (101->104)"].append(tm - last[101])
        if tag == 102 and 101 in last: d["step released(101) -> last cta in tail(102)"].append(tm - last[101])
        if tag == 103 and 102 in last: d["tail (102->103)"].append(tm - last[102])
        if tag == 105 and 102 in last: d["tail: fence (102->105)"].append(tm - last[102])
        if tag == 106 and 105 in last: d["tail: loads (105->106)"].append(tm - last[105])
        if tag == 107 and 106 in last: d["tail: merge (106->107)"].append(tm - last[106])
        if tag == 103 and 107 in last: d["tail: ticket reset (107->103)"].append(tm - last[107])
        if tag == 201 and 200 in last: d["fwd set-up (200->201)"].append(tm - last[200])
    for k, v in d.items():
        v = np.asarray(v[5:], dtype=np.float64)
        print("%-45s mean %7.2f us  median %7.2f  (n=%d)" % (k, v.mean() / 1e3, np.median(v) / 1e3, v.size))


if __name__ == "__main__":
    build() if sys.argv[1] == "build" else run()
