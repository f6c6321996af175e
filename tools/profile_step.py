"""Profiling harness (run under ncu with --profile-from-start off).

    python tools/profile_step.py step   [--workload c2]   one PPO iteration of bench.py's workload inside the capture range
    python tools/profile_step.py gae_c4 [--envs 1048576]  GAE micro-benchmark shape, both variants
    python tools/profile_step.py kernels                  each hand-written kernel once at the C2 shapes (for --set full)
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["step", "gae_c4", "kernels"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--T", type=int, default=2048)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    cudart = torch.cuda.cudart()
    if args.mode == "gae_c4":
        from xuanpolicy_b200 import ops
        T, N = args.T, args.envs
        gen = torch.Generator(device="cuda").manual_seed(1234)
        rew = torch.randn((T, N), device="cuda", generator=gen)
        val = torch.randn((T, N), device="cuda", generator=gen)
        term = (torch.rand((T, N), device="cuda", generator=gen) < 1 / 200).float()
        boot = torch.randn(N, device="cuda", generator=gen)
        adv, ret = torch.empty_like(rew), torch.empty_like(rew)
        for v in ("ldg", "tma"):
            ops.gae(rew, val, term, boot, adv, ret, 0.99, 0.95, variant=v)
        torch.cuda.synchronize()
        cudart.cudaProfilerStart()
        for v in ("ldg", "tma", "tma"):
            ops.gae(rew, val, term, boot, adv, ret, 0.99, 0.95, variant=v)
        torch.cuda.synchronize()
        cudart.cudaProfilerStop()
        return
    wl = bench.WORKLOADS[args.workload]
    agent = bench.build_agent(wl, 1, "device", False)
    for _ in range(3):
        agent.train(agent.n_steps)
    torch.cuda.synchronize()
    if args.mode == "step":
        cudart.cudaProfilerStart()
        agent.train(agent.n_steps)
        torch.cuda.synchronize()
        cudart.cudaProfilerStop()
    else:
        flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
        peak, _ = bench.measured_peaks()
        def once(fn, flush_, iters=1, warm=0):     # one launch per kernel inside the capture range
            flush_.add_(1)
            fn()
            torch.cuda.synchronize()
            return 1.0, 1.0
        bench.time_kernel = once
        cudart.cudaProfilerStart()
        bench.kernel_rooflines(agent, flush, peak, {"updates": 64}, with_c4=False, world=1)
        torch.cuda.synchronize()
        cudart.cudaProfilerStop()


if __name__ == "__main__":
    main()
