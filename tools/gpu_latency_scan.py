#!/usr/bin/env python3
"""Latency of one scheduler round as a function of resident envs per SM: DMB_ENVS_PER_CTA = k, N = 148 * k envs
(k warps on every SM, one round), after a warm-up rollout that desynchronises the episodes.  Run on the GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ROOT)
    from deepmimic_mujoco_b200.env import DPVecEnv
    k = int(sys.argv[1]); rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    E = 148 * k * rounds
    env = DPVecEnv(E, motions=("walk",), seed=0, reward_mode=4, auto_reset=True)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    acts = torch.rand(16, E, env.sim.nu, device="cuda", generator=g) - 0.5
    for t in range(60):
        env.step(acts[t % 16])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(100):
        env.step(acts[t % 16])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    print(f"{k:2d} envs/CTA x {rounds} round(s)  N={E:5d}  {ms*1e3:7.1f} us/step  {E/ms/1e3:6.3f} M env-steps/s  launch {env.sim.launch_info()}")
else:
    for k in (1, 2, 4, 7, 14, 21, 28):
        subprocess.run([sys.executable, __file__, str(k)], env=dict(os.environ, DMB_ENVS_PER_CTA=str(k)))
    subprocess.run([sys.executable, __file__, "1"], env=dict(os.environ, DMB_ENVS_PER_CTA="1", DMB_LOCKSTEP="0"))
    subprocess.run([sys.executable, __file__, "28"], env=dict(os.environ, DMB_ENVS_PER_CTA="28", DMB_LOCKSTEP="0"))
