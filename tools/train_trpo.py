#!/usr/bin/env python3
"""Example: TRPO on the batched CUDA env -- the GPU-resident counterpart of
`python3 trpo.py --task train` in the reference (trpo.py:438-490).  One process per GPU:
  python tools/train_trpo.py --envs 1024 --horizon 32 --iters 20
  torchrun --nproc-per-node 8 tools/train_trpo.py ...        (gradients averaged over NCCL)
"""
import argparse, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepmimic_mujoco_b200.env import DPVecEnv
from deepmimic_mujoco_b200.policy import MlpPolicy
from deepmimic_mujoco_b200.rollout import SegmentGenerator, add_vtarg_and_adv
from deepmimic_mujoco_b200.trpo import TRPO

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1024); ap.add_argument("--horizon", type=int, default=32)
ap.add_argument("--iters", type=int, default=20); ap.add_argument("--motion", default="walk")
ap.add_argument("--reward-mode", type=int, default=4); ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--vf-batch", type=int, default=4096)
ap.add_argument("--pretrained_weight_path", default=None,
                help="TensorFlow checkpoint saved by the reference (trpo.py:207-208, 516); read without TensorFlow")
ap.add_argument("--checkpoint_dir", default=None, help="save the policy there in the reference's TensorFlow checkpoint "
                "format every --save_per_iter iterations (trpo.py:220-224, 494, 514) -- loadable by the reference")
ap.add_argument("--save_per_iter", type=int, default=100)
a = ap.parse_args()
world, rank, lrank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
env = DPVecEnv(a.envs, motions=(a.motion,), seed=a.seed, first_env_id=rank * a.envs, reward_mode=a.reward_mode)
pi = MlpPolicy(seed=a.seed + 10000 * rank)           # workerseed = seed + 10000 * rank (trpo.py:341)
if a.pretrained_weight_path:
    pi.load_tf_checkpoint(a.pretrained_weight_path, "pi")
learner = TRPO(pi, vf_batch=a.vf_batch)              # broadcasts rank 0's parameters
gen = SegmentGenerator(pi, env, a.horizon)
t0, steps = time.time(), 0
for it in range(a.iters):
    seg = next(gen)
    add_vtarg_and_adv(seg, learner.gamma, learner.lam)
    if rank == 0 and a.checkpoint_dir and it % a.save_per_iter == 0:
        os.makedirs(a.checkpoint_dir, exist_ok=True)
        pi.save_tf_checkpoint(os.path.join(a.checkpoint_dir, f"trpo-{a.motion}-{a.seed}"))
    st = learner.update(seg)
    steps += a.envs * a.horizon * world
    if rank == 0:
        n = len(seg["ep_lens"])
        print(f"iter {it:3d} steps {steps:9d}  rew/step {seg['rew'].mean().item():.4f}  EpLenMean {seg['ep_lens'].float().mean().item() if n else 0:.1f} "
              f"EpRewMean {seg['ep_rets'].mean().item() if n else 0:.2f}  kl {st['meankl']:.4f} surr {st['surrgain']:.4f} "
              f"step {st['stepsize']:.3f} vferr {st['vferr']:.3f} ev_tdlam_before {st.get('ev_tdlam_before', float('nan')):.3f}  {steps/(time.time()-t0)/1e3:.0f}k steps/s", flush=True)
env.close()
if world > 1:
    dist.destroy_process_group()
