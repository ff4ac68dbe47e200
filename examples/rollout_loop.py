"""A jax_ppo-style rollout loop on the B200 env (reference agents/jax_ppo.py:1141-1236: policy -> step_env_wrapped ->
storage, num_ppo_steps per rollout), everything device-resident:

    python examples/rollout_loop.py [--envs 4096] [--rollouts 4] [--steps 128]

One launch per env step (`step_observe_device`: fused step + auto-reset + RGB frame), one small launch for the episode
statistics (`rollout_stats.EpisodeStatistics`, the statistics half of step_env_wrapped), observations written straight
into the rollout storage's frame of the step.  The "policy" is a stand-in (random actions drawn on the device); a real
agent reads `obs` (N,64,64,3) uint8 / float32 and writes `actions` (N,3) int32."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gym_cellular_automata_b200.forest_fire.bulldozer import AdvancedForestFireBulldozerEnv
from gym_cellular_automata_b200.rollout_stats import EpisodeStatistics


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--rollouts", type=int, default=4)
    ap.add_argument("--steps", type=int, default=128)      # num_ppo_steps (agents/args.py:59)
    ap.add_argument("--obs", default="rgb_u8", choices=["rgb_u8", "rgb_f32"])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    N = a.envs
    env = AdvancedForestFireBulldozerEnv(64, 64, key=1, num_envs=N, speed_move=0.12 * 4, speed_act=0.03 * 4, use_hidden=True,
                                         substeps=1, seed=0, hidden="device", obs_mode=a.obs, auto_reset=True,
                                         balance_every=8, device=dev)
    (obs, ctx), info = env.reset()
    stats = EpisodeStatistics(N, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    rewards = torch.empty((a.steps, N), dtype=torch.float32, device=dev)
    dones = torch.empty((a.steps, N), dtype=torch.uint8, device=dev)
    for r in range(a.rollouts):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(a.steps):
            # policy stand-in: (move 0..8, shoot 0..1, extension id 0..2)
            actions = torch.stack([torch.randint(0, 9, (N,), device=dev, generator=gen),
                                   torch.randint(0, 2, (N,), device=dev, generator=gen),
                                   torch.randint(0, 3, (N,), device=dev, generator=gen)], 1).to(torch.int32)
            out, obs = env.step_observe_device(actions)            # ONE launch: step + auto-reset + frame
            stats.update(actions, out.step_reward, out.terminated, out.obs_night)
            rewards[t].copy_(out.reward)
            dones[t].copy_(out.terminated)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        fin = int(stats.amount_finished.item())
        print(f"rollout {r}: {a.steps} steps x {N} envs in {dt * 1e3:.1f} ms = {a.steps * N / dt / 1e6:.1f} M env-steps/s; "
              f"episodes finished so far {fin}; mean reward {float(rewards.mean()):.4f}")


if __name__ == "__main__":
    main()
