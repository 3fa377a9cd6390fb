"""Soak: many env steps of the full-size Baoding batch with the policy in the loop; reports status flags, non-finite counts, episode
statistics and whether throughput drifts (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200.envs import make_vec_env
from myochallenge_b200.policy import RecurrentPolicy

n, steps = 32768, int(os.environ.get("SOAK_STEPS", "1500"))
dev = "cuda:0"
env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=dev, seed=0, clip_actions=True)
pol = RecurrentPolicy(env.sim.nobs, env.sim.nu, 256, (256, 256), (256, 256), max_batch=n, device=dev)
pol.init_random(0, -2.0); pol.seed(1)
obs = env.reset_device(); h, c = pol.initial_state(n); st = torch.ones(n, dtype=torch.uint8, device=dev)
flags = 0; dones = 0; truncs = 0; bad = 0
t0 = time.time(); marks = []
for t in range(steps):
    a, v, lp, _ = pol.forward(obs, (h, c), st)
    obs, r, d, tr = env.step_device(a); st = d
    if t % 100 == 99:
        torch.cuda.synchronize()
        marks.append(time.time() - t0); t0 = time.time()
        flags |= env.sim.status()
        bad += int((~torch.isfinite(obs)).sum()) + int((~torch.isfinite(r)).sum()) + int((~torch.isfinite(h)).sum())
    dones += int(d.sum()) if t % 50 == 0 else 0
    truncs += int(tr.sum()) if t % 50 == 0 else 0
print(f"steps {steps} x {n} worlds: status flags {flags}, non-finite values {bad}, done fraction (sampled) {dones / (n * (steps // 50 + 1)):.4f}, "
      f"truncated fraction {truncs / (n * (steps // 50 + 1)):.4f}")
print("seconds per 100 steps:", [round(x, 2) for x in marks])
