"""One PPO minibatch (gradient + Adam) on synthetic rollout data at the bench's sizes, for an ncu launch list (development aid)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200.policy import RecurrentPolicy
from myochallenge_b200.ppo import PPOUpdate
from myochallenge_b200.rollout import RecurrentRolloutBuffer

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--worlds", type=int, default=2048)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
dev = "cuda:0"
pol = RecurrentPolicy(86, 39, 256, (256, 256), (256, 256), max_batch=a.worlds, device=dev)
pol.init_random(0, -2.0)
upd = PPOUpdate(pol, a.steps, a.worlds, "bf16", learning_rate=2.5e-5, ent_coef=3e-5, max_grad_norm=0.8)
buf = RecurrentRolloutBuffer(a.steps, a.worlds, 86, 39, 256, dev)
g = torch.Generator(device=dev).manual_seed(0)
buf.observations.normal_(generator=g); buf.actions.normal_(generator=g).mul_(0.2); buf.values.normal_(generator=g); buf.log_probs.normal_(generator=g).mul_(0.1).add_(30)
buf.advantages.normal_(generator=g); buf.returns.copy_(buf.values + buf.advantages)
buf.episode_starts.copy_((torch.rand(a.steps, a.worlds, device=dev, generator=g) < 0.01).to(torch.uint8))
idx = torch.arange(a.worlds, dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(a.reps):
    e0.record()
    upd.minibatch_grad(buf, idx); upd.adam_step()
    e1.record(); torch.cuda.synchronize()
    print(f"rep {r}: {e0.elapsed_time(e1):.2f} ms, launches so far {upd.launch_count}", flush=True)
