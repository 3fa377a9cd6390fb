"""Ad-hoc GPU timing of one full PPO iteration (rollout + update) on Baoding worlds (development aid)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200.envs import make_vec_env
from myochallenge_b200.ppo import RecurrentPPO
from myochallenge_b200.rollout import DeviceVecNormalize, collect_rollouts

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32768)
ap.add_argument("--steps", type=int, default=128)
ap.add_argument("--batch-worlds", type=int, default=2048)
ap.add_argument("--epochs", type=int, default=10)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--use-sde", action="store_true")
ap.add_argument("--minibatches", type=int, default=0, help="time only this many minibatch steps (0: full train())")
args = ap.parse_args()
dev = "cuda:0"
env = make_vec_env("CustomMyoChallengeBaodingP2-v1", args.n, device=dev, seed=0, clip_actions=True)
vn = DeviceVecNormalize(env, gamma=0.99)
agent = RecurrentPPO("MlpLstmPolicy", vn, n_steps=args.steps, batch_size=args.steps * args.batch_worlds, n_epochs=args.epochs, learning_rate=2.5e-5,
                     clip_range=0.2, ent_coef=3e-5, max_grad_norm=0.8, gae_lambda=0.95, precision=args.precision, use_sde=args.use_sde,
                     policy_kwargs=dict(lstm_hidden_size=256, net_arch=[dict(pi=[256, 256], vf=[256, 256])], log_std_init=-2.0))
agent._obs = vn.reset_device().clone()
agent._starts = torch.ones(args.n, dtype=torch.uint8, device=dev)
agent._state = agent.policy.initial_state(args.n)
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(args.iters):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    agent._obs, agent._starts = collect_rollouts(vn, agent.policy, agent.buffer, agent._state, agent._obs, agent._starts)
    e1.record()
    first = agent.update.minibatch_grad(agent.buffer, torch.arange(args.batch_worlds, dtype=torch.int32, device=dev)).tolist()
    print("   first minibatch before any step of this iteration: approx_kl %.5f clip_fraction %.5f policy_loss %.5f" % (first[3], first[4], first[0]),
          "| action std", float(agent.buffer.actions.std()), flush=True)
    if args.minibatches:
        idx = torch.randperm(args.n)[: args.batch_worlds].to(dev, dtype=torch.int32)
        for _ in range(args.minibatches):
            agent.update.minibatch_grad(agent.buffer, idx)
            agent.update.adam_step(agent.update.all_reduce_grad())
        nmb = args.minibatches
        log = dict(zip(("policy_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction", "loss"), agent.update.stats.tolist()))
    else:
        log = agent.update.train(agent.buffer, args.epochs, agent._gen)
        nmb = log["train/n_updates"]
    e2.record(); torch.cuda.synchronize()
    tr, tu = e0.elapsed_time(e1), e1.elapsed_time(e2)
    samples = args.n * args.steps
    print(f"iter {it}: rollout {tr:.1f} ms ({samples / tr * 1e3:.3e} env-steps/s), update {tu:.1f} ms for {nmb} minibatches "
          f"({tu / max(nmb, 1):.2f} ms each, {nmb * args.steps * args.batch_worlds / tu * 1e3:.3e} samples/s), "
          f"whole iteration {samples / (tr + tu) * 1e3:.3e} env-steps/s", flush=True)
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in log.items()}, flush=True)
print("launches: ppo", agent.update.launch_count, "mem GB", torch.cuda.max_memory_allocated() / 1e9)
