"""Where a collect_rollouts step spends its time (development aid): CUDA-event brackets around the pieces."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200.envs import make_vec_env
from myochallenge_b200.ppo import RecurrentPPO
from myochallenge_b200 import rollout as R

n, T = 32768, 32
dev = "cuda:0"
env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=dev, seed=0, clip_actions=True)
vn = R.DeviceVecNormalize(env, gamma=0.99)
agent = RecurrentPPO("MlpLstmPolicy", vn, n_steps=T, batch_size=T * 2048, n_epochs=1,
                     policy_kwargs=dict(lstm_hidden_size=256, net_arch=[dict(pi=[256, 256], vf=[256, 256])], log_std_init=-2.0))
acc = collections.defaultdict(list)

def timed(name, fn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        acc[name].append((e0, e1))
        return r
    return w

env.step_device = timed("world_kernel(step_device)", env.step_device)
agent.policy.forward = timed("policy.forward", agent.policy.forward)
vn.obs_rms.update = timed("obs_rms.update", vn.obs_rms.update)
agent.buffer.add = timed("buffer.add", agent.buffer.add)
agent.buffer.put_obs = timed("buffer.put_obs", agent.buffer.put_obs)
vn.step_device = timed("vn.step_device(total)", vn.step_device)
obs = vn.reset_device().clone()
st = torch.ones(n, dtype=torch.uint8, device=dev)
state = agent.policy.initial_state(n)
for it in range(2):
    acc.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    obs, st = R.collect_rollouts(vn, agent.policy, agent.buffer, state, obs, st)
    e1.record(); torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1) / T:.3f} ms per step")
    for k, v in acc.items():
        print(f"   {k:32s} {sum(a.elapsed_time(b) for a, b in v) / T:.3f} ms per step ({len(v)} calls)")
