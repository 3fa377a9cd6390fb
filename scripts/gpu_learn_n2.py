"""Multi-GPU check of RecurrentPPO.learn (run under torchrun, one rank per GPU): after every rollout all ranks must hold the same
VecNormalize moments and the same parameters (per-rank bases in the moment merge, rank-consistent target_kl stop)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from myochallenge_b200.envs import make_vec_env
from myochallenge_b200.ppo import RecurrentPPO
from myochallenge_b200.rollout import DeviceVecNormalize
from myochallenge_b200.vec_env import rank_seed

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
n = 2048
env = make_vec_env("CustomMyoChallengeBaodingP2-v1", n, device=dev, seed=rank_seed(0, rank), clip_actions=True, enable_rsi=True, rsi_probability=0.5)
vn = DeviceVecNormalize(env, gamma=0.99)
agent = RecurrentPPO("MlpLstmPolicy", vn, n_steps=16, batch_size=16 * 512, n_epochs=3, learning_rate=3e-4, target_kl=0.002, seed=0,
                     policy_kwargs=dict(lstm_hidden_size=128, net_arch=[dict(pi=[64], vf=[64])], log_std_init=-1.0))

def check(tag):
    for name, t in (("obs_rms", vn.obs_rms.state), ("ret_rms", vn.ret_rms.state), ("params", agent.update.params)):
        got = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(got, t.contiguous())
        for k in range(1, world):
            assert torch.equal(got[0], got[k]), f"{tag}: {name} differs between rank 0 and rank {k}: {float((got[0] - got[k]).abs().max())}"
        assert torch.isfinite(t).all(), f"{tag}: {name} not finite"

logs = []
agent.learn(3 * 16 * n * world, callback=lambda a, log: (check(f"rollout {len(logs)}"), logs.append(log), True)[-1])
check("end")
if rank == 0:
    print("OK", world, "ranks;", len(logs), "rollouts; n_updates per rollout", [l["train/n_updates"] for l in logs], "obs count", float(vn.obs_rms.count),
          "expected", 1e-4 + world * n * (1 + 3 * 16), "approx_kl", [round(l["train/approx_kl"], 5) for l in logs])
    assert abs(float(vn.obs_rms.count) - (1e-4 + world * n * (1 + 3 * 16))) < 1.0
dist.destroy_process_group()
