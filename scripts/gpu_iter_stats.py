"""Development aid: per-substep Newton iteration counts of the Baoding worlds - distribution, persistence, and what
grouping worlds by recent iteration count would do to the per-CTA maximum (the lockstep barrier waits for it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200 import BatchSim, Model, _capi
from myochallenge_b200.assets import asset_path

n, T, W = 4096, 60, 13
m = Model(asset_path("hand/myo_hand_baoding.mjb"))
cfg = m.default_task_cfg(_capi.TASK_BAODING); cfg.task_choice_random = 1
sim = BatchSim(m, n, cfg, device="cuda:0", seed=0)
sim.reset()
a = torch.rand(n, sim.nu, device="cuda:0") * 2 - 1
for _ in range(5):
    sim.step(a)
ctrl = torch.rand(n, sim.nu, device="cuda:0")
its = []
for t in range(T):
    if t % 10 == 0:
        ctrl = torch.rand(n, sim.nu, device="cuda:0")
    sim.mj_step(ctrl, 1)
    its.append(sim.stage("solver_iter")[:, 0].float().clone())
K = torch.stack(its)          # [T, n]
print("hist", torch.bincount(K.long().flatten()).tolist(), "mean", K.mean().item())
k0, k1 = K[:-1].flatten(), K[1:].flatten()
print("lag-1 corr", torch.corrcoef(torch.stack([k0, k1]))[0, 1].item())
S = K.reshape(T // 10, 10, n).sum(1)      # per env step totals
print("env-step totals: mean", S.mean().item(), "std", S.std().item(), "lag-1 corr", torch.corrcoef(torch.stack([S[:-1].flatten(), S[1:].flatten()]))[0, 1].item())
def cost(order, Kt):   # sum over groups of W of the per-substep max
    g = Kt[:, order][:, : n // W * W].reshape(Kt.shape[0], -1, W)
    return g.max(2).values.float().mean().item()
rand = torch.arange(n, device="cuda:0")
print("mean of per-group max, natural order:", cost(rand, K[10:]))
for e in range(1, T // 10):
    order = torch.argsort(S[e - 1], stable=True)
    print(f" env step {e}: natural {cost(rand, K[10*e:10*e+10]):.3f}  sorted by previous step's total {cost(order, K[10*e:10*e+10]):.3f}  own mean {K[10*e:10*e+10].mean().item():.3f}")
