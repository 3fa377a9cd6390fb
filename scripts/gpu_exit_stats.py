"""Development aid (build with MYO_NVCC_EXTRA_MYO_KERNELS="... -DMYO_EXIT_STATS"): why the Newton solve of the last substep of an
env step stopped - 1: full step without a side change, 2: step below the fp32 floor, 3: stalled, 4: gradient tolerance, 0: iteration cap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200 import BatchSim, Model, _capi
from myochallenge_b200.assets import asset_path

n, T = 8192, 80
m = Model(asset_path("hand/myo_hand_baoding.mjb"))
cfg = m.default_task_cfg(_capi.TASK_BAODING); cfg.task_choice_random = 1
sim = BatchSim(m, n, cfg, device="cuda:0", seed=0)
sim.reset()
rows = []
for t in range(T):
    a = (0.135 * torch.randn(n, sim.nu, device="cuda:0")).clamp(-1, 1)
    sim.step(a)
    if t >= 20:
        sim.mj_step(None, 1)          # one more substep under the same controls through the stage-dumping path
        rows.append(sim.stage("solver_iter")[:, 0].clone())
K = torch.stack(rows).flatten().long()
it, why = K & 255, K >> 8
print("iterations hist", torch.bincount(it).tolist(), "mean", it.float().mean().item())
for r in range(5):
    sel = why == r
    print(f"exit reason {r}: {sel.float().mean().item():.3f} of solves; iterations hist {torch.bincount(it[sel], minlength=7).tolist()}")
