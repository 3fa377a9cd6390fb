"""How often do worlds hit the contact capacity (ncon_max)? Distribution of ncon / nefc at steady state (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200 import BatchSim, Model, _capi
from myochallenge_b200.assets import asset_path
from myochallenge_b200.envs import make_task_cfg
from myochallenge_b200.policy import RecurrentPolicy

n = 8192
m = Model(asset_path("hand/myo_hand_baoding.mjb"))
for mode in ("policy", "uniform"):
    cfg = make_task_cfg(m, "CustomMyoChallengeBaodingP2-v1", clip_actions=True)
    sim = BatchSim(m, n, cfg, device="cuda:0", seed=0)
    obs = sim.reset()
    pol = RecurrentPolicy(sim.nobs, sim.nu, 256, (256, 256), (256, 256), max_batch=n, device="cuda:0")
    pol.init_random(0, -2.0); pol.seed(1)
    h, c = pol.initial_state(n); st = torch.ones(n, dtype=torch.uint8, device="cuda:0")
    hist = torch.zeros(32, dtype=torch.long, device="cuda:0")
    over = 0; tot = 0
    for t in range(260):
        if mode == "policy":
            a, _, _, _ = pol.forward(obs, (h, c), st)
        else:
            a = torch.rand(n, sim.nu, device="cuda:0") * 2 - 1
        obs, r, d, tr = sim.step(a); st = d
        if t >= 200 and t % 10 == 0:
            sim.status()                                   # clear
            sim.mj_step(None, 1)                           # one more substep with zero ctrl, stage buffers filled
            nc = sim.stage("ncon")[:, 0].long()
            hist += torch.bincount(nc.clamp(max=31), minlength=32)
            flags = sim.stage("status")[:, 0]
            over += int(((flags & 2) != 0).sum()); tot += n
    print(mode, "ncon histogram", hist.tolist()[:18], "overflow fraction of world-substeps", over / max(tot, 1), flush=True)
