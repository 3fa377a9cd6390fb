"""Ad-hoc GPU timing of the world kernel (development aid, not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200 import BatchSim, Model, _capi
from myochallenge_b200.assets import asset_path

def run(path, kind, n, steps=20, spinup=3):
    m = Model(os.path.join(os.environ["MYO_MODEL_DIR"], path) if os.environ.get("MYO_MODEL_DIR") else asset_path(path))
    cfg = m.default_task_cfg(kind)
    if kind == _capi.TASK_BAODING:
        cfg.task_choice_random = 1
    sim = BatchSim(m, n, cfg, device="cuda:0", seed=0)
    sim.reset()
    a = torch.rand(n, sim.nu, device="cuda:0") * 2 - 1
    for _ in range(spinup):
        sim.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sim.step(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{path} n={n} {ms:.3f} ms/step {n / ms * 1e3:.3e} env-steps/s status={sim.status()} info={sim.launch_info()} done_frac={sim.done.float().mean().item():.3f}", flush=True)

if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="all")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--spinup", type=int, default=3, help="untimed steps first (150+: steady-state episode phases)")
    ap.add_argument("--lib", default="", help="alternate build of the library (development experiments)")
    args = ap.parse_args()
    if args.lib:
        _capi._LIB = _capi.bind(args.lib)
    if args.model in ("all", "finger"):
        for n in ([args.n] if args.n else [4096, 65536]):
            run("finger/myo_finger_v0.mjb", _capi.TASK_POSE, n, args.steps, args.spinup)
    if args.model in ("all", "hand"):
        for n in ([args.n] if args.n else [4096, 32768]):
            run("hand/myo_hand_baoding.mjb", _capi.TASK_BAODING, n, args.steps, args.spinup)
