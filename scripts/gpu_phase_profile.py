"""Development aid: per-phase cycle shares of the world kernel (library built with -DMYO_PROFILE)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from myochallenge_b200 import _capi
_capi._LIB = _capi.bind(os.path.join(ROOT, "scripts", "_prof", "libmyo_prof.so"))
from myochallenge_b200 import BatchSim, Model
from myochallenge_b200.assets import asset_path

NAMES = ["tree_fwd", "tendon", "tree_bwd", "mass_bias", "factor", "collision", "constraints", "actuation", "solveM", "newton", "integrate", "  nt:hessian", "  nt:chol_factor", "  nt:chol_solve", "  nt:linesearch+dots", "barrier_pre_integrate"]

def run(path, kind, n, steps=5, spinup=3):
    m = Model(os.path.join(os.environ["MYO_MODEL_DIR"], path) if os.environ.get("MYO_MODEL_DIR") else asset_path(path))
    cfg = m.default_task_cfg(kind)
    if kind == _capi.TASK_BAODING:
        cfg.task_choice_random = 1
    sim = BatchSim(m, n, cfg, device="cuda:0", seed=0)
    sim.reset()
    a = torch.rand(n, sim.nu, device="cuda:0") * 2 - 1
    for _ in range(spinup):
        sim.step(a)
    buf = (C.c_ulonglong * 16)()
    _capi._LIB.myo_debug_profile(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        sim.step(a)
    e1.record(); torch.cuda.synchronize()
    _capi._LIB.myo_debug_profile(buf)
    tot = sum(buf[:11]) + buf[15]
    sub = steps * n * cfg.frame_skip
    print(f"{path} n={n}: {e0.elapsed_time(e1)/steps:.3f} ms/step; cycles per substep per world = {tot/sub:.0f}")
    sim.mj_step(None, 1)
    it = sim.stage("solver_iter").float(); ne = sim.stage("nefc").float(); nc = sim.stage("ncon").float()
    print(f"   solver iters mean {it.mean():.2f} max {it.max():.0f}; nefc mean {ne.mean():.2f} max {ne.max():.0f}; ncon mean {nc.mean():.2f} max {nc.max():.0f}")
    for k, nm in enumerate(NAMES):
        print(f"   {nm:12s} {buf[k]/sub:10.0f} cyc  {100*buf[k]/tot:5.1f}%")

if __name__ == "__main__":
    if os.environ.get("ONLY_HAND"):
        run("hand/myo_hand_baoding.mjb", _capi.TASK_BAODING, 32768, spinup=int(os.environ.get("SPINUP", "150")))
        sys.exit(0)
    run("finger/myo_finger_v0.mjb", _capi.TASK_POSE, 4096)
    hand = os.path.join(ROOT, "myochallenge_b200", "assets", "hand", "myo_hand_baoding.mjb")
    if os.path.exists(hand):
        run("hand/myo_hand_baoding.mjb", _capi.TASK_BAODING, 4096)
        run("hand/myo_hand_baoding.mjb", _capi.TASK_BAODING, 32768, spinup=int(os.environ.get("SPINUP", "150")))
