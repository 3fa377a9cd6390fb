"""Development aid: where does the die-reorient step spend its time? Iterations / contacts per substep and per-phase cycles."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from myochallenge_b200 import _capi
if os.environ.get("PROF"):
    _capi._LIB = _capi.bind(os.path.join(ROOT, "scripts", "_prof", "libmyo_prof.so"))
from myochallenge_b200.envs import make_vec_env
from myochallenge_b200.policy import RecurrentPolicy

dev = "cuda:0"
n = 16384
env = make_vec_env("CustomMyoChallengeDieReorientP2-v0", n, device=dev, seed=0, clip_actions=True)
pol = RecurrentPolicy(env.sim.nobs, env.sim.nu, 256, (256, 256), (256, 256), max_batch=n, device=dev)
pol.init_random(0, -2.0); pol.seed(1)
obs = env.reset_device(); h, c = pol.initial_state(n); st = torch.ones(n, dtype=torch.uint8, device=dev)
for t in range(60):
    a, _, _, _ = pol.forward(obs, (h, c), st)
    obs, r, d, tr = env.step_device(a); st = d
sim = env.sim
if os.environ.get("PROF"):
    buf = (C.c_ulonglong * 16)()
    _capi._LIB.myo_debug_profile(buf)
    for t in range(5):
        a, _, _, _ = pol.forward(obs, (h, c), st)
        obs, r, d, tr = env.step_device(a); st = d
    _capi._LIB.myo_debug_profile(buf)
    names = ["tree_fwd", "tendon", "tree_bwd", "mass_bias", "factor", "collision", "constraints", "actuation", "solveM", "newton", "integrate", "nt:hessian", "nt:chol", "nt:chol_solve", "nt:ls+dots", "barrier"]
    tot = sum(buf[:11])
    for k, nm in enumerate(names):
        print(f"  {nm:12s} {buf[k] / (5 * n * 5):10.0f} cyc {100 * buf[k] / tot:5.1f}%")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(10):
    obs, r, d, tr = env.step_device(a)
e1.record(); torch.cuda.synchronize()
print("env step only", e0.elapsed_time(e1) / 10, "ms")
sim.mj_step(None, 1)
it = sim.stage("solver_iter").float(); ne = sim.stage("nefc").float(); nc = sim.stage("ncon").float()
print("iters hist", torch.bincount(it.long().flatten()).tolist(), "ncon hist", torch.bincount(nc.long().flatten()).tolist(), "nefc mean", ne.mean().item(), "max", ne.max().item())
print("die z mean", obs[:, 48].mean().item(), "min", obs[:, 48].min().item(), "status", sim.status())
