"""Development aid: device-resident stepping of n worlds as k sub-batches on k streams (policy of one sub-batch in the shadow of
another's world kernel) against the single-stream loop. Same spin-up / staggering as bench.py."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from myochallenge_b200.envs import make_vec_env
from myochallenge_b200.policy import RecurrentPolicy
from myochallenge_b200.rollout import PipelinedStepper

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32768)
ap.add_argument("--parts", type=int, nargs="+", default=[1, 2, 4])
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--spinup", type=int, default=200)
args = ap.parse_args()
dev = torch.device("cuda:0")
ENV_ID = "CustomMyoChallengeBaodingP2-v1"
for k in args.parts:
    envs = [make_vec_env(ENV_ID, args.n // k, device=dev, seed=100 * k + i, clip_actions=True) for i in range(k)]
    pol = RecurrentPolicy(envs[0].sim.nobs, envs[0].sim.nu, lstm_hidden=256, pi=(256, 256), vf=(256, 256), max_batch=args.n // k, device=dev)
    pol.init_random(seed=0, log_std_init=-2.0); pol.seed(1)
    st = PipelinedStepper(envs, pol)
    st.reset()
    st.spin_up(args.spinup)
    for _ in range(5):
        st.step()
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        st.step()
    st.join()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(f"{k} sub-batch(es): {ms:.3f} ms per {args.n}-world step = {args.n / ms * 1e3:.4e} env-steps/s, status {[e.sim.status() for e in envs]}", flush=True)
    del st, envs, pol
    torch.cuda.empty_cache()
