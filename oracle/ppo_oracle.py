"""TEST INFRASTRUCTURE - torch (CPU, autograd) restatement of the policy-update half of the reference's training loop.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the product
(``myochallenge_b200/``) never does.

The arithmetic lives in third-party code that is NOT in /root/reference (pinned by requirements.txt:126,135,150:
``sb3-contrib==1.6.2``, ``stable-baselines3==1.6.2``, ``torch==1.12.1``) and is reached from
/root/reference/src/train/trainer.py:67-71 (``agent.learn``) with the hyper-parameters of
/root/reference/docs/summary.md:86-117.  Restated here, in sb3-contrib's own structure (NOT the product's):

* ``split_sequences`` / ``pad``   sb3_contrib/common/recurrent/buffers.py create_sequencers: a sequence starts at every
                                  episode start and at every change of env; sequences are right-padded and masked
* ``evaluate_actions``            sb3_contrib/common/recurrent/policies.py RecurrentActorCriticPolicy.evaluate_actions /
                                  _process_sequence, with ``torch.nn.LSTM`` / ``torch.nn.Linear`` modules built from the
                                  SB3 state-dict keys; DiagGaussianDistribution.log_prob / entropy
* ``ppo_loss``                    sb3_contrib/ppo_recurrent/ppo_recurrent.py RecurrentPPO.train: masked advantage
                                  normalisation, clipped surrogate, (clipped) value loss, entropy bonus
* ``adam_step``                   torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (the real ones: torch is installed)

PARITY UNPINNED against sb3-contrib itself (not installable here).  Pinned: torch's own LSTM / Linear / Adam / clip
implementations are the ones executed, and the golden policy of the reference checkpoint (tests/golden/policy_phase1.npz)
provides real weights and a real observation batch for the gradient tests.
"""
import math

import numpy as np
import torch


def build_modules(sd, dtype=torch.float64):
    """torch modules carrying the SB3 state dict (``lstm_actor``, ``lstm_critic``, ``mlp_extractor.*``, heads)."""
    H = sd["lstm_actor.weight_hh_l0"].shape[1]
    O = sd["lstm_actor.weight_ih_l0"].shape[1]
    mods = {}
    for name in ("lstm_actor", "lstm_critic"):
        lstm = torch.nn.LSTM(O, H, num_layers=1).to(dtype)
        with torch.no_grad():
            for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
                getattr(lstm, k).copy_(torch.as_tensor(sd[f"{name}.{k}"], dtype=dtype))
        mods[name] = lstm
    for net in ("policy_net", "value_net"):
        layers, l = [], 0
        while f"mlp_extractor.{net}.{2 * l}.weight" in sd:
            w = torch.as_tensor(sd[f"mlp_extractor.{net}.{2 * l}.weight"], dtype=dtype)
            lin = torch.nn.Linear(w.shape[1], w.shape[0]).to(dtype)
            with torch.no_grad():
                lin.weight.copy_(w); lin.bias.copy_(torch.as_tensor(sd[f"mlp_extractor.{net}.{2 * l}.bias"], dtype=dtype))
            layers += [lin, torch.nn.ReLU()]
            l += 1
        mods[net] = torch.nn.Sequential(*layers)
    for head in ("action_net", "value_net_head"):
        key = "action_net" if head == "action_net" else "value_net"
        w = torch.as_tensor(sd[f"{key}.weight"], dtype=dtype)
        lin = torch.nn.Linear(w.shape[1], w.shape[0]).to(dtype)
        with torch.no_grad():
            lin.weight.copy_(w); lin.bias.copy_(torch.as_tensor(sd[f"{key}.bias"], dtype=dtype))
        mods[head] = lin
    mods["log_std"] = torch.nn.Parameter(torch.as_tensor(sd["log_std"], dtype=dtype).clone())
    return mods


def named_parameters(mods):
    """(SB3 state-dict key, parameter) pairs."""
    out = [("log_std", mods["log_std"])]
    for name in ("lstm_actor", "lstm_critic"):
        for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            out.append((f"{name}.{k}", getattr(mods[name], k)))
    for net in ("policy_net", "value_net"):
        for i, layer in enumerate(mods[net]):
            if isinstance(layer, torch.nn.Linear):
                out += [(f"mlp_extractor.{net}.{i}.weight", layer.weight), (f"mlp_extractor.{net}.{i}.bias", layer.bias)]
    out += [("action_net.weight", mods["action_net"].weight), ("action_net.bias", mods["action_net"].bias),
            ("value_net.weight", mods["value_net_head"].weight), ("value_net.bias", mods["value_net_head"].bias)]
    return out


def split_sequences(episode_starts):
    """episode_starts [T][B] (bool) -> list of (world b, first step, last step + 1): a new sequence at every episode
    start and at every change of world (env-major flattening, as create_sequencers sees it)."""
    T, B = episode_starts.shape
    seqs = []
    for b in range(B):
        t0 = 0
        for t in range(1, T):
            if episode_starts[t, b]:
                seqs.append((b, t0, t)); t0 = t
        seqs.append((b, t0, T))
    return seqs


def evaluate_actions(mods, obs, actions, episode_starts, h0, c0):
    """obs [T][B][O], actions [T][B][A], episode_starts [T][B], h0 / c0 [2][B][H] (states the rollout started from).
    Sequences are split, right-padded to the longest and run through ``torch.nn.LSTM`` from their start states (zero
    when the sequence begins with an episode start, as ``(1 - episode_start) * state`` makes them); padded steps are
    dropped again.  Returns values [T][B], log_prob [T][B], entropy [T][B]."""
    T, B, O = obs.shape
    dtype = obs.dtype
    es = np.asarray(episode_starts).astype(bool)
    seqs = split_sequences(es)
    L = max(t1 - t0 for _, t0, t1 in seqs)
    n_seq = len(seqs)
    padded = torch.zeros(L, n_seq, O, dtype=dtype)
    for s, (b, t0, t1) in enumerate(seqs):
        padded[: t1 - t0, s] = obs[t0:t1, b]
    latents = []
    for net, name in ((0, "lstm_actor"), (1, "lstm_critic")):
        H = h0.shape[2]
        hs = torch.zeros(1, n_seq, H, dtype=dtype); cs = torch.zeros(1, n_seq, H, dtype=dtype)
        for s, (b, t0, _) in enumerate(seqs):
            if t0 == 0 and not es[0, b]:
                hs[0, s] = h0[net, b]; cs[0, s] = c0[net, b]
        out, _ = mods[name](padded, (hs, cs))              # [L][n_seq][H]
        full = torch.zeros(T, B, H, dtype=dtype)
        pieces = {}
        for s, (b, t0, t1) in enumerate(seqs):
            pieces[(b, t0)] = out[: t1 - t0, s]
        cols = []
        for b in range(B):
            cols.append(torch.cat([pieces[k] for k in sorted(k for k in pieces if k[0] == b)], dim=0))
        full = torch.stack(cols, dim=1)
        latents.append(full)
    latent_pi = mods["policy_net"](latents[0])
    latent_vf = mods["value_net"](latents[1])
    mean = mods["action_net"](latent_pi)
    values = mods["value_net_head"](latent_vf).squeeze(-1)
    dist = action_distribution(mods["log_std"], mean, latent_pi)
    log_prob = dist.log_prob(actions).sum(-1)
    entropy = dist.entropy().sum(-1)
    return values, log_prob, entropy


def action_distribution(log_std, mean, latent_pi):
    """DiagGaussianDistribution (log_std [A]) or StateDependentNoiseDistribution (use_sde=True: log_std [latent_dim][A],
    full_std, use_expln=False, learn_features=False so the latent is detached, epsilon 1e-6; stable_baselines3/common/
    distributions.py proba_distribution)."""
    if log_std.dim() == 1:
        return torch.distributions.Normal(mean, torch.ones_like(mean) * log_std.exp())
    variance = (latent_pi.detach() ** 2) @ (log_std.exp() ** 2)
    return torch.distributions.Normal(mean, torch.sqrt(variance + 1e-6))


def ppo_loss(mods, obs, actions, episode_starts, old_values, old_log_prob, advantages, returns, h0, c0, clip_range=0.2,
             clip_range_vf=None, ent_coef=0.0, vf_coef=0.5, normalize_advantage=True):
    values, log_prob, entropy = evaluate_actions(mods, obs, actions, episode_starts, h0, c0)
    adv = advantages
    if normalize_advantage:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    ratio = torch.exp(log_prob - old_log_prob)
    policy_loss = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
    if clip_range_vf is None:
        values_pred = values
    else:
        values_pred = old_values + torch.clamp(values - old_values, -clip_range_vf, clip_range_vf)
    value_loss = ((returns - values_pred) ** 2).mean()
    entropy_loss = -entropy.mean()
    loss = policy_loss + ent_coef * entropy_loss + vf_coef * value_loss
    with torch.no_grad():
        log_ratio = log_prob - old_log_prob
        stats = dict(policy_loss=policy_loss.item(), value_loss=value_loss.item(), entropy_loss=entropy_loss.item(),
                     approx_kl=((torch.exp(log_ratio) - 1) - log_ratio).mean().item(),
                     clip_fraction=(torch.abs(ratio - 1) > clip_range).double().mean().item(), loss=loss.item())
    return loss, stats


def gradients(sd, batch, dtype=torch.float64, **hyper):
    """state dict + minibatch (dict of numpy arrays, keys as ``ppo_loss`` arguments) -> ({key: grad}, stats)."""
    mods = build_modules(sd, dtype)
    t = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in batch.items() if k != "episode_starts"}
    loss, stats = ppo_loss(mods, t["obs"], t["actions"], batch["episode_starts"], t["old_values"], t["old_log_prob"], t["advantages"],
                           t["returns"], t["h0"], t["c0"], **hyper)
    loss.backward()
    return {k: p.grad.detach().clone() for k, p in named_parameters(mods)}, stats


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-5, max_grad_norm=0.5):
    """One clip_grad_norm_ + torch.optim.Adam step on flat vectors, run by torch's own implementations.
    ``step``: 1-based step number; exp_avg / exp_avg_sq: state before the step. Returns the new (params, exp_avg,
    exp_avg_sq, grad_norm)."""
    p = torch.nn.Parameter(torch.as_tensor(params).clone())
    p.grad = torch.as_tensor(grads).clone()
    norm = torch.nn.utils.clip_grad_norm_([p], max_grad_norm) if max_grad_norm and max_grad_norm > 0 else p.grad.norm()
    opt = torch.optim.Adam([p], lr=lr, betas=betas, eps=eps)
    if step > 1:
        opt.state[p] = dict(step=torch.tensor(float(step - 1)), exp_avg=torch.as_tensor(exp_avg).clone(),
                            exp_avg_sq=torch.as_tensor(exp_avg_sq).clone())
    opt.step()
    st = opt.state[p]
    return p.detach(), st["exp_avg"], st["exp_avg_sq"], float(norm)


def masked_recurrence_loss(sd, batch, dtype=torch.float64, **hyper):
    """The product's formulation (whole sequences, state multiplied by (1 - episode_start) inside the recurrence),
    written with plain torch ops: the CPU tests check it equals the split-and-pad formulation above, value and
    gradient, which is the claim the CUDA design rests on."""
    t = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in batch.items() if k != "episode_starts"}
    P = {k: torch.nn.Parameter(torch.as_tensor(v, dtype=dtype).clone()) for k, v in sd.items()}
    keep = torch.as_tensor(1.0 - np.asarray(batch["episode_starts"], np.float64), dtype=dtype)
    T = keep.shape[0]
    lat = []
    for net, name in ((0, "lstm_actor"), (1, "lstm_critic")):
        h, c = t["h0"][net], t["c0"][net]
        outs = []
        for s in range(T):
            k = keep[s].unsqueeze(1)
            h, c = h * k, c * k
            g = t["obs"][s] @ P[f"{name}.weight_ih_l0"].T + P[f"{name}.bias_ih_l0"] + h @ P[f"{name}.weight_hh_l0"].T + P[f"{name}.bias_hh_l0"]
            i, f, gg, o = g.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        x = torch.stack(outs)
        mlp = "policy_net" if net == 0 else "value_net"
        l = 0
        while f"mlp_extractor.{mlp}.{2 * l}.weight" in P:
            x = torch.relu(x @ P[f"mlp_extractor.{mlp}.{2 * l}.weight"].T + P[f"mlp_extractor.{mlp}.{2 * l}.bias"])
            l += 1
        lat.append(x)
    mean = lat[0] @ P["action_net.weight"].T + P["action_net.bias"]
    values = (lat[1] @ P["value_net.weight"].T + P["value_net.bias"]).squeeze(-1)
    ls = P["log_std"]
    z = (t["actions"] - mean) * torch.exp(-ls)
    log_prob = (-0.5 * z * z - ls - 0.5 * math.log(2 * math.pi)).sum(-1)
    adv = t["advantages"]
    if hyper.get("normalize_advantage", True):
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    cr = hyper.get("clip_range", 0.2)
    ratio = torch.exp(log_prob - t["old_log_prob"])
    policy_loss = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - cr, 1 + cr)).mean()
    cvf = hyper.get("clip_range_vf")
    vp = values if cvf is None else t["old_values"] + torch.clamp(values - t["old_values"], -cvf, cvf)
    value_loss = ((t["returns"] - vp) ** 2).mean()
    entropy_loss = -(0.5 + 0.5 * math.log(2 * math.pi) + ls).sum()
    loss = policy_loss + hyper.get("ent_coef", 0.0) * entropy_loss + hyper.get("vf_coef", 0.5) * value_loss
    loss.backward()
    return loss.item(), {k: p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p) for k, p in P.items()}


def synthetic_batch(sd, T, B, seed=0, start_prob=0.1, dtype=np.float32):
    """A rollout-like minibatch consistent with the policy ``sd`` (old log-probs / values evaluated by the policy itself on
    perturbed parameters, so ratios are near but not equal to 1 and some are clipped)."""
    rng = np.random.default_rng(seed)
    O = sd["lstm_actor.weight_ih_l0"].shape[1]; H = sd["lstm_actor.weight_hh_l0"].shape[1]; A = sd["action_net.weight"].shape[0]
    obs = rng.normal(0, 1, (T, B, O)).clip(-10, 10)
    es = rng.random((T, B)) < start_prob
    h0 = rng.normal(0, 0.3, (2, B, H)); c0 = rng.normal(0, 0.5, (2, B, H))
    mods = build_modules({k: np.asarray(v, np.float64) for k, v in sd.items()})
    with torch.no_grad():
        dummy = torch.zeros(T, B, A, dtype=torch.float64)
        v, _, _ = evaluate_actions(mods, torch.as_tensor(obs), dummy, es, torch.as_tensor(h0), torch.as_tensor(c0))
        latent_pi = mods["policy_net"](_latent(mods, "lstm_actor", obs, es, h0[0], c0[0]))
        mean = mods["action_net"](latent_pi)
        std = action_distribution(mods["log_std"], mean, latent_pi).scale.numpy()
    actions = mean.numpy() + std * rng.normal(0, 1, (T, B, A))
    z = (actions - mean.numpy()) / std
    logp = (-0.5 * z * z - np.log(std) - 0.5 * math.log(2 * math.pi)).sum(-1)
    old_log_prob = logp + rng.normal(0, 0.15, (T, B))      # as if collected by a slightly different policy
    old_values = v.numpy() + rng.normal(0, 0.1, (T, B))
    advantages = rng.normal(0.2, 1.0, (T, B))
    returns = old_values + advantages
    b = dict(obs=obs, actions=actions, episode_starts=es, old_values=old_values, old_log_prob=old_log_prob, advantages=advantages,
             returns=returns, h0=h0, c0=c0)
    return {k: (np.asarray(x, dtype) if k != "episode_starts" else x) for k, x in b.items()}


def _latent(mods, name, obs, es, h0, c0):
    T, B, _ = obs.shape
    h, c = torch.as_tensor(h0).unsqueeze(0), torch.as_tensor(c0).unsqueeze(0)
    outs = []
    for t in range(T):
        k = torch.as_tensor(1.0 - es[t].astype(np.float64)).view(1, B, 1)
        o, (h, c) = mods[name](torch.as_tensor(obs[t]).unsqueeze(0), (h * k, c * k))
        outs.append(o[0])
    return torch.stack(outs)
