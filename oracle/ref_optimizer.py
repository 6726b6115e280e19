"""CPU restatement of the reference optimizer step (GAE, experience prep, PPO train step).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Same arithmetic, same operation
order as ``optimizer.py`` so that at (hidden 256, GRU) results are bit-identical to the
reference executed in this container (``tests/test_oracle_vs_reference.py``); generalised
to any hidden width and to an LSTM state tuple, which the reference cannot express.
"""
import numpy as np
import torch
from scipy.signal import lfilter

from .ref_policy import EPS, INPUT_KEYS, OUTPUT_KEYS, RefPolicy, masked_softmax

GAMMA = 0.98      # optimizer.py:421
LAMBDA = 0.97     # optimizer.py:421
E_CLIP = 0.1      # optimizer.py:229
MAX_GRAD_NORM = 0.5   # optimizer.py:204


def discount(x, gamma):
    """Reverse first-order IIR ``y_t = x_t + gamma*y_{t+1}``; float64 accumulate, fp32 out (``optimizer.py:53-54``)."""
    return lfilter([1], [1, -gamma], x[::-1], axis=0)[::-1].astype(np.float32)


def advantage_returns(rewards, values, gamma=GAMMA, lam=LAMBDA):
    """GAE-lambda advantages and rewards-to-go (``optimizer.py:57-64``).

    ``rewards``/``values`` carry one trailing bootstrap element (0 for terminated rollouts,
    ``optimizer.py:417-420``).  deltas are formed in fp32, the scans run in float64.
    """
    deltas = rewards[:-1] + gamma * values[1:] - values[:-1]
    return discount(deltas, gamma * lam), discount(rewards, gamma)[:-1]


class RefSequence:
    """Plain record, fields as ``optimizer.py:176-190``."""

    def __init__(self, **kw):
        self.advantages = None
        self.returns = None
        self.__dict__.update(kw)


def _pad_time(t, pad):
    """Zero-pad dim 0 (time) at the end, as the ``dim_pad`` table of ``optimizer.py:367-380``."""
    if pad == 0:
        return t
    spec = [0, 0] * (t.dim() - 1) + [0, pad]
    return torch.nn.functional.pad(t, spec, mode="constant", value=0).detach()


def experiences_from_rollout(policy, data, seq_len):
    """Slice a rollout into ``seq_len`` chunks with carried hidden state (``optimizer.py:328-430``).

    No-grad chunk-wise forward, old log-probs of the taken actions per head, whole-rollout GAE
    (padding of the last chunk is inside the scan), then split per chunk.
    """
    obs, masks, actions, rewards = data["observations"], data["masks"], data["actions"], data["rewards"]
    L = rewards.shape[0]
    hidden = policy.init_hidden()
    seqs, values, reward_sums = [], [], []
    with torch.no_grad():
        for i1 in range(0, L, seq_len):
            pad = max(0, seq_len - (L - i1))
            i2 = i1 + seq_len - pad
            s_obs = {k: _pad_time(v[i1:i2], pad) for k, v in obs.items()}
            s_masks = {k: _pad_time(v[i1:i2], pad) for k, v in masks.items()}
            s_actions = {k: _pad_time(v[i1:i2], pad) for k, v in actions.items()}
            s_rewards = rewards[i1:i2]
            if pad:
                s_rewards = np.pad(s_rewards, ((0, pad), (0, 0)), mode="constant")
            hidden_in = hidden
            logits, s_values, hidden = policy.sequence(**s_obs, hidden=hidden_in)
            old = {}
            for k in logits:                                           # optimizer.py:387-390
                lp = masked_softmax(logits[k], s_masks[k].unsqueeze(0))
                old[k] = torch.masked_select(lp, s_actions[k]).detach()
            values.append(s_values)
            reward_sums.append(np.sum(s_rewards, axis=1).ravel())      # optimizer.py:397
            hid = tuple(h.detach() for h in hidden_in) if isinstance(hidden_in, tuple) else hidden_in.detach()
            seqs.append(RefSequence(game_id=data.get("game_id"), weight_version=data.get("weight_version"),
                                    team_id=data.get("team_id"), observations=s_obs, actions=s_actions,
                                    masks=s_masks, values=s_values.detach(), rewards=s_rewards,
                                    hidden=hid, log_probs_sel=old))
    v = np.append(torch.cat(values).cpu().numpy().ravel(), np.array(0.0, dtype=np.float32))   # optimizer.py:417-418
    r = np.append(np.concatenate(reward_sums), np.array(0.0, dtype=np.float32))               # optimizer.py:419-420
    adv, ret = advantage_returns(rewards=r, values=v, gamma=GAMMA, lam=LAMBDA)
    for s, a, q in zip(seqs, np.split(adv, len(seqs)), np.split(ret, len(seqs))):
        s.advantages = torch.from_numpy(a)
        s.returns = torch.from_numpy(q)
    return seqs


def mean_gradient_norm(params):
    """Mean over params-with-grad of the per-tensor L2 norm (``optimizer.py:691-695``)."""
    return torch.stack([p.grad.data.norm(2) for p in params if p.grad is not None]).mean()


def stack_batch(experiences):
    """Stacks a list of sequences into batch tensors (``optimizer.py:587-615``)."""
    adv = torch.stack([e.advantages for e in experiences])
    ret = torch.stack([e.returns for e in experiences]).detach()
    if isinstance(experiences[0].hidden, tuple):
        hidden = tuple(torch.cat([e.hidden[i] for e in experiences], dim=1).detach() for i in range(2))
    else:
        hidden = torch.cat([e.hidden for e in experiences], dim=1).detach()
    actions = {k: torch.stack([e.actions[k] for e in experiences]) for k in OUTPUT_KEYS}
    masks = {k: torch.stack([e.masks[k] for e in experiences]) for k in OUTPUT_KEYS}
    obs = {k: torch.stack([e.observations[k] for e in experiences]) for k in INPUT_KEYS}
    old = {k: torch.cat([e.log_probs_sel[k] for e in experiences]).detach() for k in OUTPUT_KEYS}
    return adv, ret, hidden, actions, masks, obs, old


def ppo_loss(logits, values, actions, masks, old, adv_raw, returns, entropy_coef, vf_coef, e_clip=E_CLIP):
    """The PPO clipped-surrogate / entropy / value loss of ``optimizer.py:587-665`` on stacked tensors."""
    adv = ((adv_raw - adv_raw.mean()) / (adv_raw.std() + EPS)).detach()      # optimizer.py:588-589
    policy_loss, entropies = {}, {}
    for k in logits:
        step = actions[k].sum(dim=-1) != 0                                   # optimizer.py:626
        if step.sum() == 0:
            policy_loss[k] = torch.zeros([])
            entropies[k] = torch.zeros([])
            continue
        lp = masked_softmax(logits[k], masks[k])
        lp_sel = torch.masked_select(lp, actions[k])
        a_sel = adv[step]
        ratio = torch.exp(lp_sel - old[k])
        surr1 = ratio * a_sel.view(-1)
        surr2 = torch.clamp(ratio, 1.0 - e_clip, 1.0 + e_clip) * a_sel
        policy_loss[k] = -torch.min(surr1, surr2).mean()
        n_actions = step.sum()
        lp_m = torch.masked_select(lp, masks[k])
        entropies[k] = -(torch.exp(lp_m) * lp_m).sum() / n_actions          # optimizer.py:643-646
    p_loss = torch.stack(list(policy_loss.values())).mean()                  # optimizer.py:649-650
    e_loss = -entropy_coef * torch.stack(list(entropies.values())).sum() if entropy_coef > 0 else torch.tensor(0.0)
    v_loss = vf_coef * (0.5 * (returns - values.squeeze(-1)).pow(2).mean()) if vf_coef > 0 else torch.tensor(0.0)
    loss = p_loss + e_loss + v_loss
    return loss, p_loss, e_loss, v_loss, entropies


class RefOptimizer:
    """Holds a policy + Adam and performs ``DotaOptimizer.train`` (``optimizer.py:581-689``) on CPU."""

    def __init__(self, policy, seq_len, learning_rate=5e-5, entropy_coef=5e-4, vf_coef=0.5, forward_module=None):
        self.policy_base = policy
        self.policy = forward_module if forward_module is not None else policy   # DDP wrapper when distributed
        self.seq_len = seq_len
        self.entropy_coef = entropy_coef
        self.vf_coef = vf_coef
        self.e_clip = E_CLIP
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=learning_rate)  # optimizer.py:275

    def experiences_from_rollout(self, data):
        return experiences_from_rollout(self.policy_base, data, self.seq_len)

    def loss_only(self, experiences):
        adv, ret, hidden, actions, masks, obs, old = stack_batch(experiences)
        logits, values, _ = self.policy(**obs, hidden=hidden)
        out = ppo_loss(logits, values, actions, masks, old, adv, ret, self.entropy_coef, self.vf_coef, self.e_clip)
        return out, logits, values

    def train(self, experiences):
        (loss, p_loss, e_loss, v_loss, entropies), _, _ = self.loss_only(experiences)
        if torch.isnan(loss):                                                # optimizer.py:667-669
            raise ValueError("loss={}, policy_loss={}, entropy_loss={}, value_loss={}".format(loss, p_loss, e_loss, v_loss))
        self.optimizer.zero_grad()
        loss.backward()
        params = list(self.policy.parameters())
        gn = mean_gradient_norm(params)
        torch.nn.utils.clip_grad_norm_(params, MAX_GRAD_NORM)
        gn_clipped = mean_gradient_norm(params)
        if torch.isnan(gn):
            raise ValueError("grad_norm={}".format(gn))
        self.optimizer.step()
        losses = {"loss": loss, "policy_loss": p_loss, "entropy_loss": e_loss, "value_loss": v_loss}
        return losses, entropies, {"unclipped": gn, "clipped": gn_clipped}


def make_ref_optimizer(hidden_size=256, cell="gru", seq_len=16, seed=7, **kw):
    """``torch.manual_seed(7); Policy()`` as ``optimizer.py:34,220``."""
    torch.manual_seed(seed)
    return RefOptimizer(RefPolicy(hidden_size=hidden_size, cell=cell), seq_len=seq_len, **kw)
