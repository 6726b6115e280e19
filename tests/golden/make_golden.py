"""Generates tests/golden/reference_h256_gru.npz by running the REAL reference in place.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference (TimZaman/dotaclient @ 8615b90) ships no golden vectors of its own, so
these fixtures -- outputs of the unmodified reference ``DotaOptimizer.experiences_from_rollout`` and
``DotaOptimizer.train`` (``optimizer.py:328-430,581-689``) on a seeded synthetic rollout -- are what
pins the oracle (``oracle/``) and, through it, the CUDA path.  Nothing is copied from the reference;
only its outputs are recorded.
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from dotaclient_b200.synthetic import make_rollout  # noqa: E402

HEADS = ("enum", "x", "y", "target_unit", "ability")
SEQ_LEN, ROLLOUT_LEN, ROLLOUT_SEED, EPOCHS = 16, 40, 11, 3


def main():
    torch.set_num_threads(1)          # bit-stable reductions
    ref = reference_shim.make_reference_optimizer(seq_len=SEQ_LEN)
    init_sums = {k: float(v.double().sum()) for k, v in ref.policy_base.state_dict().items()}
    data = make_rollout(ROLLOUT_LEN, ROLLOUT_SEED)
    with torch.no_grad():
        seqs = ref.experiences_from_rollout(copy.deepcopy(data))
    out = {"seq_len": SEQ_LEN, "rollout_len": ROLLOUT_LEN, "rollout_seed": ROLLOUT_SEED, "epochs": EPOCHS}
    out["advantages"] = np.stack([s.advantages.numpy() for s in seqs])
    out["returns"] = np.stack([s.returns.numpy() for s in seqs])
    out["values"] = np.stack([s.values.numpy().reshape(-1) for s in seqs])
    out["hidden"] = np.stack([s.hidden.numpy().reshape(-1) for s in seqs])
    for k in HEADS:
        out["old_logp_" + k] = torch.cat([s.log_probs_sel[k] for s in seqs]).numpy()
    # forward at the initial weights on the stacked batch (what train() sees at :619)
    obs = {k: torch.stack([s.observations[k] for s in seqs]) for k in ref.policy_base.INPUT_KEYS}
    hidden = torch.cat([s.hidden for s in seqs], dim=1)
    with torch.no_grad():
        logits, values, _ = ref.policy_base(**obs, hidden=hidden)
    for k in HEADS:
        out["logits_" + k] = logits[k].numpy()
    out["forward_values"] = values.numpy()
    losses, ents, gns = [], [], []
    for ep in range(EPOCHS):
        l, e, g = ref.train(seqs)
        losses.append([float(l[k]) for k in ("loss", "policy_loss", "entropy_loss", "value_loss")])
        ents.append([float(e[k]) for k in HEADS])
        gns.append([float(g["unclipped"]), float(g["clipped"])])
        if ep == 0:
            names = [n for n, _ in ref.policy_base.named_parameters()]
            out["param_names"] = np.array(names)
            out["grad_norms_ep0"] = np.array([float(p.grad.norm(2)) if p.grad is not None else -1.0
                                              for _, p in ref.policy_base.named_parameters()])
            out["grad_rnn_bias_hh_ep0"] = dict(ref.policy_base.named_parameters())["rnn.bias_hh_l0"].grad.numpy().copy()
            out["grad_value_w_ep0"] = dict(ref.policy_base.named_parameters())["affine_value.weight"].grad.numpy().copy()
    out["losses"] = np.array(losses, dtype=np.float64)
    out["entropies"] = np.array(ents, dtype=np.float64)
    out["grad_norms"] = np.array(gns, dtype=np.float64)
    sd = ref.policy_base.state_dict()
    out["init_param_sums"] = np.array([init_sums[k] for k in sd])
    out["final_param_sums"] = np.array([float(v.double().sum()) for v in sd.values()])
    out["final_param_abs_sums"] = np.array([float(v.double().abs().sum()) for v in sd.values()])
    out["final_rnn_bias_hh"] = sd["rnn.bias_hh_l0"].numpy().copy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_h256_gru.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    # GAE known-answer vector, from the reference's advantage_returns itself
    O, _, _ = reference_shim.load()
    r = np.array([1, 1, 1, 0], np.float32)
    v = np.array([.5, .5, .5, 0], np.float32)
    a, q = O.advantage_returns(r, v, 0.98, 0.97)
    rng = np.random.RandomState(3)
    r2 = np.append(rng.randn(300).astype(np.float32), np.float32(0))
    v2 = np.append(rng.randn(300).astype(np.float32), np.float32(0))
    a2, q2 = O.advantage_returns(r2, v2, 0.98, 0.97)
    np.savez_compressed(os.path.join(os.path.dirname(path), "gae_reference.npz"), r=r, v=v, adv=a, ret=q, r2=r2, v2=v2,
                        adv2=a2, ret2=q2)


if __name__ == "__main__":
    main()
