"""CPU restatement of the reference ``Policy`` network, parametrised on width and cell.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference hard-codes a 256-wide GRU (``policy.py:65-66,69-75,78``).  BASELINE.json
sweeps hidden in {128,256,512} and names an LSTM, which the reference cannot express,
so this module restates ``policy.py:51-178`` with ``256 -> hidden_size`` and
``nn.GRU -> nn.GRU | nn.LSTM``.  At (256, 'gru') it has the identical ``state_dict``
(34 keys, same order) and is verified bit-identical to the reference forward by
``tests/test_oracle_vs_reference.py``.

Everything runs in stock torch CPU fp32 -- that IS the reference's arithmetic
(``torch==1.0.0`` CPU in ``docker/Dockerfile:17``; torch 2.11 here).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

EPS = float(np.finfo(np.float32).eps)  # policy.py:15

# (attribute suffix, observation key, units) in the order the reference concatenates them
# (policy.py:99-131).  MAX_UNITS = 1+5+16+16+1+1 (policy.py:45).
UNIT_GROUPS = (
    ("ah", "allied_heroes", 1),
    ("eh", "enemy_heroes", 5),
    ("anh", "allied_nonheroes", 16),
    ("enh", "enemy_nonheroes", 16),
    ("ath", "allied_towers", 1),
    ("eth", "enemy_towers", 1),
)
INPUT_KEYS = ("env",) + tuple(g[1] for g in UNIT_GROUPS)  # policy.py:48-49
ACTION_OUTPUT_COUNTS = {"enum": 4, "x": 9, "y": 9, "target_unit": 40, "ability": 3}  # policy.py:46
OUTPUT_KEYS = tuple(ACTION_OUTPUT_COUNTS)


class RefPolicy(nn.Module):
    """Restates ``policy.py:36-167``; module creation order == reference (so seeded init matches)."""

    def __init__(self, hidden_size=256, cell="gru"):
        super().__init__()
        assert cell in ("gru", "lstm")
        self.hidden_size = hidden_size
        self.cell = cell
        H = hidden_size
        self.affine_env = nn.Linear(3, 128)                       # policy.py:54
        self.affine_unit_basic_stats = nn.Linear(12, 128)         # policy.py:56
        for suffix, _, _ in UNIT_GROUPS:                          # policy.py:58-63
            setattr(self, "affine_unit_" + suffix, nn.Linear(128, 128))
        self.affine_pre_rnn = nn.Linear(896, H)                   # policy.py:65
        rnn_cls = nn.GRU if cell == "gru" else nn.LSTM
        self.rnn = rnn_cls(input_size=H, hidden_size=H, num_layers=1, batch_first=True)  # policy.py:66
        self.affine_head_enum = nn.Linear(H, 4)                   # policy.py:69
        self.affine_move_x = nn.Linear(H, 9)                      # policy.py:70
        self.affine_move_y = nn.Linear(H, 9)                      # policy.py:71
        self.affine_unit_attention = nn.Linear(H, 128)            # policy.py:73
        self.affine_head_ability = nn.Linear(H, 3)                # policy.py:74
        self.affine_value = nn.Linear(H, 1)                       # policy.py:75

    def init_hidden(self):                                        # policy.py:77-78
        h = torch.zeros([1, 1, self.hidden_size], dtype=torch.float32)
        if self.cell == "lstm":
            return (h, torch.zeros_like(h))
        return h

    def sequence(self, hidden, **obs):                            # policy.py:86-90
        return self(**{k: v.unsqueeze(0) for k, v in obs.items()}, hidden=hidden)

    def forward(self, env, allied_heroes, enemy_heroes, allied_nonheroes, enemy_nonheroes,
                allied_towers, enemy_towers, hidden):
        groups = dict(allied_heroes=allied_heroes, enemy_heroes=enemy_heroes,
                      allied_nonheroes=allied_nonheroes, enemy_nonheroes=enemy_nonheroes,
                      allied_towers=allied_towers, enemy_towers=enemy_towers)
        emb, emb_max = {}, {}
        for suffix, key, _ in UNIT_GROUPS:                        # policy.py:99-127
            basic = F.relu(self.affine_unit_basic_stats(groups[key]))
            emb[suffix] = getattr(self, "affine_unit_" + suffix)(basic)   # (b, s, units, 128)
            emb_max[suffix] = torch.max(emb[suffix], dim=2)[0]            # (b, s, 128)
        # Reference quirk, REQUIRED for parity: the enemy-tower max is taken from the
        # enemy-nonhero embedding (policy.py:127), so eth only reaches the attention head.
        emb_max["eth"] = torch.max(emb["enh"], dim=2)[0]
        unit_embedding = torch.cat([emb[s] for s, _, _ in UNIT_GROUPS], dim=2)    # policy.py:130-131
        unit_embedding = unit_embedding.transpose(3, 2)                           # policy.py:132
        x = torch.cat([F.relu(self.affine_env(env))] + [emb_max[s] for s, _, _ in UNIT_GROUPS], dim=2)
        x = F.relu(self.affine_pre_rnn(x))                                        # policy.py:138
        x, hidden = self.rnn(x, hidden)                                           # policy.py:141
        attention = self.affine_unit_attention(x).unsqueeze(2)                    # policy.py:144-145
        # Op creation order matters for bit-exact parity: autograd sums the five head gradients into
        # the rnn output in reverse creation order, so create them exactly as policy.py:148-155 does.
        move_x = self.affine_move_x(x)                                            # policy.py:148
        move_y = self.affine_move_y(x)                                            # policy.py:149
        head_enum = self.affine_head_enum(x)                                      # policy.py:150
        target_unit = torch.matmul(attention, unit_embedding).squeeze(2)          # policy.py:152-153
        ability = self.affine_head_ability(x)                                     # policy.py:154
        logits = {"enum": head_enum, "x": move_x, "y": move_y, "target_unit": target_unit, "ability": ability}
        value = self.affine_value(x)                                              # policy.py:155
        return logits, value, hidden


def masked_softmax(logits, mask, dim=2):
    """Log-probs normalised over ``mask`` only, no max-subtraction (``policy.py:169-178``)."""
    masked_exp = torch.exp(logits).clone()
    masked_exp[~mask] = 0.0
    return logits - torch.log(masked_exp.sum(dim, keepdim=True))


def sample_index(logits, mask, u):
    """Index function behind ``Policy.sample_action`` (``policy.py:23-33,190-195``).

    The reference draws with ``torch.multinomial(masked_probs[-1], 1)`` whose RNG stream cannot
    be reproduced on a GPU, so parity pins the *index function*: inverse-CDF over the masked
    probabilities for a given uniform ``u`` in [0,1).  fp32 sequential cumulative sum, first
    index whose cumulative mass exceeds ``u * total`` among valid entries.
    """
    lp = masked_softmax(logits.view(1, 1, -1), mask.view(1, 1, -1)).view(-1)
    probs = torch.exp(lp).clone()
    probs[~mask.view(-1)] = 0.0
    total = np.float32(0.0)
    for p in probs.numpy():
        total = np.float32(total + p)
    target = np.float32(np.float32(u) * total)
    acc = np.float32(0.0)
    last_valid = -1
    for i, p in enumerate(probs.numpy()):
        if not bool(mask.view(-1)[i]):
            continue
        last_valid = i
        acc = np.float32(acc + p)
        if acc > target:
            return i
    return last_valid
