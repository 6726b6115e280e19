"""Deterministic synthetic rollouts in the reference's experience wire layout.

Layout = what ``agent.py:350-416`` packs and ``optimizer.py:314-336`` consumes: a dict with
``observations`` (7 fp32 tensors ``[L, units, feat]``), ``masks`` / ``actions`` (5 tensors
``[L, n]``; one-hot or all-zero rows), ``rewards`` (``np.float32 [L, 10]``), plus ids.

Masks and actions are ``torch.bool``: the agent's uint8 masks no longer index under
torch >= 2 (``policy.py:174``), see DESIGN.md "drift".  Distributions follow SURVEY.md 8(d).
"""
import numpy as np
import torch

OBS_SHAPES = {
    "env": (3,),
    "allied_heroes": (1, 12),
    "enemy_heroes": (5, 12),
    "allied_nonheroes": (16, 12),
    "enemy_nonheroes": (16, 12),
    "allied_towers": (1, 12),
    "enemy_towers": (1, 12),
}
HEAD_SIZES = {"enum": 4, "x": 9, "y": 9, "target_unit": 40, "ability": 3}


def make_rollout(length, seed, game_id=0, team_id=2, weight_version=1, with_canvas=False):
    """One rollout of ``length`` steps.  ``seed`` fully determines it (CPU generator)."""
    g = torch.Generator().manual_seed(int(seed))
    L = int(length)
    obs = {k: torch.randn((L,) + shp, generator=g, dtype=torch.float32) for k, shp in OBS_SHAPES.items()}
    rewards = (torch.randn((L, 10), generator=g, dtype=torch.float32) * 0.01).numpy()
    enum = torch.randint(0, 4, (L,), generator=g)
    masks = {k: torch.zeros((L, n), dtype=torch.bool) for k, n in HEAD_SIZES.items()}
    actions = {k: torch.zeros((L, n), dtype=torch.bool) for k, n in HEAD_SIZES.items()}
    rows = torch.arange(L)
    masks["enum"][:] = True
    actions["enum"][rows, enum] = True
    move, attack, ability = enum == 1, enum == 2, enum == 3
    for k in ("x", "y"):
        pick = torch.randint(0, 9, (L,), generator=g)
        masks[k][move] = True
        actions[k][rows[move], pick[move]] = True
    valid = torch.rand((L, 40), generator=g) < 0.5
    valid[:, 0] = False                                   # the own hero is never a target (policy.py:255)
    forced = torch.randint(1, 40, (L,), generator=g)
    valid[rows, forced] = True                            # at least one valid unit
    score = torch.rand((L, 40), generator=g).masked_fill(~valid, -1.0)
    target = score.argmax(dim=1)                          # uniform over the valid units
    masks["target_unit"][attack] = valid[attack]
    actions["target_unit"][rows[attack], target[attack]] = True
    pick = torch.randint(0, 3, (L,), generator=g)
    masks["ability"][ability] = True
    actions["ability"][rows[ability], pick[ability]] = True
    data = {
        "game_id": game_id, "team_id": team_id, "player_id": 0, "weight_version": weight_version,
        "observations": obs, "masks": masks, "actions": actions, "rewards": rewards,
    }
    if with_canvas:
        data["canvas"] = np.zeros((256, 256, 3), dtype=np.uint8)   # agent.py:424-437
    return data


def rollout_seed(rank, index):
    """``7 + 1000*rank + i``."""
    return 7 + 1000 * int(rank) + int(index)


def ragged_lengths(n, seq_len, seed):
    """Correctness configs: L ~ U[S/2, 3S] so multi-chunk carry and tail padding are exercised."""
    rng = np.random.RandomState(seed)
    return [int(v) for v in rng.randint(max(1, seq_len // 2), 3 * seq_len + 1, size=n)]
