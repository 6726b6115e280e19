"""GPU tests of the unit-encoder kernel chain and the target-unit head against plain torch on the CPU
(the same ops the oracle's ``RefPolicy.forward`` performs, ``policy.py:99-136,144-153``).

Tolerances: fp32 kernels + 3xTF32 tensor-core GEMMs vs fp32 CPU: forward atol 2e-5 on O(1) activations,
gradients rtol 2e-4 with an absolute floor scaled by the token count (sums over tokens)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

UNITS = (1, 5, 16, 16, 1, 1)


def _reference(env, w_e, b_e, w_b, b_b, units, weights, biases):
    emb = [F.linear(F.relu(F.linear(u, w_b, b_b)), w, b) for u, w, b in zip(units, weights, biases)]
    mx = [e.max(dim=-2)[0] for e in emb]
    mx[5] = mx[3]                                        # policy.py:127
    return torch.cat(emb, dim=-2), torch.cat([F.relu(F.linear(env, w_e, b_e))] + mx, dim=-1)   # policy.py:97,129-136


@pytest.mark.parametrize("lead", [(7,), (3, 5), (1,), (130,)])
def test_unit_encoder_forward_backward(lead):
    from dotaclient_b200 import encoder_ops
    g = torch.Generator().manual_seed(sum(lead))
    env = torch.randn(*lead, 3, generator=g)
    w_e = (torch.randn(128, 3, generator=g) * 0.5).requires_grad_(True)
    b_e = (torch.randn(128, generator=g) * 0.1).requires_grad_(True)
    w_b = (torch.randn(128, 12, generator=g) * 0.3).requires_grad_(True)
    b_b = (torch.randn(128, generator=g) * 0.1).requires_grad_(True)
    units = [torch.randn(*lead, n, 12, generator=g) for n in UNITS]
    weights = [(torch.randn(128, 128, generator=g) * 0.1).requires_grad_(True) for _ in UNITS]
    biases = [(torch.randn(128, generator=g) * 0.1).requires_grad_(True) for _ in UNITS]
    ue_r, xm_r = _reference(env, w_e, b_e, w_b, b_b, units, weights, biases)
    g_ue = torch.randn(ue_r.shape, generator=g)
    g_xm = torch.randn(xm_r.shape, generator=g)
    ((ue_r * g_ue).sum() + (xm_r * g_xm).sum()).backward()

    d = torch.device("cuda", 0)
    params = [t.detach().clone().to(d).requires_grad_(True) for t in [w_b, b_b] + weights + biases]
    pe = [t.detach().clone().to(d).requires_grad_(True) for t in (w_e, b_e)]
    ue, xm = encoder_ops.unit_encoder(env.to(d), pe[0], pe[1], params[0], params[1], [u.to(d) for u in units], params[2:8],
                                      params[8:14])
    assert ue.shape == ue_r.shape and xm.shape == xm_r.shape
    torch.testing.assert_close(ue.detach().cpu(), ue_r.detach(), rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(xm.detach().cpu(), xm_r.detach(), rtol=1e-4, atol=2e-5)
    ((ue * g_ue.to(d)).sum() + (xm * g_xm.to(d)).sum()).backward()
    n_tok = ue_r.numel() // (40 * 128)
    for mine, ref in zip(params + pe, [w_b, b_b] + weights + biases + [w_e, b_e]):
        torch.testing.assert_close(mine.grad.cpu(), ref.grad, rtol=2e-4, atol=2e-6 * max(1, n_tok) * 16)


def test_unit_max_tie_breaking_and_grad_routing():
    """Equal maxima: the FIRST unit wins (torch.max semantics on CPU) and receives the whole gradient."""
    from dotaclient_b200 import encoder_ops
    d = torch.device("cuda", 0)
    w_b = torch.zeros(128, 12)
    b_b = torch.ones(128)                                 # basic == 1 for every unit -> all units tie
    units = [torch.randn(4, n, 12) for n in UNITS]
    weights = [torch.eye(128) for _ in UNITS]
    biases = [torch.zeros(128) for _ in UNITS]
    params = [t.to(d).requires_grad_(True) for t in [w_b, b_b] + weights + biases]
    w_e, b_e = torch.zeros(128, 3, device=d, requires_grad=True), torch.full((128,), -1.0, device=d, requires_grad=True)
    ue, xcat = encoder_ops.unit_encoder(torch.randn(4, 3, device=d), w_e, b_e, params[0], params[1], [u.to(d) for u in units],
                                        params[2:8], params[8:14])
    xm = xcat[..., 128:]
    assert torch.equal(xm, torch.ones_like(xm))
    assert torch.equal(xcat[..., :128], torch.zeros_like(xcat[..., :128]))      # relu(-1) == 0: a dead env encoder ...
    ue.retain_grad()
    xcat.sum().backward()
    assert float(w_e.grad.abs().sum()) == 0.0 and float(b_e.grad.abs().sum()) == 0.0   # ... receives no gradient
    # bias gradient of group g == number of tokens routed to it: only unit 0 of each group gets d(max)
    for gidx in range(5):
        expect = 4.0 * (2.0 if gidx == 3 else 1.0)         # enh also receives the enemy-tower slot's gradient
        assert torch.allclose(params[8 + gidx].grad.cpu(), torch.full((128,), expect))
    assert float(params[13].grad.abs().sum()) == 0.0      # eth: no path from the maxima (policy.py:127)


@pytest.mark.parametrize("lead", [(9,), (4, 6), (257,)])
def test_target_unit_forward_backward(lead):
    from dotaclient_b200 import encoder_ops
    g = torch.Generator().manual_seed(len(lead) + lead[0])
    att = torch.randn(*lead, 128, generator=g).requires_grad_(True)
    ue = torch.randn(*lead, 40, 128, generator=g).requires_grad_(True)
    ref = torch.matmul(att.unsqueeze(-2), ue.transpose(-1, -2)).squeeze(-2)
    go = torch.randn(ref.shape, generator=g)
    go[..., ::3, :] = 0                                    # tokens where the head was not used: exact-zero rows
    (ref * go).sum().backward()
    d = torch.device("cuda", 0)
    a2, u2 = att.detach().to(d).requires_grad_(True), ue.detach().to(d).requires_grad_(True)
    out = encoder_ops.target_unit(a2, u2)
    torch.testing.assert_close(out.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-5)
    (out * go.to(d)).sum().backward()
    torch.testing.assert_close(a2.grad.cpu(), att.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(u2.grad.cpu(), ue.grad, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("lead,use_head", [((7,), True), ((3, 5), True), ((130,), True), ((33,), False)])
def test_unit_encoder_implicit_embedding_gradient(lead, use_head):
    """The training-path backward: the unit embedding is consumed only by the target-unit head (and the max-pool), so its
    [N,40,128] gradient is never materialised -- the group GEMMs generate their rows of it (dc_unit_group_dgrad/_wgrad).
    Compared with plain torch autograd on the CPU (policy.py:99-136,152-153)."""
    from dotaclient_b200 import encoder_ops
    g = torch.Generator().manual_seed(17 + sum(lead))
    env = torch.randn(*lead, 3, generator=g)
    w_e = (torch.randn(128, 3, generator=g) * 0.5).requires_grad_(True)
    b_e = (torch.randn(128, generator=g) * 0.1).requires_grad_(True)
    w_b = (torch.randn(128, 12, generator=g) * 0.3).requires_grad_(True)
    b_b = (torch.randn(128, generator=g) * 0.1).requires_grad_(True)
    units = [torch.randn(*lead, n, 12, generator=g) for n in UNITS]
    weights = [(torch.randn(128, 128, generator=g) * 0.1).requires_grad_(True) for _ in UNITS]
    biases = [(torch.randn(128, generator=g) * 0.1).requires_grad_(True) for _ in UNITS]
    att = torch.randn(*lead, 128, generator=g).requires_grad_(True)
    ue_r, x_r = _reference(env, w_e, b_e, w_b, b_b, units, weights, biases)
    g_x = torch.randn(x_r.shape, generator=g)
    loss_r = (x_r * g_x).sum()
    if use_head:
        tu_r = torch.matmul(att.unsqueeze(-2), ue_r.transpose(-1, -2)).squeeze(-2)
        g_tu = torch.randn(tu_r.shape, generator=g)
        g_tu[..., ::2, :] = 0                              # tokens where the head was not used
        loss_r = loss_r + (tu_r * g_tu).sum()
    loss_r.backward()

    d = torch.device("cuda", 0)
    refs = [w_b, b_b] + weights + biases + [w_e, b_e]
    params = [t.detach().clone().to(d).requires_grad_(True) for t in refs]
    att_d = att.detach().to(d).requires_grad_(True)
    ue, x = encoder_ops.unit_encoder(env.to(d), params[14], params[15], params[0], params[1], [u.to(d) for u in units],
                                     params[2:8], params[8:14])
    loss = (x * g_x.to(d)).sum()
    if use_head:
        tu = encoder_ops.target_unit(att_d, ue)
        torch.testing.assert_close(tu.detach().cpu(), tu_r.detach(), rtol=1e-4, atol=1e-4)
        loss = loss + (tu * g_tu.to(d)).sum()
    loss.backward()
    n_tok = ue_r.numel() // (40 * 128)
    for mine, ref in zip(params, refs):
        expect = ref.grad if ref.grad is not None else torch.zeros_like(ref)   # e.g. the enemy-tower layer without the head
        torch.testing.assert_close(mine.grad.cpu(), expect, rtol=2e-4, atol=2e-6 * max(1, n_tok) * 16)
    if use_head:
        torch.testing.assert_close(att_d.grad.cpu(), att.grad, rtol=1e-4, atol=1e-4)
