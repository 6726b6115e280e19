"""Quick stand-alone check of the recurrence kernels against torch.nn.GRU / nn.LSTM on the CPU (fp32), one shape per
line -- the debugging companion of tests/test_gpu_parity.py::test_rnn_forward_backward_vs_torch (run it under `timeout`).

    python tools/rnn_check.py [fwd|all] cell B S H [cell B S H ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dotaclient_b200 import ops  # noqa: E402


def check(cell, B, S, H, backward):
    torch.manual_seed(B * 1000 + S * 10 + H)
    ref = (torch.nn.GRU if cell == "gru" else torch.nn.LSTM)(H, H)
    x = torch.randn(S, B, H)
    h0, c0 = torch.randn(1, B, H) * 0.5, torch.randn(1, B, H) * 0.5
    wy = torch.randn(S, B, H)
    xr = x.clone().requires_grad_(True)
    h0r, c0r = h0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
    yr, st = ref(xr, (h0r, c0r) if cell == "lstm" else h0r)
    if backward:
        (yr * wy).sum().backward()
    d = torch.device("cuda", 0)
    p = {k: v.detach().clone().to(d).requires_grad_(True) for k, v in ref.named_parameters()}
    xg = x.to(d).requires_grad_(True)
    h0g = h0[0].to(d).requires_grad_(True)
    c0g = c0[0].to(d).requires_grad_(True) if cell == "lstm" else None
    y, hn, cn = ops.rnn_sequence(xg, p["weight_ih_l0"], p["weight_hh_l0"], p["bias_ih_l0"], p["bias_hh_l0"], h0g, c0g, cell)
    torch.cuda.synchronize()
    msg = "%-4s B=%3d S=%3d H=%3d | max|y - ref| %.2e" % (cell, B, S, H, (y.detach().cpu() - yr.detach()).abs().max().item())
    if backward:
        (y * wy.to(d)).sum().backward()
        torch.cuda.synchronize()
        msg += " | dx %.2e dh0 %.2e" % ((xg.grad.cpu() - xr.grad).abs().max().item(), (h0g.grad.cpu() - h0r.grad[0]).abs().max().item())
        for k, v in ref.named_parameters():
            msg += " %s %.2e" % (k.replace("weight_", "dw_").replace("bias_", "db_").replace("_l0", ""),
                                 ((p[k].grad.cpu() - v.grad).abs().max() / v.grad.abs().max().clamp_min(1e-12)).item())
    print(msg, flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    backward = a[0] != "fwd"
    a = a[1:]
    for i in range(0, len(a), 4):
        check(a[i], int(a[i + 1]), int(a[i + 2]), int(a[i + 3]), backward)
