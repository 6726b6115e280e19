"""Stand-alone check of the fused unit-encoder backward kernels (dc_unit_wgrad_routed, dc_unit_dgrad_fused) against dense
torch on the CPU, one line per group -- the debugging companion of tests/test_gpu_encoder.py (run it under `timeout`).

    python tools/encoder_check.py [N ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from dotaclient_b200 import _lib  # noqa: E402

UNITS = (1, 5, 16, 16, 1, 1)
OFF = (0, 1, 6, 22, 38, 39)
C = 128


def dense_reference(N, n_u, seed, with_dx2, with_head, routed=True):
    """-> inputs and (dW_route, db_route, dW_b, db_b) computed densely in fp64 on the CPU."""
    g = torch.Generator().manual_seed(seed)
    # inputs of the basic layer on a coarse binary grid: its pre-activation is then exact in fp32 in ANY summation order, so the
    # ReLU mask of this reference and the one the kernel recomputes cannot differ by rounding
    units = torch.round(torch.randn(N * n_u, 12, generator=g) * 16) / 16
    w_b, b_b = torch.round(torch.randn(C, 12, generator=g) * 0.3 * 64) / 64, torch.round(torch.randn(C, generator=g) * 0.1 * 64) / 64
    W = torch.randn(C, C, generator=g) * 0.1
    dx = torch.randn(N, 7 * C, generator=g)                         # a [N, 896] gradient row; slot 2 (and 5) used below
    am = torch.randint(0, n_u, (N, C), generator=g).to(torch.uint8)
    dl = torch.randn(N, 40, generator=g)
    dl[::2] = 0
    att = torch.randn(N, C, generator=g)
    basic = F.relu(F.linear(units, w_b, b_b))
    d = dx[:, 2 * C:3 * C] + (dx[:, 5 * C:6 * C] if with_dx2 else 0)
    route = torch.zeros(N, n_u, C, dtype=torch.float64)
    if routed:
        route.scatter_(1, am.long().unsqueeze(1), d.double().unsqueeze(1))
    dW = route.reshape(N * n_u, C).t() @ basic.double()
    db = route.sum((0, 1))
    d_emb = route
    if with_head:
        d_emb = d_emb + dl[:, 3:3 + n_u].double().unsqueeze(-1) * att.double().unsqueeze(1)
    d_basic = d_emb.reshape(N * n_u, C) @ W.double()
    gm = d_basic * (basic > 0)
    return dict(units=units, w_b=w_b, b_b=b_b, W=W, dx=dx, am=am, dl=dl, att=att, basic=basic), (dW, db, gm.t() @ units.double(), gm.sum(0))


def run(N, n_u, with_dx2, with_head, routed=True, seed=0):
    lib = _lib.load()
    st = _lib.stream_ptr()
    dev = torch.device("cuda", 0)
    inp, (dW_r, db_r, dwb_r, dbb_r) = dense_reference(N, n_u, seed, with_dx2, with_head, routed)
    t = {k: v.to(dev).contiguous() for k, v in inp.items()}
    wt = t["W"].t().contiguous()
    f = lambda x, off=0: x.data_ptr() + 4 * off  # noqa: E731
    msg = "N=%4d units=%2d dx2=%d head=%d routed=%d |" % (N, n_u, with_dx2, with_head, routed)
    if n_u > 1 and routed:
        dW, db = torch.full((C, C), 7.0, device=dev), torch.full((C,), 7.0, device=dev)
        ws = torch.empty(int(lib.dc_gemm_wgrad_workspace_bytes(C, C)), dtype=torch.uint8, device=dev)
        _lib.check(lib.dc_unit_wgrad_routed(f(t["dx"], 2 * C), f(t["dx"], 5 * C) if with_dx2 else None, 7 * C, t["am"].data_ptr(),
                                            t["basic"].data_ptr(), N, n_u, dW.data_ptr(), db.data_ptr(), ws.data_ptr(), st), "wgrad_routed")
        torch.cuda.synchronize()
        msg += " dW %.2e (ref max %.2e) db %.2e |" % ((dW.cpu().double() - dW_r).abs().max(), dW_r.abs().max(), (db.cpu().double() - db_r).abs().max())
    dwb, dbb = torch.full((C, 12), 7.0, device=dev), torch.full((C,), 7.0, device=dev)
    ws = torch.empty(int(lib.dc_unit_basic_bwd_workspace_bytes()), dtype=torch.uint8, device=dev)
    _lib.check(lib.dc_unit_dgrad_fused(f(t["dx"], 2 * C) if routed else None, f(t["dx"], 5 * C) if with_dx2 else None, 7 * C,
                                       t["am"].data_ptr() if (routed and n_u > 1) else None,
                                       f(t["dl"], 3) if with_head else None, 40, t["att"].data_ptr() if with_head else None,
                                       wt.data_ptr(), t["units"].data_ptr(), t["w_b"].data_ptr(), t["b_b"].data_ptr(), N, n_u,
                                       dwb.data_ptr(), dbb.data_ptr(), 0, ws.data_ptr(), st), "dgrad_fused")
    torch.cuda.synchronize()
    msg += " dW_b %.2e (ref max %.2e) db_b %.2e (ref max %.2e)" % ((dwb.cpu().double() - dwb_r).abs().max(), dwb_r.abs().max(),
                                                                   (dbb.cpu().double() - dbb_r).abs().max(), dbb_r.abs().max())
    print(msg, flush=True)


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [1, 7, 130, 1000, 40000]
    for N in sizes:
        for n_u in (1, 5, 16):
            run(N, n_u, with_dx2=(n_u == 16), with_head=True)
        run(N, 16, with_dx2=False, with_head=False)
        run(N, 1, with_dx2=False, with_head=True, routed=False)
