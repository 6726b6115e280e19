"""Micro-benchmark of the recurrence kernels (dc_rnn_seq_fwd / dc_rnn_seq_bwd) against cuDNN's nn.LSTM / nn.GRU, per
(cell, batch, seq, hidden) shape -- CUDA events, kernel time per launch from ops.PROFILE, results checked against the cuDNN
layer (same weights).  The recurrence is the step's dependency-bound piece (2*S sequential steps); this is the harness for
working on its per-step latency:

    python tools/rnn_bench.py                      # the BASELINE shapes: C2 (256 x 512 x 128 LSTM), GRU, H=256/512
    python tools/rnn_bench.py lstm 256 512 128     # one shape
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dotaclient_b200 import ops  # noqa: E402


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def bench(cell, B, S, H):
    d = torch.device("cuda", 0)
    torch.manual_seed(0)
    layer = (torch.nn.LSTM if cell == "lstm" else torch.nn.GRU)(H, H).to(d)
    x = torch.randn(S, B, H, device=d, requires_grad=True)
    h0 = torch.zeros(B, H, device=d)
    c0 = torch.zeros(B, H, device=d) if cell == "lstm" else None
    G = 4 if cell == "lstm" else 3
    params = [layer.weight_ih_l0, layer.weight_hh_l0, layer.bias_ih_l0, layer.bias_hh_l0]

    def ours():
        y, hn, cn = ops.rnn_sequence(x, *params, h0, c0, cell)
        y.sum().backward()
        return y

    def cudnn():
        hx = (h0.unsqueeze(0), c0.unsqueeze(0)) if cell == "lstm" else h0.unsqueeze(0)
        y, _ = layer(x, hx)
        y.sum().backward()
        return y

    y_ours, y_ref = ours().detach(), cudnn().detach()
    err = (y_ours - y_ref).abs().max().item()
    ops.PROFILE.reset(enabled=True)
    n = 5
    t_ours = timeit(ours, n)
    prof = ops.PROFILE.summary(n + 2)
    ops.PROFILE.reset(enabled=False)
    t_ref = timeit(cudnn, n)                       # cuDNN default: TF32 tensor cores allowed (torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    t_ref32 = timeit(cudnn, n)                     # cuDNN restricted to fp32 arithmetic (the like-for-like comparison)
    torch.backends.cudnn.allow_tf32 = True
    fwd, bwd = prof.get("rnn_fwd", 0.0), prof.get("rnn_bwd", 0.0)
    nbytes = 12.0 * S * B * (G + 1) * H
    print("%-4s B=%4d S=%4d H=%3d | fwd %.3f ms (%.2f us/step) bwd %.3f ms (%.2f us/step) | %.0f GB/s algorithmic | "
          "layer fwd+bwd incl. GEMMs: ours %.3f ms, cuDNN(tf32) %.3f ms, cuDNN(fp32) %.3f ms | max|y - cudnn| %.2e"
          % (cell, B, S, H, fwd, 1e3 * fwd / S, bwd, 1e3 * bwd / S, nbytes / ((fwd + bwd) * 1e-3) / 1e9 if fwd + bwd > 0 else 0.0,
             t_ours, t_ref, t_ref32, err))


def main():
    if len(sys.argv) == 5:
        shapes = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))]
    else:
        shapes = [("lstm", 256, 512, 128), ("gru", 256, 512, 128), ("lstm", 128, 512, 128), ("lstm", 512, 512, 128),
                  ("gru", 512, 512, 256), ("lstm", 512, 512, 256), ("gru", 1024, 16, 256), ("lstm", 512, 1024, 512)]
    for s in shapes:
        bench(*s)


if __name__ == "__main__":
    main()
