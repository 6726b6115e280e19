"""Micro-benchmark of dc_gemm_tf32x3 against cuBLAS fp32 / TF32 on the model's GEMM shapes (CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dotaclient_b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    d = torch.device("cuda", 0)
    shapes = [("unit-embedding (16 units)", 131072 * 16, 128, 128), ("i2h lstm H128", 131072, 512, 128),
              ("pre_rnn", 131072, 128, 896), ("i2h lstm H512 (c4)", 524288 // 4, 2048, 512)]
    for name, M, N, K in shapes:
        a = torch.randn(M, K, device=d)
        b = torch.randn(N, K, device=d) * 0.1
        bias = torch.randn(N, device=d)
        out = torch.empty(M, N, device=d)
        ref = torch.addmm(bias, a, b.t())
        got = ops.gemm_tf32x3(a, b, bias, out=out)
        err = (got - ref).abs().max().item()
        t_ours = timeit(lambda: ops.gemm_tf32x3(a, b, bias, out=out))
        torch.backends.cuda.matmul.allow_tf32 = False
        t_fp32 = timeit(lambda: torch.addmm(bias, a, b.t(), out=out))
        torch.backends.cuda.matmul.allow_tf32 = True
        t_tf32 = timeit(lambda: torch.addmm(bias, a, b.t(), out=out))
        torch.backends.cuda.matmul.allow_tf32 = False
        flops = 2.0 * M * N * K
        bytes_ = 4.0 * (M * K + N * K + M * N)
        print("%-28s M=%8d N=%4d K=%4d | ours %.3f ms (%.1f TF/s eff, %.0f GB/s) | cublas fp32 %.3f ms | cublas tf32 %.3f ms | max|diff vs fp32| %.2e"
              % (name, M, N, K, t_ours, flops / t_ours / 1e9, bytes_ / t_ours / 1e6, t_fp32, t_tf32, err))

    # the unit-embedding GEMMs as the encoder issues them: one unit group (16 of 40 rows per token) of [N, 40, 128] in place
    from dotaclient_b200 import _lib
    lib = _lib.load()
    st = _lib.stream_ptr()
    N_tok, n_u, C = 131072, 16, 128
    R = N_tok * n_u
    ue = torch.empty(N_tok, 40, C, device=d)
    basic = torch.randn(R, C, device=d)
    w = torch.randn(C, C, device=d) * 0.1
    bias = torch.randn(C, device=d)
    off = 6 * C * 4
    t_c = timeit(lambda: _lib.check(lib.dc_gemm_tf32x3_blocked(basic.data_ptr(), C, 0, 0, w.data_ptr(), C, bias.data_ptr(),
                                                              ue.data_ptr() + off, C, n_u, 40 * C, R, C, C, 0, st), "gemm"))
    t_a = timeit(lambda: _lib.check(lib.dc_gemm_tf32x3_blocked(ue.data_ptr() + off, C, n_u, 40 * C, w.data_ptr(), C, None,
                                                              basic.data_ptr(), C, 0, 0, R, C, C, 0, st), "gemm"))
    ref = torch.addmm(bias, basic, w.t()).view(N_tok, n_u, C)
    _lib.check(lib.dc_gemm_tf32x3_blocked(basic.data_ptr(), C, 0, 0, w.data_ptr(), C, bias.data_ptr(), ue.data_ptr() + off, C, n_u,
                                          40 * C, R, C, C, 0, st), "gemm")
    err = (ue[:, 6:22] - ref).abs().max().item()
    print("unit group 16/40: C blocked (fwd) %.3f ms | A blocked (dgrad) %.3f ms | %.0f / %.0f GB/s | max|diff| %.2e"
          % (t_c, t_a, 8.0 * R * C / t_c / 1e6, 8.0 * R * C / t_a / 1e6, err))

    # weight gradient of the same group: dW = d_ue_g^T basic (dY two-level rows, X plain), db = column sums
    ws = torch.empty(int(lib.dc_gemm_wgrad_workspace_bytes(C, C)), dtype=torch.uint8, device=d)
    dw, db = torch.empty(C, C, device=d), torch.empty(C, device=d)
    t_w = timeit(lambda: _lib.check(lib.dc_gemm_wgrad_tf32x3_blocked(ue.data_ptr() + off, C, n_u, 40 * C, basic.data_ptr(), C, R, C, C,
                                                                    dw.data_ptr(), C, db.data_ptr(), 0, ws.data_ptr(), st), "wgrad"))
    ref_w = (ue[:, 6:22].reshape(R, C).double().t() @ basic.double()).float()
    print("unit group 16/40 weight gradient: %.3f ms | %.0f GB/s | max|diff|/max|ref| %.2e"
          % (t_w, 8.0 * R * C / t_w / 1e6, ((dw - ref_w).abs().max() / ref_w.abs().max()).item()))


if __name__ == "__main__":
    main()
