"""Torch-facing wrappers of the C-ABI kernels (device memory + stream plumbing only).

Every function here requires CUDA tensors and enqueues hand-written sm_100a kernels from
``libdotaclient_b200.so`` on torch's current stream.  No CPU path exists: CPU tensors raise.
"""
import ctypes

import torch

from . import _lib

CELL_ID = {"gru": 0, "lstm": 1}
HEAD_KEYS = ("enum", "x", "y", "target_unit", "ability")     # policy.py:46
HEAD_SIZES = (4, 9, 9, 40, 3)
GATES = {"gru": 3, "lstm": 4}


class _Profile:
    """Optional per-kernel CUDA-event timing on the launching stream + a count of OUR kernel launches.

    Used by bench.py (roofline) -- events are recorded around each C-ABI call on torch's current stream.
    """

    def __init__(self):
        self.reset(False)

    def reset(self, enabled=False):
        self.enabled = enabled
        self.pairs = []
        self.launches = 0
        self.bytes = {}            # name -> algorithmic HBM bytes of the timed calls (inputs read once + outputs written once)

    class _Span:
        def __init__(self, prof, name, n_launches, nbytes=0):
            self.prof, self.name, self.n, self.nbytes = prof, name, n_launches, nbytes

        def __enter__(self):
            if self.prof.enabled:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
            return self

        def __exit__(self, *exc):
            if self.prof.enabled:
                self.e1.record()
                self.prof.pairs.append((self.name, self.e0, self.e1))
                self.prof.launches += self.n
                self.prof.bytes[self.name] = self.prof.bytes.get(self.name, 0) + self.nbytes
            return False

    def span(self, name, n_launches, nbytes=0):
        return _Profile._Span(self, name, n_launches, nbytes)

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.pairs:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out

    def summary(self, steps=None):
        """ms per kernel name, averaged per step when ``steps`` is given (else per call)."""
        tot = self.totals()
        if steps:
            return {k: v / steps for k, v in tot.items()}
        counts = {}
        for name, _, _ in self.pairs:
            counts[name] = counts.get(name, 0) + 1
        return {k: v / counts[k] for k, v in tot.items()}


PROFILE = _Profile()


# Events of in-flight host->device copies issued on a side stream (ExperienceBatch.to), keyed by the destination's
# data pointer; the consumer makes the compute stream wait right before the tensor's first use.
H2D_EVENTS = {}


def wait_h2d(*tensors):
    if not H2D_EVENTS:
        return
    for t in tensors:
        if t is None:
            continue
        ev = H2D_EVENTS.pop(t.data_ptr(), None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dotaclient_b200 kernels need CUDA tensors (no CPU fallback); got %s" % t.device)


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.to(torch.float32).contiguous()
    return t


def _u8(t):
    """bool/uint8 tensor -> contiguous byte view (0/1)."""
    t = t.contiguous()
    if t.dtype == torch.bool:
        return t.view(torch.uint8)
    if t.dtype != torch.uint8:
        t = (t != 0).view(torch.uint8)
    return t


# --------------------------------------------------------------------------------------------- GAE
def gae_scan(rewards, values, seg_off, gamma=0.98, lam=0.97, boot_value=None, boot_reward=None):
    """GAE advantages + rewards-to-go for many rollouts (``optimizer.py:53-64,397,417-421``).

    rewards [n_rows] or [n_rows, n_sub] fp32; values [n_rows]; seg_off int64 [n_seg+1] (device).
    """
    _need_cuda(rewards, values, seg_off)
    rewards, values = _f32c(rewards), _f32c(values)
    n_sub = 1 if rewards.dim() == 1 else rewards.shape[1]
    n_rows = values.numel()
    assert rewards.numel() == n_rows * n_sub
    seg_off = seg_off.to(torch.int64).contiguous()
    adv = torch.empty(n_rows, dtype=torch.float32, device=values.device)
    ret = torch.empty_like(adv)
    lib = _lib.load()
    with PROFILE.span("gae_scan", 1):
        _lib.check(lib.dc_gae_scan(rewards.data_ptr(), n_sub, values.data_ptr(), seg_off.data_ptr(),
                                   seg_off.numel() - 1, _lib.ptr(boot_value), _lib.ptr(boot_reward), float(gamma),
                                   float(lam), adv.data_ptr(), ret.data_ptr(), _lib.stream_ptr()), "dc_gae_scan")
    return adv, ret


# --------------------------------------------------------------------------------------------- RNN
_workspaces = {}


def _rnn_workspace(cell, B, H, device):
    key = (cell, H, device)
    nbytes = max(int(_lib.load().dc_rnn_workspace_bytes(CELL_ID[cell], B, H)), 16)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _rnn_forward_impl(x, w_ih, w_hh, b_ih, b_hh, h0, c0, cell):
    """i2h GEMM (tcgen05 3xTF32, ``dc_gemm_tf32x3``) + recurrence kernel.  Returns (x2, w_ih, w_hh, gates, ybuf, cbuf)."""
    _need_cuda(x, w_ih, w_hh, b_ih, b_hh, h0, c0)
    S, B, Hin = x.shape
    H = w_hh.shape[1]
    N = S * B
    x2 = _f32c(x.detach()).view(N, Hin)
    w_ih, w_hh, b_ih, b_hh = _f32c(w_ih.detach()), _f32c(w_hh.detach()), _f32c(b_ih.detach()), _f32c(b_hh.detach())
    # [N, G*H] = x W_ih^T + b_ih on tcgen05 (3xTF32); shapes outside the kernel's (G*H % 128, Hin % 32) raise DC_EUNSUPPORTED
    gates = gemm_tf32x3(x2, w_ih, b_ih)
    ybuf = torch.empty((S + 1, B, H), dtype=torch.float32, device=x.device)
    cbuf = torch.empty((S + 1, B, H), dtype=torch.float32, device=x.device)
    ybuf[0].copy_(h0.detach().reshape(B, H))
    if cell == "lstm":
        cbuf[0].copy_(c0.detach().reshape(B, H))
    ws = _rnn_workspace(cell, B, H, x.device)
    lib = _lib.load()
    with PROFILE.span("rnn_fwd", 1, 4 * S * B * ((4 if cell == "lstm" else 3) + 1) * H):
        _lib.check(lib.dc_rnn_seq_fwd(CELL_ID[cell], gates.data_ptr(), w_hh.data_ptr(), b_hh.data_ptr(),
                                      ybuf.data_ptr(), cbuf.data_ptr(), B, S, H, ws.data_ptr(), _lib.stream_ptr()),
                   "dc_rnn_seq_fwd")
    return x2, w_ih, w_hh, gates, ybuf, cbuf


def rnn_forward_states(x_tm, w_ih, w_hh, b_ih, b_hh, h0, c0, cell):
    """No-grad forward returning the full state buffers: ybuf [S+1,B,H] (slot 0 = h0, slot t+1 = h_t) and
    cbuf [S+1,B,H] (LSTM cell states).  Used by experience prep to read the hidden state at chunk boundaries."""
    with torch.no_grad():
        _, _, _, _, ybuf, cbuf = _rnn_forward_impl(x_tm, w_ih, w_hh, b_ih, b_hh, h0, c0, cell)
    return ybuf, cbuf


class RnnSequence(torch.autograd.Function):
    """Time-major GRU/LSTM layer: i2h GEMM (tcgen05 3xTF32) + hand-written recurrence kernels.

    forward(x [S,B,Hin], w_ih [G*H,Hin], w_hh [G*H,H], b_ih, b_hh, h0 [B,H], c0 [B,H]|None, cell)
      -> y [S,B,H], h_n [B,H], c_n [B,H] (zeros-size-0 tensor for GRU)
    Semantics of ``torch.nn.GRU/LSTM(batch_first=...)`` as used in ``policy.py:66,141``.
    """

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, h0, c0, cell):
        S, B, Hin = x.shape
        H = w_hh.shape[1]
        x2, w_ih, w_hh, gates, ybuf, cbuf = _rnn_forward_impl(x, w_ih, w_hh, b_ih, b_hh, h0, c0, cell)
        ctx.cell, ctx.dims = cell, (S, B, Hin, H, GATES[cell])
        ctx.save_for_backward(x2, w_ih, w_hh, gates, ybuf, cbuf)
        y = ybuf[1:]
        h_n = ybuf[S].clone()
        if cell == "lstm":
            c_n = cbuf[S].clone()
        else:
            c_n = ybuf.new_empty(0)
            ctx.mark_non_differentiable(c_n)
        return y, h_n, c_n

    @staticmethod
    def backward(ctx, dy, dhn, dcn):
        x2, w_ih, w_hh, gates, ybuf, cbuf = ctx.saved_tensors
        cell = ctx.cell
        S, B, Hin, H, G = ctx.dims
        N = S * B
        if getattr(ctx, "_consumed", False):
            raise RuntimeError("RnnSequence backward ran twice: the saved gate buffer is consumed in place")
        ctx._consumed = True
        dy = _f32c(dy) if dy is not None else torch.zeros((S, B, H), dtype=torch.float32, device=x2.device)
        dhn = _f32c(dhn) if dhn is not None else None
        dcn = _f32c(dcn) if (dcn is not None and cell == "lstm" and dcn.numel()) else None
        dh0 = torch.empty((B, H), dtype=torch.float32, device=x2.device)
        dc0 = torch.empty((B, H), dtype=torch.float32, device=x2.device) if cell == "lstm" else None
        ws = _rnn_workspace(cell, B, H, x2.device)
        lib = _lib.load()
        with PROFILE.span("rnn_bwd", 1, 8 * S * B * ((4 if cell == "lstm" else 3) + 1) * H):
            _lib.check(lib.dc_rnn_seq_bwd(CELL_ID[cell], gates.data_ptr(), w_hh.data_ptr(), ybuf.data_ptr(),
                                          cbuf.data_ptr(), dy.data_ptr(), _lib.ptr(dhn), _lib.ptr(dcn), dh0.data_ptr(),
                                          _lib.ptr(dc0), B, S, H, ws.data_ptr(), _lib.stream_ptr()), "dc_rnn_seq_bwd")
        dgi = gates                                   # [N, G*H], overwritten in place by the kernel
        hprev = ybuf[:S].view(N, H)                   # h_{t-1} for every token (slot t)
        dx = gemm_tf32x3(dgi, w_ih.t().contiguous()).view(S, B, Hin) if ctx.needs_input_grad[0] else None   # dx = dgi W_ih
        dw_ih, db_ih = gemm_wgrad_tf32x3(dgi, x2)                              # dW_ih = dgi^T x, db_ih = colsum(dgi)
        if cell == "lstm":
            dw_hh = gemm_wgrad_tf32x3(dgi, hprev, want_bias=False)[0]
            db_hh = db_ih
        else:
            dghn = cbuf[1:].view(N, H)                # n-gate part of dgh (= dgi_n * r)
            dw_hh = torch.empty_like(w_hh)
            gemm_wgrad_tf32x3(dgi[:, :2 * H], hprev, want_bias=False, dw_out=dw_hh[:2 * H])
            _, db_n = gemm_wgrad_tf32x3(dghn, hprev, dw_out=dw_hh[2 * H:])
            db_hh = torch.cat([db_ih[:2 * H], db_n])
        return dx, dw_ih, dw_hh, db_ih, db_hh, dh0, dc0, None


def rnn_sequence(x_tm, w_ih, w_hh, b_ih, b_hh, h0, c0, cell):
    return RnnSequence.apply(x_tm, w_ih, w_hh, b_ih, b_hh, h0, c0, cell)


# --------------------------------------------------------------------------------------------- PPO loss
def ppo_loss_fwd_bwd(logits, masks, actions, old_logp, adv_raw, ret, value, e_clip, entropy_coef, vf_coef):
    """Fused PPO loss + gradients (``optimizer.py:587-589,621-665`` and their backward).

    logits/masks/actions: sequences of 5 tensors [..., n_h] in HEAD_KEYS order (any leading dims,
    same token order everywhere); old_logp [..., 5]; adv_raw/ret/value [...].
    Returns (out[16] fp32, n_actions[5] int32, dlogits list, dvalue) -- all on device, no sync.
    """
    logits = [_f32c(l.detach()) for l in logits]
    _need_cuda(*logits)
    N = logits[0].numel() // HEAD_SIZES[0]
    masks = [_u8(m) for m in masks]
    actions = [_u8(a) for a in actions]
    for h in range(5):
        assert logits[h].numel() == N * HEAD_SIZES[h] and masks[h].numel() == N * HEAD_SIZES[h] \
            and actions[h].numel() == N * HEAD_SIZES[h], "head %d shape mismatch" % h
    old_logp, adv_raw, ret, value = _f32c(old_logp), _f32c(adv_raw), _f32c(ret), _f32c(value.detach())
    assert old_logp.numel() == N * 5 and adv_raw.numel() == N and ret.numel() == N and value.numel() == N
    dev = logits[0].device
    dlogits = [torch.empty_like(l) for l in logits]
    dvalue = torch.empty_like(value)
    out = torch.empty(_lib.LOSS_SLOTS, dtype=torch.float32, device=dev)
    n_actions = torch.empty(5, dtype=torch.int32, device=dev)
    ws = torch.empty(_lib.PPO_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)
    lib = _lib.load()
    with PROFILE.span("ppo_loss", 2):
        _lib.check(lib.dc_ppo_loss_fwd_bwd(_lib.ptr5(logits), _lib.ptr5(masks), _lib.ptr5(actions),
                                           old_logp.data_ptr(), adv_raw.data_ptr(), ret.data_ptr(), value.data_ptr(),
                                           N, float(e_clip), float(entropy_coef), float(vf_coef), _lib.ptr5(dlogits),
                                           dvalue.data_ptr(), out.data_ptr(), n_actions.data_ptr(), ws.data_ptr(),
                                           _lib.stream_ptr()), "dc_ppo_loss_fwd_bwd")
    return out, n_actions, dlogits, dvalue


def selected_logp(logits, masks, actions):
    """Dense [N,5] log-prob of the taken action per head (0 where none) -- ``optimizer.py:387-390``."""
    logits = [_f32c(l.detach()) for l in logits]
    _need_cuda(*logits)
    N = logits[0].numel() // HEAD_SIZES[0]
    masks = [_u8(m) for m in masks]
    actions = [_u8(a) for a in actions]
    out = torch.empty((N, 5), dtype=torch.float32, device=logits[0].device)
    lib = _lib.load()
    with PROFILE.span("selected_logp", 1):
        _lib.check(lib.dc_selected_logp(_lib.ptr5(logits), _lib.ptr5(masks), _lib.ptr5(actions), N, out.data_ptr(),
                                        _lib.stream_ptr()), "dc_selected_logp")
    return out


# --------------------------------------------------------------------------------------------- grad finish
def grad_flags(flat_grad, total, seg_head, n_actions):
    lib = _lib.load()
    with PROFILE.span("grad_flags", 1):
        _lib.check(lib.dc_grad_flags(flat_grad.data_ptr(), total, seg_head.data_ptr(), seg_head.numel(),
                                     n_actions.data_ptr(), _lib.stream_ptr()), "dc_grad_flags")


def grad_finish(flat_param, flat_grad, exp_avg, exp_avg_sq, steps, seg_lo, seg_hi, seg_head, total, lr, betas, eps,
                max_norm, loss_out, metrics, workspace):
    lib = _lib.load()
    with PROFILE.span("grad_finish", 3):
        _lib.check(lib.dc_grad_finish(flat_param.data_ptr(), flat_grad.data_ptr(), exp_avg.data_ptr(),
                                      exp_avg_sq.data_ptr(), steps.data_ptr(), seg_lo.data_ptr(), seg_hi.data_ptr(),
                                      seg_head.data_ptr(),
                                      seg_head.numel(), total, float(lr), float(betas[0]), float(betas[1]), float(eps),
                                      float(max_norm), _lib.ptr(loss_out), metrics.data_ptr(), workspace.data_ptr(),
                                      _lib.stream_ptr()), "dc_grad_finish")


# --------------------------------------------------------------------------------------------- tensor-core GEMM
def gemm_tf32x3_supported(M, N, K):
    return bool(_lib.load().dc_gemm_tf32x3_supported(int(M), int(N), int(K)))


def gemm_tf32x3(a, b, bias=None, relu=False, out=None):
    """``out[M,N] = a[M,K] @ b[N,K]^T (+ bias) (ReLU)`` on tcgen05 tensor cores with the 3xTF32 split
    (fp32-level accuracy).  ``a``/``b``/``out`` are 2-D fp32 CUDA tensors whose rows are contiguous (row stride
    may exceed the width: column-slice views are fine)."""
    _need_cuda(a, b, bias)
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1]
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    lib = _lib.load()
    with PROFILE.span("gemm_tf32x3", 1, 4 * (M * K + N * K + M * N)):
        _lib.check(lib.dc_gemm_tf32x3(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), _lib.ptr(bias), out.data_ptr(),
                                      out.stride(0), M, N, K, 1 if relu else 0, _lib.stream_ptr()), "dc_gemm_tf32x3")
    return out


class LinearTC(torch.autograd.Function):
    """``y = x W^T + b`` (optionally ReLU): forward, data gradient and weight gradient (+ bias gradient from the same pass
    over ``dy``) all on the tcgen05 3xTF32 GEMMs.  No library GEMM: unsupported shapes raise ``DC_EUNSUPPORTED``."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        shp = x.shape
        x2 = _f32c(x.detach()).reshape(-1, shp[-1])
        w = _f32c(weight.detach())
        y = gemm_tf32x3(x2, w, None if bias is None else _f32c(bias.detach()), relu=relu)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x2, w, y if relu else None)
        return y.view(*shp[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        N, K = w.shape
        dy2 = _f32c(dy).reshape(-1, N)
        if ctx.relu:
            dy2 = torch.ops.aten.threshold_backward(dy2, y, 0.0)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = gemm_tf32x3(dy2, w.t().contiguous()).view(*dy.shape[:-1], K)
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw, db = gemm_wgrad_tf32x3(dy2, x2, want_bias=ctx.has_bias)     # dW = dy^T x, db = colsum(dy), one pass over dy
        return dx, dw, db, None


def linear(x, weight, bias=None, relu=False):
    """Dense layer on the tcgen05 3xTF32 GEMM (out features % 128 == 0, in features % 32 == 0; CUDA tensors only)."""
    _need_cuda(x, weight, bias)
    return LinearTC.apply(x, weight, bias, relu)


_wgrad_ws = {}


def gemm_wgrad_supported(T, No, Ni):
    return T > 0 and No % 128 == 0 and Ni % 128 == 0


def gemm_wgrad_tf32x3(dy, x, want_bias=True, dw_out=None, db_out=None, accumulate=False):
    """``dW[No,Ni] = dy[T,No]^T @ x[T,Ni]`` and ``db[No] = dy.sum(0)`` on tcgen05 (3xTF32, split-K, deterministic).

    ``dy`` / ``x`` are 2-D fp32 CUDA tensors with contiguous rows (column-slice views allowed)."""
    _need_cuda(dy, x)
    assert dy.dim() == 2 and x.dim() == 2 and dy.shape[0] == x.shape[0]
    if dy.stride(1) != 1:
        dy = dy.contiguous()
    if x.stride(1) != 1:
        x = x.contiguous()
    T, No = dy.shape
    Ni = x.shape[1]
    dev = dy.device
    if dw_out is None:
        dw_out = torch.empty((No, Ni), dtype=torch.float32, device=dev)
    if want_bias and db_out is None:
        db_out = torch.empty(No, dtype=torch.float32, device=dev)
    lib = _lib.load()
    key = (No, Ni, dev)
    ws = _wgrad_ws.get(key)
    if ws is None:
        ws = torch.empty(int(lib.dc_gemm_wgrad_workspace_bytes(No, Ni)), dtype=torch.uint8, device=dev)
        _wgrad_ws[key] = ws
    with PROFILE.span("gemm_wgrad", 2, 4 * (T * No + T * Ni + No * Ni)):
        _lib.check(lib.dc_gemm_wgrad_tf32x3(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), T, No, Ni,
                                            dw_out.data_ptr(), dw_out.stride(0), _lib.ptr(db_out) if want_bias else None,
                                            1 if accumulate else 0, ws.data_ptr(), _lib.stream_ptr()),
                   "dc_gemm_wgrad_tf32x3")
    return dw_out, (db_out if want_bias else None)


def select_actions(logits, masks, u):
    """Hierarchical action selection for ``A`` agents in one launch (``policy.py:190-216``).

    ``logits`` / ``masks``: five ``[A, n_h]`` tensors in ``HEAD_KEYS`` order (rows may be strided column slices);
    ``u``: ``[A, 5]`` uniforms in [0, 1).  Returns ``(chosen [A,5] int32, logp [A,5] fp32)``; ``chosen`` is -1 for the
    sub-heads the sampled enum does not use.  The index function is pinned to ``oracle.ref_policy.sample_index``."""
    _need_cuda(u, *logits, *masks)
    A = u.shape[0]
    ls, lds, ms = [], [], []
    for h, (l, m) in enumerate(zip(logits, masks)):
        l2 = l.reshape(A, HEAD_SIZES[h]).float()
        if l2.stride(1) != 1:
            l2 = l2.contiguous()
        ls.append(l2)
        lds.append(l2.stride(0))
        ms.append(m.reshape(A, HEAD_SIZES[h]).contiguous().view(torch.uint8))
    u2 = _f32c(u).reshape(A, 5)
    chosen = torch.empty((A, 5), dtype=torch.int32, device=u.device)
    logp = torch.empty((A, 5), dtype=torch.float32, device=u.device)
    with PROFILE.span("select_actions", 1):
        _lib.check(_lib.load().dc_select_actions(_lib.ptr5(ls), (ctypes.c_int64 * 5)(*lds), _lib.ptr5(ms), u2.data_ptr(), A,
                                                 chosen.data_ptr(), logp.data_ptr(), _lib.stream_ptr()), "dc_select_actions")
    return chosen, logp


# --------------------------------------------------------------------------------------------- packed small heads
PACK_COLS = {"enum": (0, 4), "x": (4, 13), "y": (13, 22), "ability": (22, 25), "value": (25, 26)}   # columns of the packed GEMM
PACK_WIDTH = 128


def ppo_loss_packed(packed, logits_tu, masks, actions, old_logp, adv_raw, ret, e_clip, entropy_coef, vf_coef):
    """Fused PPO loss where the four small heads and the value head are column ranges of ONE packed ``[N,128]``
    tensor-core GEMM output (``PACK_COLS``) and the target-unit logits are a separate ``[N,40]`` tensor.

    Returns (out[16], n_actions[5], d_packed [N,128], d_logits_tu [N,40]): the gradients go straight back into the two
    producers, so no slice/cat kernels run and the five tiny K=131072 weight-gradient GEMMs become one tcgen05 wgrad.
    """
    _need_cuda(packed, logits_tu)
    p2 = _f32c(packed.detach()).reshape(-1, PACK_WIDTH)
    N = p2.shape[0]
    tu = _f32c(logits_tu.detach()).reshape(N, 40)
    masks = [_u8(m) for m in masks]
    actions = [_u8(a) for a in actions]
    old_logp, adv_raw, ret = _f32c(old_logp), _f32c(adv_raw), _f32c(ret)
    dev = p2.device
    d_packed = torch.zeros_like(p2)            # the 102 padding columns must carry a zero gradient
    d_tu = torch.empty_like(tu)
    out = torch.empty(_lib.LOSS_SLOTS, dtype=torch.float32, device=dev)
    n_actions = torch.empty(5, dtype=torch.int32, device=dev)
    ws = torch.empty(_lib.PPO_WORKSPACE_BYTES, dtype=torch.uint8, device=dev)

    def col(t, key):
        return t.data_ptr() + 4 * PACK_COLS[key][0]
    c = _lib._c
    lptr = _lib._ptr5(col(p2, "enum"), col(p2, "x"), col(p2, "y"), tu.data_ptr(), col(p2, "ability"))
    dptr = _lib._ptr5(col(d_packed, "enum"), col(d_packed, "x"), col(d_packed, "y"), d_tu.data_ptr(), col(d_packed, "ability"))
    ld = (c.c_int64 * 5)(PACK_WIDTH, PACK_WIDTH, PACK_WIDTH, 40, PACK_WIDTH)
    lib = _lib.load()
    with PROFILE.span("ppo_loss", 2):
        _lib.check(lib.dc_ppo_loss_fwd_bwd_strided(lptr, ld, _lib.ptr5(masks), _lib.ptr5(actions), old_logp.data_ptr(),
                                                   adv_raw.data_ptr(), ret.data_ptr(), col(p2, "value"), PACK_WIDTH, N,
                                                   float(e_clip), float(entropy_coef), float(vf_coef), dptr, ld,
                                                   col(d_packed, "value"), PACK_WIDTH, out.data_ptr(), n_actions.data_ptr(),
                                                   ws.data_ptr(), _lib.stream_ptr()), "dc_ppo_loss_fwd_bwd_strided")
    return out, n_actions, d_packed.view_as(packed), d_tu.view_as(logits_tu)
