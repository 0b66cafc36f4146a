"""Bring-up probe for the raw C ABI on a B200 (run under gpurun).  Not part of the product or the test-suite:
prints error metrics for debug logits, forward statistics and backward gradients against torch fp64 math."""
import ctypes
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flyp_b200 import _lib

torch.manual_seed(0)
lib = _lib.load()
dev = torch.device("cuda:0")


def ws_for(n_rows, n_cols, dim):
    sz = ctypes.c_size_t()
    _lib.check(lib.flyp_clip_workspace_bytes(n_rows, n_cols, dim, 0, ctypes.byref(sz)))
    return torch.empty(sz.value, dtype=torch.uint8, device=dev), sz.value


def make(n, d, corr=True):
    x = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=-1)
    y = torch.nn.functional.normalize(torch.randn(n, d, device=dev), dim=-1)
    t = torch.nn.functional.normalize(0.5 * x + 0.5 * y, dim=-1) if corr else y
    return x.bfloat16().contiguous(), t.bfloat16().contiguous()


def ref(I, T, s, g):
    I = I.double().requires_grad_(True)
    T = T.double().requires_grad_(True)
    sc = torch.tensor(float(s), dtype=torch.float64, device=dev, requires_grad=True)
    S = sc * I @ T.t()
    n = S.shape[0]
    lab = torch.arange(n, device=dev)
    loss = 0.5 * (torch.nn.functional.cross_entropy(S, lab, reduction="none") +
                  torch.nn.functional.cross_entropy(S.t(), lab, reduction="none"))
    (loss * g.double()).sum().backward()
    return loss.detach(), I.grad, T.grad, sc.grad, S.detach()


def relerr(a, b):
    a = a.double(); b = b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def run(n, d, s, corr=True, check_logits=True):
    I, T = make(n, d, corr)
    scale = torch.tensor([s], dtype=torch.float32, device=dev)
    ws, wsz = ws_for(n, n, d)
    st = torch.cuda.current_stream().cuda_stream
    g = torch.rand(n, device=dev, dtype=torch.float32) / n
    lref, dIref, dTref, dsref, Sref = ref(I, T, s, g)
    if check_logits:
        out = torch.zeros(n, n, dtype=torch.float32, device=dev)
        _lib.check(lib.flyp_debug_logits(I.data_ptr(), T.data_ptr(), n, n, d, 0, out.data_ptr(), ws.data_ptr(), wsz, st))
        torch.cuda.synchronize()
        print(f"[n={n} d={d}] logits relerr {relerr(out * s, Sref):.3e}  argmax-eq {(out.argmax(1) == Sref.argmax(1)).float().mean().item():.4f}")
    row_lse = torch.zeros(n, device=dev); col_stat = torch.zeros(3 * n, device=dev); diag = torch.zeros(n, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    col_lse = torch.zeros(n, device=dev); loss = torch.zeros(n, device=dev)
    _lib.check(lib.flyp_clip_fwd_local(I.data_ptr(), T.data_ptr(), scale.data_ptr(), n, n, d, 0, 0, row_lse.data_ptr(),
                                       diag.data_ptr(), col_stat.data_ptr(), status.data_ptr(), ws.data_ptr(), wsz, st))
    _lib.check(lib.flyp_clip_fwd_finish(col_stat.data_ptr(), 1, diag.data_ptr(), n, n, 0,
                                        col_lse.data_ptr(), loss.data_ptr(), st))
    torch.cuda.synchronize()
    print(f"[n={n} d={d} s={s}] status={status.item()} loss relerr {relerr(loss, lref):.3e}  "
          f"row_lse err {relerr(row_lse, torch.logsumexp(Sref, 1)):.3e} col_lse err {relerr(col_lse, torch.logsumexp(Sref, 0)):.3e}")
    dsc = torch.zeros(1, device=dev)
    for gd, gt in ((1, torch.float32), (0, torch.bfloat16)):
        dI = torch.zeros(n, d, dtype=gt, device=dev); dT = torch.zeros(n, d, dtype=gt, device=dev)
        _lib.check(lib.flyp_clip_bwd_local(I.data_ptr(), T.data_ptr(), scale.data_ptr(), n, n, d, 0, 0, row_lse.data_ptr(),
                                           col_lse.data_ptr(), g.data_ptr(), g.data_ptr(), 1.0, gd, dI.data_ptr(), dT.data_ptr(),
                                           dsc.data_ptr(), ws.data_ptr(), wsz, st))
        torch.cuda.synchronize()
        print(f"[n={n} d={d} s={s}] grad {gt}: dI relerr {relerr(dI, dIref):.3e} dT relerr {relerr(dT, dTref):.3e} "
              f"ds relerr {abs(dsc.item() - dsref.item()) / abs(dsref.item()):.3e}")
    return I, T, scale, ws, wsz, row_lse, col_stat, diag, status, col_lse, loss, g, dI, dT, dsc


def bench(n, d, s, iters=10):
    I, T, scale, ws, wsz, row_lse, col_stat, diag, status, col_lse, loss, g, dI, dT, dsc = run(n, d, s, check_logits=False)
    st = torch.cuda.current_stream().cuda_stream

    def fwd():
        lib.flyp_clip_fwd_local(I.data_ptr(), T.data_ptr(), scale.data_ptr(), n, n, d, 0, 0, row_lse.data_ptr(),
                                diag.data_ptr(), col_stat.data_ptr(), status.data_ptr(), ws.data_ptr(), wsz, st)
        lib.flyp_clip_fwd_finish(col_stat.data_ptr(), 1, diag.data_ptr(), n, n, 0,
                                 col_lse.data_ptr(), loss.data_ptr(), st)

    def bwd():
        lib.flyp_clip_bwd_local(I.data_ptr(), T.data_ptr(), scale.data_ptr(), n, n, d, 0, 0, row_lse.data_ptr(),
                                col_lse.data_ptr(), g.data_ptr(), g.data_ptr(), 1.0, 0, dI.data_ptr(), dT.data_ptr(),
                                dsc.data_ptr(), ws.data_ptr(), wsz, st)

    for name, fn, flops in (("fwd", fwd, 2.0 * n * n * d), ("bwd", bwd, 6.0 * n * n * d)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"[bench n={n} d={d}] {name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (algorithmic)")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0))
    if what in ("all", "small"):
        run(128, 64, 14.2857)
        run(128, 512, 14.2857)
        run(256, 512, 14.2857)
        run(512, 512, 14.2857)
        run(37, 512, 14.2857)
        run(300, 768, 100.0)
        run(300, 768, 100.0, corr=False)
        run(512, 512, 14.2857, corr=False)
        run(1024, 1024, 14.2857)
    if what in ("all", "bench"):
        bench(8192, 512, 14.2857)
        bench(32768, 512, 14.2857, iters=3)
