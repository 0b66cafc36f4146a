"""Drop-in replacement for joliang17/FLYP ``clip/loss.py``: ``ClipLoss`` and ``gather_features``.

Same constructor, same ``forward(image_features, text_features, logit_scale, ...)`` signature, same per-sample loss
vector (``reduction='none'``, clip/loss.py:208-211), same gathered ordering (rank-major, clip/loss.py:66-67) and the
same gradient semantics for every ``(local_loss, gather_with_grad)`` combination - but the work is done by the
sm_100a kernels behind ``libflypclip.so``: the B x B logit matrix is never written to memory, and across ranks only
O(B) statistics are exchanged in addition to the feature all-gather the reference also performs.

Differences that are deliberate and documented:
  * ``use_horovod=True`` raises (NCCL / torch.distributed only, no multi-backend dispatch);
  * the label-aware variants (``ground_labels``, ``ignore``, ``google_sup_loss``; clip/loss.py:123-192) are implemented
    for world_size == 1 (flyp_b200/labeled.py) and, like the reference's, return a scalar;
  * new opt-in keyword ``normalize`` (default False) L2-normalises both inputs first with a fused kernel
    (what the callers do at clip/model.py:375-376);
  * there is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from ._lib import FlypError

try:
    import torch.distributed as dist
    has_distributed = dist.is_available()
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False


# ---------------------------------------------------------------------------------------------------- normalise
class _L2Normalize(torch.autograd.Function):
    """x / ||x||_2 row-wise, no epsilon (clip/model.py:375-376)."""

    @staticmethod
    def forward(ctx, x):
        y, inv = ops.l2norm_fwd(x)
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        return ops.l2norm_bwd(y, dy, inv)


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    return _L2Normalize.apply(x)


# ---------------------------------------------------------------------------------------------------- gather
def _all_gather_rows(x: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """Rank-major concatenation of equally sized [b, D] blocks, written in place by one collective (no torch.cat)."""
    x = x.contiguous()
    out = torch.empty((world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    return out


class _GatherWithGrad(torch.autograd.Function):
    """all_gather whose backward is reduce-scatter(SUM), as torch.distributed.nn.all_gather (clip/loss.py:48-52)."""

    @staticmethod
    def forward(ctx, x, world_size, group):
        ctx.world_size, ctx.group = world_size, group
        return _all_gather_rows(x, world_size, group)

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        out = torch.empty((grad.shape[0] // ctx.world_size,) + tuple(grad.shape[1:]), dtype=grad.dtype,
                          device=grad.device)
        dist.reduce_scatter_tensor(out, grad, op=dist.ReduceOp.SUM, group=ctx.group)
        return out, None, None


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=False, rank=0, world_size=1,
                    use_horovod=False, group=None):
    """clip/loss.py:19-69.  Returns (all_image_features, all_text_features), global row = rank * b + local row.
    gather_with_grad=False and not local_loss: the local slot carries the gradient of the local tensor (:62-65)."""
    assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
    if use_horovod:
        raise NotImplementedError("flyp_b200 gathers with NCCL through torch.distributed only (use_horovod=True is "
                                  "not supported)")
    if gather_with_grad:
        all_image = _GatherWithGrad.apply(image_features, world_size, group)
        all_text = _GatherWithGrad.apply(text_features, world_size, group)
    else:
        with torch.no_grad():
            all_image = _all_gather_rows(image_features, world_size, group)
            all_text = _all_gather_rows(text_features, world_size, group)
        if not local_loss:
            b = image_features.shape[0]
            all_image = _SpliceLocal.apply(all_image, image_features, rank * b)
            all_text = _SpliceLocal.apply(all_text, text_features, rank * b)
    return all_image, all_text


class _SpliceLocal(torch.autograd.Function):
    """Value of ``gathered`` (already containing the local block); gradient flows to ``local`` only."""

    @staticmethod
    def forward(ctx, gathered, local, start):
        ctx.start, ctx.n = start, local.shape[0]
        return gathered.view_as(gathered)

    @staticmethod
    def backward(ctx, grad):
        return None, grad[ctx.start:ctx.start + ctx.n], None


# ---------------------------------------------------------------------------------------------------- one-sided CE
class _ContrastiveCE(torch.autograd.Function):
    """loss[i] = CE(scale * a_i @ b.T, target_i), reduction='none' (src/models/ce_ablation.py:122-123 and the
    local_loss blocks clip/loss.py:109-111).  Logits are never materialised."""

    @staticmethod
    def forward(ctx, a, b, scale, labels, label_offset, grad_dtype):
        s = ops._scale_tensor(scale, a.device)
        loss, lse = ops.ce_fwd(a, b, s, labels, label_offset)
        ctx.save_for_backward(a, b, s, lse, loss, labels if labels is not None else torch.empty(0))
        ctx.has_labels = labels is not None
        ctx.label_offset = label_offset
        ctx.grad_dtype = grad_dtype
        ctx.scale_is_tensor = torch.is_tensor(scale)
        ctx.scale_shape = scale.shape if torch.is_tensor(scale) else None
        ctx.scale_dtype = scale.dtype if torch.is_tensor(scale) else None
        return loss.to(a.dtype)

    @staticmethod
    def backward(ctx, g):
        a, b, s, lse, loss, labels = ctx.saved_tensors
        need_a, need_b, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d_a, d_b, d_s = ops.ce_bwd(a, b, s, labels if ctx.has_labels else None, ctx.label_offset, lse, loss, g,
                                   grad_dtype=ctx.grad_dtype, need_a=need_a, need_b=need_b, need_scale=need_s)
        gs = None
        if need_s and ctx.scale_is_tensor:
            gs = torch.zeros(ctx.scale_shape, dtype=torch.float32, device=a.device).reshape(-1)
            gs[:1] = d_s
            gs = gs.reshape(ctx.scale_shape).to(ctx.scale_dtype)
        return (d_a if need_a else None), (d_b if need_b else None), gs, None, None, None


def contrastive_cross_entropy(a, b, logit_scale, labels=None, label_offset=0, reduction="none", grad_dtype=None):
    """Fused ``F.cross_entropy(logit_scale * a @ b.T, labels, reduction=...)``; ``labels=None`` means
    ``arange(n) + label_offset``."""
    loss = _ContrastiveCE.apply(a, b, logit_scale, labels, label_offset, grad_dtype)
    if reduction == "none":
        return loss
    if reduction == "mean":
        if labels is not None:
            # F.cross_entropy(reduction='mean') averages over the targets that are not ignore_index (-100)
            return (loss.float().sum() / (labels != -100).sum()).to(loss.dtype)
        return loss.float().mean().to(loss.dtype)
    if reduction == "sum":
        return loss.float().sum().to(loss.dtype)
    raise ValueError(f"unknown reduction {reduction!r}")


# ---------------------------------------------------------------------------------------------------- symmetric loss
class _ClipLossFn(torch.autograd.Function):
    """Symmetric loss over the row block of this rank.  ``img`` are the local rows, ``txt_all`` / ``img_all`` the
    gathered matrices (for world_size == 1 they are the inputs themselves)."""

    @staticmethod
    def forward(ctx, img, txt, scale, rank, world_size, group, gather_with_grad, grad_dtype):
        dev = img.device
        s = ops._scale_tensor(scale, dev)
        b = img.shape[0]
        if world_size > 1:
            img_all = _all_gather_rows(img, world_size, group)
            txt_all = _all_gather_rows(txt, world_size, group)
            off = rank * b
        else:
            img_all, txt_all, off = img, txt, 0
        n = txt_all.shape[0]
        row_lse, row_nll, col_stat, status = ops.clip_fwd_local(img, txt_all, s, off)
        if world_size > 1:
            col_stat_all = _all_gather_rows(col_stat, world_size, group)
            packed = _all_gather_rows(torch.stack([row_lse, row_nll]), world_size, group)  # [W*2, b]
            packed = packed.view(world_size, 2, b)
            row_lse_all = packed[:, 0].reshape(-1).contiguous()
            row_nll_all = packed[:, 1].reshape(-1).contiguous()
        else:
            col_stat_all, row_lse_all, row_nll_all = col_stat, row_lse, row_nll
        # loss of every global row (the reference returns the full vector on every rank, clip/loss.py:113-114,208)
        col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat_all, world_size, row_nll_all, n, 0, loss_dtype=img.dtype)
        ctx.save_for_backward(img, txt, img_all, txt_all, s, row_lse_all, row_nll_all, col_lse, col_nll)
        ctx.meta = (rank, world_size, group, gather_with_grad, grad_dtype, off, b,
                    torch.is_tensor(scale), scale.shape if torch.is_tensor(scale) else None,
                    scale.dtype if torch.is_tensor(scale) else None)
        ctx.status = status
        return loss

    @staticmethod
    def backward(ctx, g):
        img, txt, img_all, txt_all, s, row_lse_all, row_nll_all, col_lse, col_nll = ctx.saved_tensors
        rank, world_size, group, gwg, grad_dtype, off, b, s_is_tensor, s_shape, s_dtype = ctx.meta
        need_img, need_txt, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        g = g.to(torch.float32).contiguous()
        mul = float(world_size) if gwg else 1.0
        d_img = d_txt = d_s = None
        if world_size == 1:
            d_img, d_txt, d_s = ops.clip_bwd_local(img, txt, s, 0, row_lse_all, row_nll_all, col_lse, col_nll, g, g,
                                                   grad_mul=mul, grad_dtype=grad_dtype, need_img=need_img,
                                                   need_txt=need_txt, need_scale=need_s)
        else:
            sl = slice(off, off + b)
            # image rows of this rank against all texts: complete d_img; the scale gradient of this row block
            if need_img or need_s:
                d_img, _, d_s = ops.clip_bwd_local(img, txt_all, s, off, row_lse_all[sl].contiguous(),
                                                   row_nll_all[sl].contiguous(), col_lse, col_nll, g[sl].contiguous(), g,
                                                   grad_mul=mul, grad_dtype=grad_dtype, need_img=True, need_txt=False,
                                                   need_scale=need_s)
            # text rows of this rank against all images (the transposed problem): complete d_txt
            if need_txt:
                d_txt, _, _ = ops.clip_bwd_local(txt, img_all, s, off, col_lse[sl].contiguous(),
                                                 col_nll[sl].contiguous(), row_lse_all, row_nll_all,
                                                 g[sl].contiguous(), g, grad_mul=mul, grad_dtype=grad_dtype,
                                                 need_img=True, need_txt=False, need_scale=False)
            if need_s:
                # every rank differentiates the same replicated loss: d(scale) sums the row blocks of all ranks
                dist.all_reduce(d_s, op=dist.ReduceOp.SUM, group=group)
        gs = None
        if need_s and s_is_tensor:
            if s_dtype == torch.float32 and len(s_shape) <= 1 and (len(s_shape) == 0 or s_shape[0] == 1):
                gs = d_s.view(s_shape)
            else:
                gs = torch.zeros(s_shape, dtype=torch.float32, device=img.device).reshape(-1)
                gs[:1] = d_s
                gs = gs.reshape(s_shape).to(s_dtype)
        return (d_img if need_img else None), (d_txt if need_txt else None), gs, None, None, None, None, None


class _ClipLossStepFn(torch.autograd.Function):
    """The symmetric loss through the whole-step C entry points (flyp_b200/step.py): one call per direction.
    ``comm=None``: the single-GPU loss of the FLYP loop.  With a PeerComm: the row-sharded loss with every exchange done
    over peer memory (flyp_b200/comm.py) - feature pushes by the copy engines (one multicast copy per matrix on
    NVSwitch) that overlap the forward kernel, flag-polling tensor-core kernels, statistics and d(scale) by (multicast)
    remote stores."""

    @staticmethod
    def forward(ctx, img, txt, scale, comm, gather_with_grad, grad_dtype):
        from . import step
        s = ops._scale_tensor(scale, img.device)
        need_bwd = any(ctx.needs_input_grad[:3])
        loss, st = step.step_forward(comm, img, txt, s, img.dtype, need_backward=need_bwd)   # loss in the feature dtype
        ctx.st = st
        ctx.meta = (gather_with_grad, grad_dtype, torch.is_tensor(scale), scale.shape if torch.is_tensor(scale) else None,
                    scale.dtype if torch.is_tensor(scale) else None)
        return loss

    @staticmethod
    def backward(ctx, g):
        from . import step
        st = ctx.st
        gwg, grad_dtype, s_is_tensor, s_shape, s_dtype = ctx.meta
        need_img, need_txt, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        mul = float(st.world) if gwg else 1.0
        # every rank differentiates the same replicated loss: d(scale) is the sum over the row blocks of all ranks
        d_img, d_txt, tot = step.step_backward(st, g, mul, grad_dtype, need_img, need_txt, need_s)
        gs = None
        if need_s and s_is_tensor:
            if s_dtype == torch.float32 and len(s_shape) <= 1 and (len(s_shape) == 0 or s_shape[0] == 1):
                gs = tot.view(s_shape)
            else:
                gs = torch.zeros(s_shape, dtype=torch.float32, device=st.img.device).reshape(-1)
                gs[:1] = tot
                gs = gs.reshape(s_shape).to(s_dtype)
        return (d_img if need_img else None), (d_txt if need_txt else None), gs, None, None, None


class _LocalLossPeerFn(torch.autograd.Function):
    """local_loss=True over peer memory (clip/loss.py:109-111,200-201): the loss of the LOCAL rows as two one-directional
    cross-entropies against the gathered matrices, which arrive in this rank's exchange segment while the kernels run.
    gather_with_grad=False: gradients flow to the local operands only (the gathered matrices carry none, :54-61);
    True: the gradients w.r.t. the gathered matrices are reduce-scattered (sum) to their owners, like the backward of
    torch.distributed.nn.all_gather (:48-52) - the one real B x D exchange of this mode, an NCCL collective."""

    @staticmethod
    def forward(ctx, img, txt, scale, comm, gather_with_grad, grad_dtype, group):
        from . import comm as peer
        s = ops._scale_tensor(scale, img.device)
        li, lt, st = peer.local_fwd(comm, img, txt, s)
        ctx.st = st
        ctx.meta = (gather_with_grad, grad_dtype, group, torch.is_tensor(scale),
                    scale.shape if torch.is_tensor(scale) else None, scale.dtype if torch.is_tensor(scale) else None)
        return ((li + lt) * 0.5).to(img.dtype)

    @staticmethod
    def backward(ctx, g):
        from . import comm as peer
        st = ctx.st
        gwg, grad_dtype, group, s_is_tensor, s_shape, s_dtype = ctx.meta
        d_img, d_txt, d_s, d_img_all, d_txt_all = peer.local_bwd(st, g, grad_dtype, gwg)
        if gwg:
            for local, gathered in ((d_img, d_img_all), (d_txt, d_txt_all)):
                part = torch.empty_like(local)
                dist.reduce_scatter_tensor(part, gathered.contiguous(), op=dist.ReduceOp.SUM, group=group)
                local += part
        gs = None
        if ctx.needs_input_grad[2] and s_is_tensor:
            gs = torch.zeros(s_shape, dtype=torch.float32, device=st.img.device).reshape(-1)
            gs[:1] = d_s
            gs = gs.reshape(s_shape).to(s_dtype)
        return d_img, d_txt, gs, None, None, None, None


class ClipLoss(nn.Module):
    """clip/loss.py:72-211 with the same constructor and forward signature."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False, normalize=False, grad_dtype: Optional[torch.dtype] = None, group=None):
        super().__init__()
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        # extensions (defaults reproduce the reference)
        self.normalize = normalize
        self.grad_dtype = grad_dtype
        self.group = group
        # Multi-rank exchange.  Ranks of ONE node exchange over NVLink peer memory (flyp_b200/comm.py); only ranks
        # spread over several nodes (no peer mapping possible) use torch.distributed collectives.  This is decided once,
        # collectively, on the first forward - it is not a user-facing backend switch.  (FLYP_EXCHANGE=collective forces
        # the multi-node path on one node: tests and A/B measurements.)
        self._peer = None          # PeerComm, or False once the set-up was tried and refused
        self._peer_shape = None

        # cache state (kept for attribute compatibility; the kernels need no label tensor)
        self.prev_num_logits = 0
        self.labels = {}

    def _labels(self, device, num_logits):
        # clip/loss.py:195-206
        if self.prev_num_logits != num_logits or device not in self.labels:
            labels = torch.arange(num_logits, device=device, dtype=torch.long)
            if self.world_size > 1 and self.local_loss:
                labels = labels + num_logits * self.rank
            if self.cache_labels:
                self.labels[device] = labels
                self.prev_num_logits = num_logits
        else:
            labels = self.labels[device]
        return labels

    def _peer_comm(self, feats):
        """The peer-memory communicator for blocks shaped like ``feats`` (created collectively on first use; every rank
        takes the same decision).  None -> the ranks do not share a node: torch.distributed collectives."""
        import os
        if os.environ.get("FLYP_EXCHANGE", "") == "collective" or feats.dtype not in (torch.bfloat16, torch.float32):
            return None
        shape = (feats.shape[0], feats.shape[1])
        if self._peer is not None and self._peer_shape == shape:
            return self._peer or None
        if self._peer:
            # peers may still be copying into this rank's segment (their side streams): drain the device and meet the
            # other ranks before the segment is unmapped
            torch.cuda.synchronize(feats.device)
            dist.barrier(group=self.group)
            self._peer.close()
        from .comm import PeerComm
        self._peer_shape = shape
        self._peer = PeerComm.from_process_group(self.rank, self.world_size, shape[0], shape[1], feats.device,
                                                 self.group) or False
        return self._peer or None

    def forward(self, image_features, text_features, logit_scale, ground_labels=None, ignore=False,
                google_sup_loss=False):
        assert not (ignore and google_sup_loss), 'please specify only one'
        if self.use_horovod and self.world_size > 1:
            raise NotImplementedError("use_horovod=True is not supported (NCCL via torch.distributed only)")
        if not image_features.is_cuda:
            raise FlypError("flyp_b200.ClipLoss needs CUDA tensors (sm_100a); there is no CPU fallback")
        if self.normalize:
            image_features = l2_normalize(image_features)
            text_features = l2_normalize(text_features)
        device = image_features.device
        if ground_labels is not None:
            # clip/loss.py:123-192: the label-aware variants return a scalar.  In the reference they only make sense for
            # world_size == 1 (the labels are local while the gathered logits are global: the shapes do not match).
            if self.world_size > 1:
                raise NotImplementedError("ground_labels with world_size > 1 is ill-defined in the reference "
                                          "(clip/loss.py:124-127 compares local labels with gathered logits)")
            from .labeled import labeled_clip_loss
            return labeled_clip_loss(image_features, text_features, logit_scale, ground_labels, ignore, google_sup_loss,
                                     self.grad_dtype)

        if self.world_size > 1 and self.local_loss:
            # clip/loss.py:109-111: two row blocks against the gathered matrices; loss for the local rows only
            if self.cache_labels:
                self._labels(device, image_features.shape[0])
            peer = self._peer_comm(image_features)
            if peer is not None:
                return _LocalLossPeerFn.apply(image_features, text_features, logit_scale, peer, self.gather_with_grad,
                                              self.grad_dtype, self.group)
            # ranks on several nodes: torch.distributed collectives
            all_image, all_text = gather_features(image_features, text_features, True, self.gather_with_grad,
                                                  self.rank, self.world_size, False, self.group)
            off = self.rank * image_features.shape[0]
            li = contrastive_cross_entropy(image_features, all_text, logit_scale, None, off,
                                           grad_dtype=self.grad_dtype)
            lt = contrastive_cross_entropy(text_features, all_image, logit_scale, None, off,
                                           grad_dtype=self.grad_dtype)
            return (li + lt) / 2

        if self.world_size == 1:
            loss = _ClipLossStepFn.apply(image_features, text_features, logit_scale, None, self.gather_with_grad,
                                         self.grad_dtype)
        else:
            peer = self._peer_comm(image_features)
            if peer is not None:
                loss = _ClipLossStepFn.apply(image_features, text_features, logit_scale, peer, self.gather_with_grad,
                                             self.grad_dtype)
            else:
                loss = _ClipLossFn.apply(image_features, text_features, logit_scale, self.rank, self.world_size,
                                         self.group, self.gather_with_grad, self.grad_dtype)
        if self.cache_labels:
            self._labels(device, loss.shape[0])
        return loss
