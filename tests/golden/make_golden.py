"""Generates tests/golden/*.npz by running the UNMODIFIED reference (joliang17/FLYP, mounted at /root/reference) on CPU.

    python tests/golden/make_golden.py          # needs /root/reference; not needed to run the tests

What is recorded (all inputs are seeded, float64 unless noted):
  clip_w1_*.npz   clip.loss.ClipLoss(world_size=1) forward (per-item vector) and autograd gradients for a random
                  upstream vector g                                       (clip/loss.py:94-211)
  clip_w2_*.npz   the same module under 2 gloo ranks for the four (local_loss, gather_with_grad) combinations:
                  per-rank outputs, gathered features and local gradients (clip/loss.py:19-69,103-114)
  ce_*.npz        the --ce_ablation head: F.cross_entropy(scale * img @ txt.T, labels) with normalised inputs
                                                                           (src/models/ce_ablation.py:115-123)
  l2norm.npz      x / x.norm(dim=-1, keepdim=True) and its autograd        (clip/model.py:375-376)
  labeled_*.npz   ClipLoss(world_size=1)(I, T, s, ground_labels=y [, ignore=True | google_sup_loss=True]): scalar loss
                  and autograd gradients                                   (clip/loss.py:123-192).
                  The reference's google_sup_loss branch modifies the output of torch.exp in place (:166, :178), so its
                  backward raises "modified by an inplace operation": for that variant only the LOSS comes from the
                  unmodified reference; the gradients are autograd through an out-of-place restatement of the same
                  lines (checked here to reproduce the reference's loss to the last bit).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def make_inputs(n, d, seed, dtype=torch.float64, mix=0.5):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=gen, dtype=torch.float64)
    y = torch.randn(n, d, generator=gen, dtype=torch.float64)
    I = torch.nn.functional.normalize(x, dim=-1)
    T = torch.nn.functional.normalize(mix * I + (1 - mix) * torch.nn.functional.normalize(y, dim=-1), dim=-1)
    g = torch.rand(n, generator=gen, dtype=torch.float64) / n
    return I.to(dtype), T.to(dtype), g.to(dtype)


def run_w1(n, d, s, seed, dtype, name, mix=0.5):
    sys.path.insert(0, REF)
    from clip.loss import ClipLoss
    I, T, g = make_inputs(n, d, seed, dtype, mix)
    I.requires_grad_(True); T.requires_grad_(True)
    theta = torch.tensor(float(np.log(s)), dtype=dtype, requires_grad=True)
    fn = ClipLoss(cache_labels=True)
    loss = fn(I, T, theta.exp())
    (loss * g).sum().backward()
    ds = theta.grad / theta.exp()          # d/d(scale) from d/d(theta): scale = exp(theta)
    np.savez(os.path.join(OUT, name), I=I.detach().numpy(), T=T.detach().numpy(), g=g.numpy(), scale=np.float64(s),
             loss=loss.detach().numpy(), dI=I.grad.numpy(), dT=T.grad.numpy(), ds=ds.detach().numpy())


def _w2_worker(rank, world, port, n, d, s, seed, local_loss, gwg, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, REF)
    from clip.loss import ClipLoss, gather_features
    I, T, g = make_inputs(n, d, seed)
    b = n // world
    Il = I[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    Tl = T[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    sc = torch.tensor(float(s), dtype=torch.float64, requires_grad=True)
    fn = ClipLoss(local_loss=local_loss, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world)
    loss = fn(Il, Tl, sc)
    gg = g[:loss.shape[0]] if local_loss else g
    (loss * gg).sum().backward()
    with torch.no_grad():
        gi, gt = gather_features(Il.detach(), Tl.detach(), local_loss, False, rank, world, False)
    ret[rank] = dict(loss=loss.detach().numpy(), dI=Il.grad.numpy(), dT=Tl.grad.numpy(), ds=sc.grad.numpy(),
                     gathered_I=gi.numpy(), gathered_T=gt.numpy())
    dist.destroy_process_group()


def run_w2(n, d, s, seed, port):
    I, T, g = make_inputs(n, d, seed)
    out = dict(I=I.numpy(), T=T.numpy(), g=g.numpy(), scale=np.float64(s))
    for local_loss in (False, True):
        for gwg in (False, True):
            mgr = mp.Manager()
            ret = mgr.dict()
            mp.spawn(_w2_worker, args=(2, port, n, d, s, seed, local_loss, gwg, ret), nprocs=2, join=True)
            port += 1
            tag = f"ll{int(local_loss)}_gwg{int(gwg)}"
            for r in range(2):
                for k, v in ret[r].items():
                    out[f"{tag}_r{r}_{k}"] = v
    np.savez(os.path.join(OUT, f"clip_w2_n{n}_d{d}.npz"), **out)


def run_ce(n, c, d, s, seed, name):
    # src/models/ce_ablation.py:115-123 restated with the reference's own ops (the file itself does not import here)
    gen = torch.Generator().manual_seed(seed)
    img = torch.randn(n, d, generator=gen, dtype=torch.float64, requires_grad=True)
    txt = torch.randn(c, d, generator=gen, dtype=torch.float64, requires_grad=True)
    labels = torch.randint(0, c, (n,), generator=gen)
    g = torch.rand(n, generator=gen, dtype=torch.float64) / n
    sc = torch.tensor(float(s), dtype=torch.float64, requires_grad=True)
    imgn = img / img.norm(dim=-1, keepdim=True)
    txtn = txt / txt.norm(dim=-1, keepdim=True)
    imgn.retain_grad(); txtn.retain_grad()
    logits = sc * imgn @ txtn.T
    loss = torch.nn.functional.cross_entropy(logits, labels, reduction="none")
    (loss * g).sum().backward()
    np.savez(os.path.join(OUT, name), img=img.detach().numpy(), txt=txt.detach().numpy(), labels=labels.numpy(),
             g=g.numpy(), scale=np.float64(s), imgn=imgn.detach().numpy(), txtn=txtn.detach().numpy(),
             loss=loss.detach().numpy(), loss_mean=np.float64(loss.mean().item()), d_imgn=imgn.grad.numpy(),
             d_txtn=txtn.grad.numpy(), d_img=img.grad.numpy(), d_txt=txt.grad.numpy(), ds=sc.grad.numpy(),
             pred=logits.detach().argmax(dim=1).numpy())


def run_l2norm():
    gen = torch.Generator().manual_seed(7)
    x = (3.0 * torch.randn(19, 24, generator=gen, dtype=torch.float64)).requires_grad_(True)
    dy = torch.randn(19, 24, generator=gen, dtype=torch.float64)
    y = x / x.norm(dim=-1, keepdim=True)
    (y * dy).sum().backward()
    np.savez(os.path.join(OUT, "l2norm.npz"), x=x.detach().numpy(), dy=dy.numpy(), y=y.detach().numpy(),
             dx=x.grad.numpy())


def _google_sup_out_of_place(I, T, s, y):
    """clip/loss.py:123-127,160-187 with `a /= b` / `a *= b` written out of place (the only change)."""
    lpi = s * I @ T.T
    lpt = s * T @ I.T
    E = (y.view(1, -1).repeat(I.shape[0], 1) == y.view(-1, 1)).type(torch.float)
    out = []
    for L in (lpi, lpt):
        e = torch.exp(L - torch.max(L, dim=1, keepdim=True).values)
        tot = torch.sum(e, dim=1, keepdim=True).repeat(1, e.shape[1])
        v = -torch.log(e / (tot - e)) * E
        out.append(torch.mean(torch.sum(v, dim=1) / torch.sum(E, dim=1)))
    return (out[0] + out[1]) / 2


def run_labeled(n, d, n_cls, s, seed, name, mix=0.5):
    sys.path.insert(0, REF)
    from clip.loss import ClipLoss
    gen = torch.Generator().manual_seed(seed + 1000)
    y = torch.randint(0, n_cls, (n,), generator=gen) * 7 - 3            # arbitrary (negative, sparse) label values
    out = dict(y=y.numpy(), scale=np.float64(s))
    for variant, kw in (("soft", {}), ("ignore", dict(ignore=True)), ("google", dict(google_sup_loss=True))):
        I, T, _ = make_inputs(n, d, seed, torch.float64, mix)
        I.requires_grad_(True); T.requires_grad_(True)
        sc = torch.tensor(float(s), dtype=torch.float64, requires_grad=True)
        loss = ClipLoss()(I, T, sc, ground_labels=y, **kw)
        assert loss.dim() == 0
        if variant == "google":
            again = _google_sup_out_of_place(I, T, sc, y)
            assert again.item() == loss.item(), (again.item(), loss.item())
            again.backward()
        else:
            loss.backward()
        out.update({"I": I.detach().numpy(), "T": T.detach().numpy(), f"{variant}_loss": np.float64(loss.item()),
                    f"{variant}_dI": I.grad.numpy(), f"{variant}_dT": T.grad.numpy(), f"{variant}_ds": sc.grad.numpy()})
    np.savez(os.path.join(OUT, name), **out)


if __name__ == "__main__":
    assert os.path.isdir(REF), "the reference is not mounted"
    run_w1(24, 16, 1 / 0.07, 1, torch.float64, "clip_w1_n24_d16_f64.npz")
    run_w1(24, 16, 1 / 0.07, 1, torch.float32, "clip_w1_n24_d16_f32.npz")
    run_w1(37, 64, 100.0, 2, torch.float64, "clip_w1_n37_d64_s100_f64.npz", mix=0.15)
    run_w1(130, 72, 1 / 0.07, 3, torch.float64, "clip_w1_n130_d72_f64.npz")
    run_w1(1, 8, 1 / 0.07, 4, torch.float64, "clip_w1_n1_d8_f64.npz")
    run_w2(24, 16, 1 / 0.07, 5, 29611)
    run_w2(264, 64, 1 / 0.07, 6, 29631)
    run_ce(20, 7, 16, 1 / 0.07, 8, "ce_n20_c7_d16.npz")
    run_ce(150, 182, 64, 100.0, 9, "ce_n150_c182_d64.npz")
    run_l2norm()
    run_labeled(24, 16, 5, 1 / 0.07, 11, "labeled_n24_d16_c5.npz")
    run_labeled(150, 64, 9, 1 / 0.07, 12, "labeled_n150_d64_c9.npz")
    run_labeled(40, 32, 40, 30.0, 13, "labeled_n40_d32_c40_s30.npz", mix=0.3)        # most classes are singletons
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))
