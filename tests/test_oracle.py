"""The oracle (oracle/clip_oracle.py) against the golden vectors recorded from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import clip_oracle as orc

W1 = ["clip_w1_n24_d16_f64.npz", "clip_w1_n37_d64_s100_f64.npz", "clip_w1_n130_d72_f64.npz", "clip_w1_n1_d8_f64.npz"]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("name", W1)
def test_clip_loss_matches_reference_fp64(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    loss = orc.clip_loss(z["I"], z["T"], z["scale"])
    assert loss.shape == z["loss"].shape
    assert rel(loss, z["loss"]) < 1e-12
    dI, dT, ds = orc.clip_loss_grads(z["I"], z["T"], z["scale"], z["g"])
    for got, want in ((dI, z["dI"]), (dT, z["dT"]), (ds, z["ds"])):
        assert rel(got, want) < 1e-10


def test_reference_fp32_run_is_within_fp32_noise_of_oracle(golden_dir):
    z = np.load(os.path.join(golden_dir, "clip_w1_n24_d16_f32.npz"))
    loss = orc.clip_loss(z["I"], z["T"], z["scale"])
    assert rel(loss, z["loss"]) < 2e-5
    dI, dT, ds = orc.clip_loss_grads(z["I"], z["T"], z["scale"], z["g"])
    assert rel(dI, z["dI"]) < 1e-4 and rel(dT, z["dT"]) < 1e-4 and rel(ds, z["ds"]) < 1e-4


@pytest.mark.parametrize("name", ["clip_w2_n24_d16.npz", "clip_w2_n264_d64.npz"])
@pytest.mark.parametrize("local_loss", [False, True])
@pytest.mark.parametrize("gwg", [False, True])
def test_distributed_semantics_match_reference(golden_dir, name, local_loss, gwg):
    z = np.load(os.path.join(golden_dir, name))
    n = z["I"].shape[0]
    b = n // 2
    Ib = [z["I"][:b], z["I"][b:]]
    Tb = [z["T"][:b], z["T"][b:]]
    tag = f"ll{int(local_loss)}_gwg{int(gwg)}"
    for r in range(2):
        # gathered ordering is bit-exact: rank-major concatenation
        assert np.array_equal(orc.gather_features(Ib), z[f"{tag}_r{r}_gathered_I"])
        assert np.array_equal(orc.gather_features(Tb), z[f"{tag}_r{r}_gathered_T"])
        loss = orc.clip_loss_distributed(Ib, Tb, z["scale"], r, local_loss)
        assert rel(loss, z[f"{tag}_r{r}_loss"]) < 1e-12
        g = z["g"][:b] if local_loss else z["g"]
        dI, dT, ds = orc.clip_loss_distributed_grads(Ib, Tb, z["scale"], r, local_loss, gwg, g)
        assert rel(dI, z[f"{tag}_r{r}_dI"]) < 1e-10
        assert rel(dT, z[f"{tag}_r{r}_dT"]) < 1e-10
        assert rel(ds, z[f"{tag}_r{r}_ds"]) < 1e-10


@pytest.mark.parametrize("name", ["ce_n20_c7_d16.npz", "ce_n150_c182_d64.npz"])
def test_ce_head_matches_reference(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    imgn, txtn = orc.l2_normalize(z["img"]), orc.l2_normalize(z["txt"])
    assert rel(imgn, z["imgn"]) < 1e-14 and rel(txtn, z["txtn"]) < 1e-14
    loss = orc.cross_entropy(imgn, txtn, z["scale"], z["labels"])
    assert rel(loss, z["loss"]) < 1e-12
    assert abs(loss.mean() - z["loss_mean"]) < 1e-12 * max(1.0, abs(z["loss_mean"]))
    dA, dB, ds = orc.cross_entropy_grads(imgn, txtn, z["scale"], z["labels"], z["g"])
    assert rel(dA, z["d_imgn"]) < 1e-10 and rel(dB, z["d_txtn"]) < 1e-10 and rel(ds, z["ds"]) < 1e-10
    assert rel(orc.l2_normalize_bwd(z["img"], dA), z["d_img"]) < 1e-10
    assert rel(orc.l2_normalize_bwd(z["txt"], dB), z["d_txt"]) < 1e-10
    assert np.array_equal(orc.argmax_predictions(imgn, txtn), z["pred"])


def test_l2norm_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "l2norm.npz"))
    assert rel(orc.l2_normalize(z["x"]), z["y"]) < 1e-14
    assert rel(orc.l2_normalize_bwd(z["x"], z["dy"]), z["dx"]) < 1e-12


def test_scale_gradient_identities():
    # ds = sum_i <dI_i, I_i> / s = sum_j <dT_j, T_j> / s
    rng = np.random.default_rng(0)
    I, T = orc.l2_normalize(rng.standard_normal((33, 24))), orc.l2_normalize(rng.standard_normal((33, 24)))
    g = rng.random(33)
    dI, dT, ds = orc.clip_loss_grads(I, T, 14.3, g)
    assert abs(np.sum(dI * I) / 14.3 - ds) < 1e-12
    assert abs(np.sum(dT * T) / 14.3 - ds) < 1e-12


LABELED = ["labeled_n24_d16_c5.npz", "labeled_n150_d64_c9.npz", "labeled_n40_d32_c40_s30.npz"]


@pytest.mark.parametrize("name", LABELED)
@pytest.mark.parametrize("variant", ["soft", "ignore", "google"])
def test_label_aware_variants_match_reference(golden_dir, name, variant):
    # clip/loss.py:123-192; the google_sup gradients come from the out-of-place restatement (make_golden.py docstring)
    z = np.load(os.path.join(golden_dir, name))
    loss = orc.labeled_clip_loss(z["I"], z["T"], z["scale"], z["y"], variant)
    assert abs(loss - z[f"{variant}_loss"]) < 1e-11 * abs(z[f"{variant}_loss"])
    dI, dT, ds = orc.labeled_clip_loss_grads(z["I"], z["T"], z["scale"], z["y"], variant)
    assert rel(dI, z[f"{variant}_dI"]) < 1e-10 and rel(dT, z[f"{variant}_dT"]) < 1e-10
    assert rel(ds, z[f"{variant}_ds"]) < 1e-10


def test_label_aware_variants_reduce_to_the_default_loss_for_distinct_labels():
    rng = np.random.default_rng(3)
    I, T = orc.l2_normalize(rng.standard_normal((21, 12))), orc.l2_normalize(rng.standard_normal((21, 12)))
    want = orc.clip_loss(I, T, 14.3).mean()
    y = rng.permutation(21)
    for v in ("soft", "ignore"):
        assert abs(orc.labeled_clip_loss(I, T, 14.3, y, v) - want) < 1e-12
