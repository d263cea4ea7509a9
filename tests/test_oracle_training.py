"""CPU: pins the oracle's model/loss/optimizer restatement against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden_train.py -> tests/golden/train_golden.npz).  fp32 vs fp32 on the same torch
build: tolerance 2e-5 relative (different op grouping only)."""
import os

import numpy as np
import pytest
import torch

from oracle import fedcola_oracle as O
import helpers as H

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_golden.npz"))
UPDATE_RUNS = {"sgd": ("SGD", 0.05, None, 0.0), "adamw": ("AdamW", 1e-3, None, 0.0),
               "prox_sgd_clip": ("SGD", 0.05, 0.1, 1.0)}


def oracle_params(kind):
    spec = H.train_spec(kind)
    sd = H.state_dict_of(spec, H.fill_arena(spec, 7))
    params, seen = {}, {}
    for s in spec.segments:
        root = s.alias_of or s.key
        if root not in seen:
            seen[root] = torch.from_numpy(sd[root].copy())
        params[s.key] = seen[root]
    return spec, params


def batches_of(ds, n, seed, B):
    a, b = H.make_samples(ds, n, seed)
    return [(a[i:i + B], b[i:i + B]) for i in range(0, n, B)]


@pytest.mark.parametrize("kind", sorted(H.TRAIN_KINDS))
def test_step0_matches_reference(kind):
    spec, params = oracle_params(kind)
    ds, _ = H.TRAIN_KINDS[kind]
    m = H.DS_MODALITY[ds]
    for p in params.values():
        p.requires_grad_(True)
    batch = H.make_samples(ds, 4, 11)
    if m == "img+txt":
        outs = O.mat_forward(params, [batch[0], batch[1]], spec.modalities, spec.num_heads, spec.depth, feat_out=True)
        loss = O.contrastive_loss(*outs)
        out = torch.cat(outs, 0)
    else:
        loss, out = O.client_loss(params, batch, m, spec.modalities, spec.num_heads, spec.depth)
    loss.backward()
    np.testing.assert_allclose(out.detach().numpy(), GOLD[f"{kind}/step0/out"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(loss.item(), GOLD[f"{kind}/step0/loss"], rtol=2e-6)
    for s in spec.unique_segments():
        g = params[s.key].grad
        g = g if g is not None else torch.zeros_like(params[s.key])
        ref = GOLD[f"{kind}/step0/g:{s.key}"]
        scale = max(float(GOLD[f"{kind}/step0/gn:{s.key}"]), 1e-12)
        np.testing.assert_allclose(H.subsample(g.numpy()), ref, rtol=1e-4, atol=2e-6 * scale + 1e-9, err_msg=s.key)


@pytest.mark.parametrize("run", sorted(UPDATE_RUNS))
@pytest.mark.parametrize("kind", sorted(H.TRAIN_KINDS))
def test_client_update_matches_reference(kind, run):
    if f"{kind}/{run}/loss" not in GOLD:
        pytest.skip("combination not in the golden set")
    opt, lr, mu, clip = UPDATE_RUNS[run]
    spec, params = oracle_params(kind)
    ds, _ = H.TRAIN_KINDS[kind]
    m = H.DS_MODALITY[ds]
    losses, res = O.client_update(params, {}, batches_of(ds, 8, 21, 4), m, spec.modalities, spec.num_heads, spec.depth,
                                  E=1, optimizer=opt, lr=lr, momentum=0.9 if opt == "SGD" else 0.0, max_grad_norm=clip,
                                  mu=mu, n_total=8)
    np.testing.assert_allclose(res[1]["loss"], GOLD[f"{kind}/{run}/loss"], rtol=2e-5)
    n_bad = n_all = 0
    for s in spec.unique_segments():
        got, ref = H.subsample(params[s.key].detach().numpy()), GOLD[f"{kind}/{run}/p:{s.key}"]
        if opt == "AdamW":
            # Adam's update is ~lr*sign(g) wherever |g| ~ eps (e.g. the key bias of qkv, whose true gradient is
            # exactly zero): a last-bit gradient difference moves such an element by up to 2*lr per step.
            # Allow <1% of all elements to differ, each by at most 2 steps * 2 lr.
            n_bad += int((np.abs(got - ref) > (2e-6 + 2e-4 * np.abs(ref))).sum())
            n_all += got.size
            assert np.abs(got - ref).max() <= 4.2 * lr, s.key
        else:
            np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-6, err_msg=s.key)
    assert n_all == 0 or n_bad / n_all < 0.01
    # upload(): merged keys only
    up = O.upload_merge({k: v.detach().numpy() for k, v in params.items()}, H.TRAIN_KINDS[kind][1], m)
    assert sorted(up.keys()) == list(GOLD[f"{kind}/{run}/upload_keys"])
