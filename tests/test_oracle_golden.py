"""The CPU oracle against the committed golden vectors (reference outputs, tests/golden/*.npz).

The fixtures were produced by oracle/make_golden.py from the LIVE reference modules; inputs are
regenerated from integer hashes (oracle/synth.py).  Runs without a GPU and without /root/reference.
"""
import os

import numpy as np
import torch

import oracle
from oracle import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_seeded_init_and_flat_layout():
    g = np.load(os.path.join(GOLD, 'init.npz'))
    params = oracle.init_convnet3d(int(g['seed']), 3, 50)
    assert list(params.keys()) == oracle.convnet3d_param_names()
    offs = np.cumsum([0] + [v.numel() for v in params.values()])
    assert np.array_equal(offs, g['offsets'])
    # offsets of the flat student vector
    assert offs.tolist() == [0, 28224, 28288, 1232512, 1232640, 3641088, 3641216, 3647616, 3647666]
    flat = oracle.flatten_params(params)
    s, samp = synth.summarize(flat, 9973)
    assert np.allclose(s, g['flat_sums'], rtol=1e-12) and np.array_equal(samp, g['flat_sample'])
    back = oracle.unflatten_params(flat, params)
    assert all(torch.equal(back[k], params[k]) for k in params)
    hal = oracle.init_hallucinator(int(g['hal_seed']))
    assert np.array_equal(hal['encoder.weight'].numpy(), g['hal_weight'])
    assert np.array_equal(hal['encoder.bias'].numpy(), g['hal_bias'])


def test_convnet3d_embed_forward():
    g = np.load(os.path.join(GOLD, 'convnet3d.npz'))
    params = synth.synth_convnet3d_params(1, num_classes=5)
    x = synth.hash_uniform((3, 8, 3, 64, 64), 11)
    assert rel(oracle.convnet3d_embed(params, x), g['embed']) < 1e-6
    assert rel(oracle.convnet3d_forward(params, x, (64, 64)), g['logits_eval']) < 1e-6
    mask = torch.from_numpy(g['dropout_mask'])
    assert rel(oracle.convnet3d_forward(params, x, (64, 64), dropout_mask=mask), g['logits_train']) < 1e-6
    assert oracle.embed_dim(8, (64, 64)) == 256 and oracle.embed_dim(16, (112, 112)) == 2048
    _, inter = oracle.convnet3d_features(params, x.permute(0, 2, 1, 3, 4), return_intermediates=True)
    for kind, d, t in inter:
        s, samp = synth.summarize(t)
        assert rel(samp, g[f'{kind}{d}_sample']) < 1e-6


def test_composer():
    g = np.load(os.path.join(GOLD, 'composer.npz'))
    hal = synth.synth_hallucinator(3)
    s = synth.hash_uniform((5, 3, 16, 16), 21).requires_grad_(True)
    d = synth.hash_uniform((5, 4, 1, 16, 16), 22).requires_grad_(True)
    w = hal['encoder.weight'].clone().requires_grad_(True)
    b = hal['encoder.bias'].clone().requires_grad_(True)
    y = oracle.compose(s, d, w, b)
    assert rel(y, g['y']) < 1e-6
    y.backward(synth.hash_uniform(tuple(y.shape), 23))
    assert rel(s.grad, g['grad_static']) < 1e-6 and rel(d.grad, g['grad_dynamic']) < 1e-6
    assert rel(w.grad, g['grad_weight']) < 1e-6 and rel(b.grad, g['grad_bias']) < 1e-6


def test_dm_loss_closed_form_gradient():
    gen = torch.Generator().manual_seed(0)
    er, es = torch.randn(6, 40, generator=gen), torch.randn(3, 40, generator=gen).requires_grad_(True)
    oracle.dm_loss(er, es).backward()
    from oracle.dm import dm_loss_grad
    assert rel(dm_loss_grad(er, es.detach()), es.grad) < 1e-6


def test_dm_s2d_iteration():
    g = np.load(os.path.join(GOLD, 'dm_s2d.npz'))
    C, per, T, H, batch_real, vpc, spc, dpc = 3, 4, 4, 64, 3, 1, 2, 2
    videos = synth.hash_uniform((C * per, T, 3, H, H), 51)
    indices_class = [list(range(c * per, (c + 1) * per)) for c in range(C)]
    static_syn = synth.hash_uniform((C * spc, 3, H, H), 52)
    dyn = synth.hash_uniform((C, dpc, T, 1, H, H), 53)
    hal = synth.synth_hallucinator(5)
    np.random.seed(9)
    params = synth.synth_convnet3d_params(60, num_classes=C)
    r = oracle.dm_s2d_iteration(params, static_syn, dyn, hal, videos, indices_class, vpc=vpc, spc=spc,
                                batch_real=batch_real, coin_dynamic=torch.from_numpy(g['coin_dynamic_v1_0']),
                                coin_static=torch.from_numpy(g['coin_static_v1_0']))
    assert np.array_equal(np.stack(r['real_idx']), g['real_idx_v1_0'])
    assert np.array_equal(r['static_idx'].numpy(), g['static_idx_v1_0'])
    assert rel(r['loss'], g['loss_v1_0']) < 1e-6
    assert rel(r['grad_hal_weight'], g['grad_hal_weight_v1_0']) < 1e-5
    s, samp = synth.summarize(r['grad_dynamic'])
    assert rel(samp, g['grad_dynamic_sample_v1_0']) < 1e-5


def test_sgd_momentum_dense_semantics():
    p, buf = torch.zeros(2), None
    p, buf = oracle.sgd_momentum_step(p, torch.tensor([1.0, 0.0]), buf, 1.0, 0.95)
    p, buf = oracle.sgd_momentum_step(p, torch.tensor([0.0, 0.0]), buf, 1.0, 0.95)
    assert torch.allclose(p, torch.tensor([-1.95, 0.0]))
