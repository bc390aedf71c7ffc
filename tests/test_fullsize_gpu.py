"""Size-independent properties at the FULL size of BASELINE.json configs[1] (50 classes, 16x3x112x112, batch_real 64,
3200 real + 50 synthetic videos per iteration), where the CPU oracle is far too slow to serve as the checker:

* determinism     — the same seeds give bitwise identical loss, embeddings and class means (the two MMA issuer threads
                    enter the tensor pipe in a fixed order); the memory gradient is identical up to the last bit of a few
                    hundred of its 80 M elements (see the test);
* chunk invariance — an embedding does not depend on the launch it was computed in (chunk size, position in the batch);
* class additivity — the DM loss and the dynamic-memory gradient of all 50 classes equal the sum over class-sharded
                    sub-problems (what the multi-GPU path relies on), and gradient rows of unselected memories are exactly 0;
* pack round trip — the resident packed bf16 operand reproduces bf16(video) exactly.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

C, T, HW, PER, BATCH_REAL = 50, 16, 112, 66, 64


@pytest.fixture(scope='module')
def world():
    from video_distillation_b200.distill import DeviceDataset
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev).manual_seed(11)
    vids = torch.empty(C * PER, T, 3, HW, HW, device=dev).normal_(generator=g)
    labels = [c for c in range(C) for _ in range(PER)]
    ds = DeviceDataset.from_device_shard(vids, labels, C, dev, 0, 1)
    yield ds
    del ds, vids
    torch.cuda.empty_cache()


def make_trainer(ds, seed=5, **kw):
    from video_distillation_b200.distill import DMS2DTrainer
    torch.manual_seed(seed)
    return DMS2DTrainer(ds, num_classes=C, im_size=(HW, HW), frames=T, vpc=1, spc=2, dpc=2, batch_real=BATCH_REAL,
                        lr_dynamic=1e4, lr_hal=1e-2, precision='bf16', device='cuda', init_on_device=True, max_batch=640, **kw)


def run_step(tr, seed=3):
    np.random.seed(seed)
    torch.cuda.manual_seed(100 + seed)
    loss = tr.step(net_seed=77)
    torch.cuda.synchronize()
    return loss.clone(), tr.dynamic_syn.grad.clone(), tr.last['emb_syn'].clone(), tr.last['mean_real'].clone()


def test_full_size_iteration_is_deterministic(world):
    ds = world
    outs = []
    for _ in range(2):
        tr = make_trainer(ds)
        if ds.x0 is None:
            ds.prepack(tr.embedder.tc, extra_slots=C)
        outs.append(run_step(tr))
    for name, a, b in zip(('loss', 'grad_dynamic', 'emb_syn', 'mean_real'), outs[0], outs[1]):
        if name == 'grad_dynamic':
            # The forward is bitwise reproducible.  The direct conv-1 dgrad is reproducible up to the LAST BIT of ~0.1 % of its
            # outputs: the baton between the two MMA issuer threads orders the issue of their MMAs, not their retirement into
            # the shared accumulator, and the tensor core's truncating accumulate makes the last bit depend on it (DESIGN.md
            # section 5, scripts/dgrad1_probe.py).  Gate: at most 1e-4 of the elements differ, each by < 1e-6 of the largest.
            diff = (a - b).abs()
            assert int((diff > 0).sum()) <= 1e-4 * a.numel(), (name, int((diff > 0).sum()))
            assert diff.max().item() <= 1e-6 * a.abs().max().item(), (name, diff.max().item(), a.abs().max().item())
            continue
        assert torch.equal(a, b), (name, (a - b).abs().max().item(), int((a != b).sum()))
    loss, g_dyn, emb_syn, mean_real = outs[0]
    assert torch.isfinite(loss) and loss.item() > 0
    # vpc = 1, dpc = 2: exactly one of the two dynamic memories of every class was selected (distill_s2d_ms.py:405)
    nz = (g_dyn.flatten(2).abs().sum(-1) > 0)
    assert nz.sum(1).eq(1).all()


def test_embedding_is_independent_of_chunking_and_position(world):
    ds = world
    tr = make_trainer(ds)
    if ds.x0 is None:
        ds.prepack(tr.embedder.tc, extra_slots=C)
    from video_distillation_b200.distill import frozen_convnet3d
    net = frozen_convnet3d(3, C, (HW, HW), T, 'cuda', seed=4, init_on_device=True)
    tr.embedder.load(net)
    tc = tr.embedder.tc
    idx = torch.randperm(C * PER, device='cuda')[:1500]
    e_all = tc.embed_resident(ds.x0, idx).clone()
    e_rev = tc.embed_resident(ds.x0, idx.flip(0)).flip(0)
    assert torch.equal(e_all, e_rev)
    e_parts = torch.cat([tc.embed_resident(ds.x0, idx[s:s + 333]).clone() for s in range(0, 1500, 333)], 0)
    assert torch.equal(e_all, e_parts)
    # the packed operand is bf16(video): embedding the fp32 videos directly (per-step packing) gives the same bits
    e_direct = tc.embed(ds.videos, index=idx[:200])
    assert torch.equal(e_all[:200], e_direct)


def test_loss_and_gradient_are_additive_over_class_shards(world):
    from video_distillation_b200 import distill
    ds = world
    tr = make_trainer(ds)
    if ds.x0 is None:
        ds.prepack(tr.embedder.tc, extra_slots=C)
    static0, dyn0 = tr.static_syn.detach().clone(), tr.dynamic_syn.detach().clone()
    hal_state = {k: v.clone() for k, v in tr.hal.state_dict().items()}
    loss_all, g_all, _, _ = run_step(tr)
    g_hal_all = tr.hal.encoder.weight.grad.clone()
    loss_sum, g_sum, g_hal_sum = 0.0, torch.zeros_like(g_all), torch.zeros_like(g_hal_all)
    for r in range(3):                                      # emulate a 3-rank class sharding on one GPU
        tr_r = make_trainer(ds, static_syn=static0, dynamic_syn=dyn0)
        tr_r.hal.load_state_dict(hal_state)
        tr_r.owned = distill.owned_classes(C, r, 3)
        tr_r.owned_t = torch.as_tensor(tr_r.owned, dtype=torch.long, device='cuda')
        tr_r.ds = _ShardView(ds, tr_r.owned)
        l, g, _, _ = run_step(tr_r)
        loss_sum += l.item()
        g_sum += g
        g_hal_sum += tr_r.hal.encoder.weight.grad
    assert abs(loss_sum - loss_all.item()) <= 1e-5 * abs(loss_all.item())
    assert torch.equal(g_sum != 0, g_all != 0)
    assert ((g_sum - g_all).norm() / g_all.norm()).item() < 1e-5          # disjoint rows: only the launch grouping differs
    assert ((g_hal_sum - g_hal_all).norm() / g_hal_all.norm()).item() < 1e-4


class _ShardView:
    """The full resident dataset seen by a trainer that owns a subset of the classes (single process)."""

    def __init__(self, ds, owned):
        self.__dict__.update(ds.__dict__)
        self._ds = ds
        self.x0 = None                                       # separate launches per shard: no shared spare slots

    def sample_all_classes(self, n):
        return self._ds.sample_all_classes(n)

    def local_index(self, idx):
        return self._ds.local_index(idx)
