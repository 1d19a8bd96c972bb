"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol the header
declares, host-only entry points behave, and the nn.Module mirrors keep the reference's
state_dict contract (key order, shapes, aliasing, initialisation).  No compute calls."""
import os
import re

import numpy as np
import pytest
import torch

from bmcnet_esr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'bmc_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(bmc_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 24
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    l = _lib.lib()                       # raises if any symbol is missing from the .so
    assert l.bmc_abi_version() == 1
    assert l.bmc_act_dtype() in (b'f16', b'bf16')


def test_model_create_rejects_unsupported_hyperparameters():
    l = _lib.lib()
    assert l.bmc_model_create(0, 2, 128, 5, 3) is None
    assert b'scale=4' in l.bmc_last_error()
    assert l.bmc_model_create(7, 4, 128, 5, 3) is None
    h = l.bmc_model_create(_lib.MODEL_BMCNET, 4, 128, 5, 3)
    assert h
    assert l.bmc_model_weight_bytes(h) > 2 * 2731680 * 0.9     # ~ one 16-bit copy of the unique weights
    assert l.bmc_model_configure(h, 0, 45, 80) == -1           # BMC_ERR_ARG
    assert l.bmc_model_configure(h, 2, 45, 80) == 0
    assert l.bmc_model_workspace_bytes(h) > 0
    assert l.bmc_model_bind_workspace(h, None, 0) == -4        # BMC_ERR_STATE: weights not loaded
    l.bmc_model_destroy(h)


def test_compute_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    from bmcnet_esr_b200.dataloader import encodings as G
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    x = torch.zeros(8)
    with pytest.raises(_lib.BmcError):
        G.events_to_channels(x, x.clone(), x.clone(), sensor_size=(4, 4))
    m = BMCNet_plain(4, 128, 5).eval()
    with pytest.raises(_lib.BmcError):
        m(torch.zeros(1, 2, 2, 8, 8), torch.zeros(1, 128, 8, 8), torch.zeros(1, 32, 8, 8), True)


@pytest.mark.parametrize('plain', [False, True])
def test_module_state_dict_matches_reference_contract(golden_dir, plain):
    from bmcnet_esr_b200.models.BMCNet import BMCNet
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    g = np.load(os.path.join(golden_dir, 'statedict_%s.npz' % ('plain' if plain else 'bmcnet')))
    m = (BMCNet_plain if plain else BMCNet)(4, 128, 5)
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g['keys']]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g['shapes']]
    ptr = {}
    assert [ptr.setdefault(v.data_ptr(), len(ptr)) for v in sd.values()] == [int(v) for v in g['alias_group']]
    assert sum(p.numel() for p in m.parameters()) == int(g['n_unique_params'])


def test_shipped_checkpoint_loads_strictly(plain_ckpt):
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    m = BMCNet_plain(4, 128, 5)
    m.load_state_dict(plain_ckpt, strict=True)
    assert torch.equal(m.neuro.conv_f2.weight, plain_ckpt['neuro.conv_f1.weight'])


def test_initialisation_follows_reference_recipe():
    # kaiming_normal(fan_in) * 0.1 and zero bias for the listed convs; conv_fs of the plain
    # model keeps the default init (reference BMCNet_plain.py:17)
    from bmcnet_esr_b200.models.BMCNet_plain import BMCNet_plain
    torch.manual_seed(0)
    m = BMCNet_plain(4, 128, 5)
    w = m.neuro.conv_h.weight
    assert abs(w.std().item() - 0.1 * (2.0 / (128 * 9)) ** 0.5) < 2e-4
    assert float(m.neuro.conv_h.bias.abs().max()) == 0.0
    assert float(m.neuro.conv_fs.bias.abs().max()) > 0.0


def test_training_launch_structure_threshold_and_bench_targets():
    """Host logic of the training path that needs no GPU: stacked multi-job launches are used while one job fills at
    most half a wave of SMs; the bench's synthetic targets have the prediction's shape and are seeded per rank."""
    import torch
    from bmcnet_esr_b200.models import _train as TR
    import bench
    assert TR._use_stack(2, 45, 80) and TR._use_stack(4, 45, 80) and not TR._use_stack(8, 45, 80)
    assert TR._use_stack(16, 22, 40)
    xs = [torch.rand(2, 2, 2, 6, 9) for _ in range(3)]
    a, b2, c = bench.train_targets(xs, 0), bench.train_targets(xs, 0), bench.train_targets(xs, 1)
    assert len(a) == 3 and a[0].shape == (2, 2, 24, 36)
    assert all(torch.equal(u, v) for u, v in zip(a, b2)) and not torch.equal(a[0], c[0])
