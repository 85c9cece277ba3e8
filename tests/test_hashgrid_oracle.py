"""CPU: self-checks of the hash-grid restatement (oracle/hashgrid_standin.py; parity unpinned, see its header)
and agreement of the library's host-side level table with it."""
import torch

from loner_b200 import ops
from oracle import hashgrid_standin as H
from oracle import loner_oracle as orc

CONFIGS = [dict(), dict(n_levels=8, log2_hashmap_size=15),
           dict(n_levels=12, base_resolution=8, per_level_scale=1.5, log2_hashmap_size=16)]


def test_default_levels_are_the_published_ones():
    hs = H.HashGridSpec()          # cfg/nerf_config/default_nerf_hash.yaml pos_encoding_sigma
    assert hs.resolutions[:4] == [16, 32, 64, 128] and hs.resolutions[-1] == 16 * 2 ** 15
    assert hs.entries[:3] == [4096, 32768, 262144] and all(e == 1 << 18 for e in hs.entries[2:])
    assert [hs.level_is_dense(l) for l in range(4)] == [True, True, True, False]
    assert hs.n_output_dims == 32


def test_spatial_hash_known_answers():
    hs = H.HashGridSpec()
    p = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [123456, 654321, 111111]])
    want = [0, 1, 2654435761, 805459861, 1 ^ 2654435761 ^ 805459861,
            123456 ^ ((654321 * 2654435761) & 0xFFFFFFFF) ^ ((111111 * 805459861) & 0xFFFFFFFF)]
    got = H.grid_index(hs, 5, p)   # level 5 is hashed, 2^18 entries
    assert got.tolist() == [w % (1 << 18) for w in want]
    dense = H.grid_index(hs, 0, torch.tensor([[3, 2, 1], [16, 16, 16]]))
    assert dense.tolist() == [3 + 2 * 16 + 1 * 256, (16 + 16 * 16 + 16 * 256) % 4096]


def test_constant_table_gives_constant_features():
    """The 8 interpolation weights are a partition of unity."""
    hs = H.HashGridSpec(n_levels=6, log2_hashmap_size=12)
    x = torch.rand(257, 3, generator=torch.Generator().manual_seed(0))
    enc = H.hashgrid_encode(x, torch.full((hs.n_params,), 0.25), hs, precision="fp32")
    assert torch.allclose(enc, torch.full_like(enc, 0.25), atol=1e-6)


def test_table_gradient_is_the_scatter_of_the_weights():
    hs = H.HashGridSpec(n_levels=3, log2_hashmap_size=10, base_resolution=4)
    g = torch.Generator().manual_seed(1)
    table = (torch.rand(hs.n_params, generator=g) - 0.5).requires_grad_(True)
    x = torch.rand(64, 3, generator=g, dtype=torch.float32).requires_grad_(True)
    enc = H.hashgrid_encode(x, table, hs, precision="fp32")
    up = torch.randn(enc.shape, generator=g)
    (enc * up).sum().backward()
    # every sample spreads a total weight of 1 per level and feature
    assert torch.isfinite(table.grad).all() and torch.isfinite(x.grad).all()
    tot = table.grad.view(-1, 2).sum(0)
    want = torch.stack([up[:, 0::2].sum(), up[:, 1::2].sum()])
    assert torch.allclose(tot, want, rtol=1e-4, atol=1e-4)
    # input gradient against a central difference (away from cell boundaries the encoding is multilinear)
    eps = 1e-4
    xd = x.detach().double()
    td = table.detach().double()
    def f(xx):
        return (H.hashgrid_encode(xx.float(), td.float(), hs, precision="fp32").double() * up.double()).sum()
    num = torch.zeros(3, dtype=torch.float64)
    i = 5
    for d in range(3):
        e = torch.zeros_like(xd); e[i, d] = eps
        num[d] = (f(xd + e) - f(xd - e)) / (2 * eps)
    assert torch.allclose(x.grad[i].double(), num, rtol=2e-2, atol=2e-2)


def test_library_level_table_matches_the_oracle():
    """Host-side code of loner_b200/csrc/hashgrid.cu (no GPU needed): parameter count, table entries, padded width."""
    for cfg in CONFIGS:
        hs = H.HashGridSpec(**cfg)
        net = ops.HashNet(n_levels=hs.n_levels, log2_hashmap_size=hs.log2_hashmap_size,
                          base_resolution=hs.base_resolution, per_level_scale=hs.per_level_scale)
        spec = orc.NetSpec(n_neurons=64, n_hidden_layers=1, hash=hs)
        assert net.table_entries == hs.n_entries
        assert net.param_count == spec.n_params
        assert net.e_pad == spec.e_pad
        assert net.layer_shapes() == spec.shapes
        for L in (2, 3, 4):                      # deeper 64-wide heads
            deep = ops.HashNet(n_levels=hs.n_levels, log2_hashmap_size=hs.log2_hashmap_size,
                               base_resolution=hs.base_resolution, per_level_scale=hs.per_level_scale, n_hidden_layers=L)
            dspec = orc.NetSpec(n_neurons=64, n_hidden_layers=L, hash=hs)
            assert deep.param_count == dspec.n_params and deep.layer_shapes() == dspec.shapes
    for bad in (dict(n_neurons=128), dict(n_hidden_layers=5), dict(n_hidden_layers=2, flags=ops.HASH_SCALAR),
                dict(n_features_per_level=4), dict(n_levels=17)):
        try:
            ops.HashNet(**bad)
        except RuntimeError:
            continue
        raise AssertionError(f"{bad} should be refused")
