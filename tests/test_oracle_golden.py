"""CPU: the oracle against the golden vectors produced from the reference's own code
(tests/golden/make_golden.py).  This is what pins the oracle (see oracle/__init__.py)."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, build_oracle_field, oracle_march, golden_grads, assert_close


def test_hashnerf_oracle_matches_reference_file():
    """HashEmbedderOracle == /root/reference/grids/hash_grid_torch.py (imported verbatim when the golden was made)."""
    from oracle.hashgrid import HashEmbedderOracle
    g = load_golden("hash_torch")
    L, F, T, base, fin = [int(v) for v in g["cfg"]]
    emb = HashEmbedderOracle(L, F, T, base, fin)
    assert np.allclose(emb.resolutions, g["resolutions"])
    with torch.no_grad():
        emb.embeddings.copy_(torch.from_numpy(g["weights"]))
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    idx = emb.indices(x).numpy()
    assert np.array_equal(idx, g["idx"]), "hashed vertex indices must be bit-exact"
    out = emb(x)
    assert_close(out, g["out"], rtol=1e-6, atol_scale=1e-7, msg="features")
    (out * torch.from_numpy(g["gout"])).sum().backward()
    assert_close(emb.embeddings.grad, g["grad_weights"], rtol=1e-5, atol_scale=1e-6, msg="grad table")
    assert_close(x.grad, g["grad_x"], rtol=1e-4, atol_scale=1e-5, msg="grad x")


@pytest.mark.parametrize("name,mode", [("trace_delta_permuto_ray", "ray"), ("trace_delta_permuto_voxel", "voxel"),
                                       ("trace_nef_tcnn_ray", "ray"), ("trace_dd_permuto_ray", "ray")])
def test_trace_oracle_matches_reference_glue(name, mode):
    """oracle.field.trace_oracle == the reference's tracer + nef source run on the oracle-backed stubs."""
    from oracle.field import trace_oracle
    g = load_golden(name)
    field = build_oracle_field(g)
    ridx, pidx, samples, depths, deltas, boundary = oracle_march(g, mode)
    o = torch.from_numpy(g["o"]).requires_grad_(True)
    d = torch.from_numpy(g["d"]).requires_grad_(True)
    t = depths.reshape(samples.shape[0], -1, 1)
    s_attached = o[ridx][:, None] + d[ridx][:, None] * t
    samples = samples + (s_attached - s_attached.detach())
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    out = trace_oracle(field, o, d, ridx, samples, depths, deltas, boundary, chans,
                       bg_color='white' if bool(g["bg_white"]) else 'black', dd=name.startswith("trace_dd"))
    for c in chans + ['alpha']:
        assert_close(out[c], g["out_" + c], rtol=1e-5, atol_scale=1e-6, msg=c)
    assert np.array_equal(out['hit'].numpy(), g["out_hit"])
    loss = sum((out[c] * torch.from_numpy(g["gw_" + c])).sum() for c in chans)
    loss.backward()
    assert_close(o.grad, g["grad_o"], rtol=1e-3, atol_scale=1e-4, msg="grad origins")
    assert_close(d.grad, g["grad_d"], rtol=1e-3, atol_scale=1e-4, msg="grad dirs")
    gg = golden_grads(g)
    named = dict(field.named_parameters())
    remap = {"grid.lattice_values": "grid.embedder.lattice_values", "delta_grid.lattice_values": "delta_grid.embedder.lattice_values",
             "grid.params": "grid.embedder.params"}
    checked = 0
    for k, p in named.items():
        gk = remap.get(k, k)
        if gk in gg:
            assert_close(p.grad if p.grad is not None else torch.zeros_like(p), gg[gk], rtol=1e-3, atol_scale=1e-4, msg="grad " + k)
            checked += 1
    assert checked >= 20
