"""GPU parity: every CUDA kernel (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: bit-exact for integers (octree cells, packed indices, lattice / hash vertex indices, boundaries) and for the
marcher's float32 depths/samples (identical op order); 1e-4 relative for features, rendered outputs and
gradients (north_star).  Honest statement: "bit-exact vs OUR CPU restatement of the upstream algorithms; upstream
binaries unavailable; the in-tree hash_grid_torch and the tracer/nef glue are matched via the reference-made goldens".
"""
import numpy as np
import pytest
import torch

from tests.util import (load_golden, build_oracle_field, build_cuda_nef, oracle_march, golden_grads, golden_params,
                        assert_close, assert_close_norm, GOLDEN_CFG)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scene(level, seed=0, frac=0.25):
    rng = np.random.default_rng(seed)
    n = 1 << level
    k = max(8, int(frac * n ** 3))
    return rng.integers(0, n, size=(k, 3)).astype(np.int16)


def _rays(N, seed=0, inside=False):
    rng = np.random.default_rng(seed)
    if inside:
        o = rng.uniform(-0.7, 0.7, (N, 3)).astype(np.float32)
    else:
        o = np.stack([rng.uniform(-0.6, 0.6, N), rng.uniform(-0.6, 0.6, N), np.full(N, 1.3)], 1).astype(np.float32)
    d = rng.normal(size=(N, 3)).astype(np.float32)
    if not inside:
        d[:, 2] = -np.abs(d[:, 2]) - 0.3
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[0] = np.array([0.0, 0.0, -1.0], np.float32)      # axis-aligned: zero components -> inf reciprocals
    d[1] = np.array([0.0, 1.0, 0.0], np.float32)
    return o, d.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# octree
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("level", [1, 4, 6])
def test_octree_build_and_query(cuda_lib, level):
    from oracle import spc as ospc
    from pagnerf_b200 import spc, ops
    pts = _scene(level, seed=level)
    oc_ref = ospc.points_to_octree(pts, level)
    p_ref, py_ref, pre_ref = ospc.scan_octree(oc_ref, level)
    oc = spc.unbatched_points_to_octree(torch.from_numpy(pts).to(DEV), level)
    assert np.array_equal(oc.cpu().numpy(), oc_ref)
    blas = spc.OctreeAS(DEV)
    blas.init(oc)
    assert blas.max_level == level
    assert np.array_equal(blas.points.cpu().numpy(), p_ref)
    assert np.array_equal(blas.pyramid.numpy(), py_ref)
    assert np.array_equal(blas.prefix.cpu().numpy(), pre_ref)
    rng = np.random.default_rng(1)
    x = rng.uniform(-1.05, 1.05, (20000, 3)).astype(np.float32)
    x[:8] = np.array([[1, 1, 1], [-1, -1, -1], [0, 0, 0], [1, 0, 0], [np.nan, 0, 0], [np.inf, 0, 0], [0.999999, 0.5, -1], [-1.0000001, 0, 0]], np.float32)
    got = ops.octree_query(blas.octree, blas.prefix, torch.from_numpy(x).to(DEV), level).cpu().numpy()
    assert np.array_equal(got, ospc.query(oc_ref, pre_ref, x, level))


def test_octree_query_empty_and_dense(cuda_lib):
    from oracle import spc as ospc
    from pagnerf_b200 import spc, ops
    blas = spc.OctreeAS(DEV)
    blas.init_dense(3)
    assert np.array_equal(blas.octree.cpu().numpy(), ospc.dense_octree(3))
    out = ops.octree_query(blas.octree, blas.prefix, torch.zeros(0, 3, device=DEV), 3)
    assert out.shape == (0,)


@pytest.mark.parametrize("level,inside", [(4, False), (6, False), (5, True)])
def test_raytrace_nuggets_bit_exact(cuda_lib, level, inside):
    from oracle import spc as ospc
    from pagnerf_b200 import spc, ops
    pts = _scene(level, seed=10 + level, frac=0.15)
    oc = ospc.points_to_octree(pts, level)
    p, py, pre = ospc.scan_octree(oc, level)
    o, d = _rays(700, seed=level, inside=inside)
    r_ref, p_ref, dep_ref = ospc.raytrace(oc, p, py, pre, o, d, level)
    blas = spc.OctreeAS(DEV); blas.init(torch.from_numpy(oc))
    ridx, pidx, depth, offsets = ops.raytrace(blas.octree, blas.prefix, torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), level)
    assert np.array_equal(ridx.cpu().numpy(), r_ref)
    assert np.array_equal(pidx.cpu().numpy(), p_ref)
    assert np.array_equal(depth.cpu().numpy().view(np.uint32), dep_ref.view(np.uint32)), "entry/exit depths bit-exact"
    assert int(offsets[-1]) == r_ref.shape[0]


@pytest.mark.parametrize("S", [7, 32, 64, 100])
def test_raymarch_ray_bit_exact(cuda_lib, S):
    from oracle import spc as ospc, raymarch as orm
    from pagnerf_b200 import spc, ops
    level = 5
    oc = ospc.points_to_octree(_scene(level, seed=3, frac=0.2), level)
    p, py, pre = ospc.scan_octree(oc, level)
    o, d = _rays(300, seed=S)
    ref = orm.raymarch_ray(oc, pre, o, d, level, S, 0.0, 2.5, seed=5)
    blas = spc.OctreeAS(DEV); blas.init(torch.from_numpy(oc))
    got = ops.raymarch_ray(blas.octree, blas.prefix, torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), level, S, 0.0, 2.5, seed=5)
    names = ["ridx", "pidx", "samples", "depths", "deltas", "boundary"]
    for n, a, b in zip(names, got[:6], ref):
        a = a.cpu().numpy()
        assert a.shape == b.shape, (n, a.shape, b.shape)
        if a.dtype == np.float32:
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), n
        else:
            assert np.array_equal(a, b), n
    # explicit jitter tensor == counter stream
    from oracle.f32 import jitter_u01
    jit = jitter_u01(5, np.arange(300 * S, dtype=np.uint64)).reshape(300, S)
    got2 = ops.raymarch_ray(blas.octree, blas.prefix, torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), level, S, 0.0, 2.5,
                            jitter=torch.from_numpy(jit).to(DEV))
    assert torch.equal(got2[2], got[2]) and torch.equal(got2[0], got[0])
    # the occupancy-bit-field marcher behind raymarch(need_pidx=False): same packed samples, bit for bit, no point indices
    got3 = ops.raymarch_ray_bits(blas.level_bits(level), torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), level, S, 0.0, 2.5, seed=5)
    assert got3[1] is None
    for i in (0, 2, 3, 4, 5, 6):
        assert torch.equal(got3[i], got[i]), names[i] if i < 6 else "offsets"


def test_raymarch_voxel_bit_exact_and_filter(cuda_lib):
    from oracle import spc as ospc, raymarch as orm
    from pagnerf_b200 import spc, ops
    level, S = 5, 3
    oc = ospc.points_to_octree(_scene(level, seed=4, frac=0.2), level)
    p, py, pre = ospc.scan_octree(oc, level)
    o, d = _rays(400, seed=2)
    ref = orm.raymarch_voxel(oc, p, py, pre, o, d, level, S, seed=9)
    blas = spc.OctreeAS(DEV); blas.init(torch.from_numpy(oc))
    got = ops.raymarch_voxel(blas.octree, blas.prefix, torch.from_numpy(o).to(DEV), torch.from_numpy(d).to(DEV), level, S, seed=9)
    for n, a, b in zip(["ridx", "pidx", "samples", "depths", "deltas", "boundary"], got[:6], ref):
        a = a.cpu().numpy()
        assert a.shape == b.shape, (n, a.shape, b.shape)
        if a.dtype == np.float32:
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), n
        else:
            assert np.array_equal(a, b), n
    keep_ref = orm.max_travel_filter(ref[0], ref[3], 0.4)
    first = ops.ray_offsets(got[0], 400)
    keep = ops.max_travel_mask(got[0], got[3], first, 0.4).cpu().numpy()
    assert np.array_equal(keep, keep_ref) and 0 < keep.sum() < keep.shape[0]


def test_raymarch_no_hits(cuda_lib):
    from pagnerf_b200 import spc, ops
    blas = spc.OctreeAS(DEV); blas.init_dense(2)
    o = torch.tensor([[3.0, 3.0, 3.0]] * 5, device=DEV)
    d = torch.tensor([[0.0, 0.0, 1.0]] * 5, device=DEV)
    for out in (ops.raymarch_ray(blas.octree, blas.prefix, o, d, 2, 16, 0.0, 1.0), ops.raymarch_voxel(blas.octree, blas.prefix, o, d, 2, 4)):
        assert out[0].numel() == 0 and out[2].shape[0] == 0 and int(out[6][-1]) == 0


# ------------------------------------------------------------------------------------------------
# encoders
# ------------------------------------------------------------------------------------------------
def _points(M, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (M, 3)).astype(np.float32)
    x[:6] = np.array([[0, 0, 0], [1, 1, 1], [-1, -1, -1], [0.5, -0.25, 0.125], [1e-6, -1e-6, 0], [0.999, -0.999, 0.3]], np.float32)
    return x


@pytest.mark.parametrize("cap,L,finest", [(2 ** 12, 8, 1e-2), (2 ** 18, 24, 1e-4), (1000, 5, 1e-1)])
def test_permuto_indices_and_forward(cuda_lib, cap, L, finest):
    from oracle.permuto import PermutoEncodingOracle
    from pagnerf_b200 import ops
    enc = PermutoEncodingOracle(cap, L, 2, np.geomspace(1.0, finest, L), seed=1)
    with torch.no_grad():
        enc.lattice_values.mul_(1e4)
    x = torch.from_numpy(_points(3000, seed=L))
    rem0, rank, idx = enc.indices(x)
    sf, sh = enc.scale_factor.to(DEV), enc.random_shift_per_level.to(DEV)
    gi, gr, gb = ops.permuto_indices(x.to(DEV), cap, sf, sh)
    assert np.array_equal(gi.cpu().numpy().view(np.uint32), idx), "lattice vertex hash indices bit-exact"
    assert np.array_equal(gr.cpu().numpy(), rank), "simplex ranks bit-exact"
    ref = enc(x)
    out = ops.permuto_encode(x.to(DEV), enc.lattice_values.detach().to(DEV), sf, sh, enc.anneal_window.to(DEV))
    assert_close(out, ref.detach(), msg="permuto features")


@pytest.mark.parametrize("n_agg", [0, 3, 8])
def test_permuto_backward(cuda_lib, n_agg):
    from oracle.permuto import PermutoEncodingOracle
    from pagnerf_b200 import ops
    cap, L = 2 ** 10, 8
    enc = PermutoEncodingOracle(cap, L, 2, np.geomspace(1.0, 1e-2, L), seed=2)
    with torch.no_grad():
        enc.lattice_values.mul_(1e4)
    x = torch.from_numpy(_points(2053, seed=3)).requires_grad_(True)   # not a multiple of 32: ragged last warp
    g = torch.randn(2053, 2 * L, generator=torch.Generator().manual_seed(0))
    (enc(x) * g).sum().backward()
    xt = x.detach().to(DEV).requires_grad_(True)
    tb = enc.lattice_values.detach().to(DEV).requires_grad_(True)
    out = ops.permuto_encode(xt, tb, enc.scale_factor.to(DEV), enc.random_shift_per_level.to(DEV), enc.anneal_window.to(DEV), n_agg)
    (out * g.to(DEV)).sum().backward()
    assert_close(tb.grad, enc.lattice_values.grad, msg="grad lattice_values")
    assert_close(xt.grad, x.grad, msg="grad positions")


def test_permuto_empty(cuda_lib):
    from pagnerf_b200.grids import PermutoGrid
    g = PermutoGrid(2, blas_level=2, num_lods=4, capacity_log_2=8)
    g.init_from_scales()
    g = g.to(DEV)
    assert g.interpolate(torch.zeros(0, 1, 3, device=DEV)).shape == (0, 1, 8)


def test_hashnerf_matches_reference_golden(cuda_lib):
    """CUDA flavour-1 hash grid vs the reference's own grids/hash_grid_torch.py outputs (golden)."""
    from pagnerf_b200.grids.hash_grid_torch import HashEmbedder
    from pagnerf_b200 import ops
    g = load_golden("hash_torch")
    L, F, T, base, fin = [int(v) for v in g["cfg"]]
    emb = HashEmbedder(L, F, T, base, fin)
    assert np.allclose(emb.level_res.numpy(), g["resolutions"])
    emb.load_state_dict({f"embeddings.{l}.weight": torch.from_numpy(g["weights"][l]) for l in range(L)})
    assert sorted(emb.state_dict().keys()) == sorted(f"embeddings.{l}.weight" for l in range(L))
    emb = emb.to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    idx = ops.hash_indices(x, 1, emb.level_res, None, emb.level_offset, emb.level_size).cpu().numpy()
    perm = [((k & 1) << 2) | (((k >> 1) & 1) << 1) | ((k >> 2) & 1) for k in range(8)]  # ours (bit d = dim d) -> reference (i,j,k x-major)
    assert np.array_equal(idx.astype(np.int64), g["idx"][:, :, perm]), "hashed vertex indices bit-exact vs reference"
    out = emb(x)
    assert_close(out, g["out"], msg="features")
    (out * torch.from_numpy(g["gout"]).to(DEV)).sum().backward()
    assert_close(emb.embeddings_weight.grad, g["grad_weights"], msg="grad table")
    assert_close(x.grad, g["grad_x"], rtol=1e-3, msg="grad x")


@pytest.mark.parametrize("L,log2T,base", [(6, 10, 16), (5, 14, 4), (14, 19, 16)])
def test_tcnn_hash_indices_forward_backward(cuda_lib, L, log2T, base):
    from oracle.hashgrid import TcnnHashGridOracle
    from pagnerf_b200.grids.hash_grid_tinycudann import TcnnEncoding
    from pagnerf_b200 import ops
    ref = TcnnHashGridOracle(L, 2, log2T, base, 2.0, seed=4)
    with torch.no_grad():
        ref.params.mul_(1e3)
    enc = TcnnEncoding(3, {"otype": "HashGrid", "n_levels": L, "n_features_per_level": 2, "log2_hashmap_size": log2T,
                           "base_resolution": base, "per_level_scale": 2})
    assert enc.params.shape == ref.params.shape
    assert np.array_equal(enc.level_size.numpy(), ref.sizes.astype(np.int64))
    enc.load_state_dict({"params": ref.params.detach()})
    enc.round_half = False
    enc = enc.to(DEV)
    x = torch.from_numpy(_points(1500, seed=L)).requires_grad_(True)
    idx_ref = np.stack(ref.indices(x))
    xt = x.detach().to(DEV).requires_grad_(True)
    idx = ops.hash_indices(xt, 0, enc.level_scale, enc.level_res, enc.level_offset, enc.level_size).cpu().numpy()
    assert np.array_equal(idx.view(np.uint32), idx_ref), "tcnn grid indices bit-exact (incl. dense levels, negative cells)"
    g = torch.randn(1500, 2 * L, generator=torch.Generator().manual_seed(1))
    (ref(x) * g).sum().backward()
    out = enc(xt)
    assert_close(out, ref(x).detach(), msg="features")
    (out * g.to(DEV)).sum().backward()
    assert_close(enc.params.grad, ref.params.grad, msg="grad params")
    assert_close(xt.grad, x.grad, rtol=1e-3, msg="grad x")
    enc.round_half = True
    ref.out_half = True
    assert_close(enc(xt), ref(x).detach(), rtol=2e-3, atol_scale=2e-3, msg="fp16-rounded features (tol 2e-3)")


# ------------------------------------------------------------------------------------------------
# decoders (through the nef plugin) and compositing (through the tracer plugin)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["trace_delta_permuto_ray", "trace_nef_tcnn_ray"])
def test_nef_forward_backward_vs_oracle(cuda_lib, name):
    g = load_golden(name)
    field = build_oracle_field(g)
    nef = build_cuda_nef(g, DEV)
    if "tcnn" in name:
        nef.grid.embedder.round_half = True
    gen = torch.Generator().manual_seed(0)
    M, S = 301, 2
    coords = (torch.rand(M, S, 3, generator=gen) * 2 - 1).requires_grad_(True)
    ray_d = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen), dim=-1).requires_grad_(True)
    chans = {'density', 'rgb', 'semantics', 'inst_embedding'}
    ref = field(coords, ray_d, chans)
    ct, dt = coords.detach().to(DEV).requires_grad_(True), ray_d.detach().to(DEV).requires_grad_(True)
    out = nef(coords=ct, ray_d=dt, channels=chans)
    tol = dict(rtol=2e-3, atol_scale=2e-3) if "tcnn" in name else {}
    gws = {}
    for c in sorted(chans):
        assert out[c].shape == ref[c].shape, c
        assert_close(out[c], ref[c].detach(), msg=c, **tol)
        gws[c] = torch.randn(ref[c].shape, generator=gen)
    sum((ref[c] * gws[c]).sum() for c in chans).backward()
    sum((out[c] * gws[c].to(DEV)).sum() for c in chans).backward()
    if "tcnn" in name:
        return  # fp16-rounded features: gradients are checked in the permuto case and in the tcnn encoder test
    named = dict(nef.named_parameters())
    remap = {"grid.lattice_values": "grid.embedder.lattice_values", "delta_grid.lattice_values": "delta_grid.embedder.lattice_values"}
    for k, p in field.named_parameters():
        assert_close(named[remap.get(k, k)].grad, p.grad, rtol=1e-3, msg="grad " + k)
    assert_close(ct.grad, coords.grad, rtol=1e-3, msg="grad coords")
    assert_close(dt.grad, ray_d.grad, rtol=1e-3, msg="grad ray_d")


def test_nef_channel_dispatch(cuda_lib):
    g = load_golden("trace_delta_permuto_ray")
    nef = build_cuda_nef(g, DEV)
    coords = torch.rand(10, 1, 3, device=DEV) * 2 - 1
    rd = torch.nn.functional.normalize(torch.randn(10, 3, device=DEV), dim=-1)
    d = nef(coords=coords, ray_d=rd, channels="density")
    assert torch.is_tensor(d) and d.shape == (10, 1, 1)
    lst = nef(coords=coords, ray_d=rd, channels=["rgb", "density"])
    assert isinstance(lst, list) and lst[0].shape == (10, 1, 3)
    dct = nef(coords=coords, ray_d=rd, channels={"semantics"})
    assert set(dct) == {"semantics"} and dct["semantics"].shape == (10, 6)
    assert torch.allclose(dct["semantics"].sum(-1), torch.ones(10, device=DEV), atol=1e-5)


@pytest.mark.parametrize("name,mode", [("trace_delta_permuto_ray", "ray"), ("trace_delta_permuto_voxel", "voxel"),
                                       ("trace_nef_tcnn_ray", "ray"), ("trace_dd_permuto_ray", "ray")])
def test_trace_matches_reference_golden(cuda_lib, name, mode):
    """Full CUDA path (march -> encode -> decode -> composite, fwd + bwd) through the tracer plugin vs the outputs
    of the REFERENCE's tracer/nef source (golden).  Marcher integers are compared against the oracle marcher."""
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden(name)
    nef = build_cuda_nef(g, DEV)
    if name.startswith("trace_dd"):      # PanopticDDensityNeF + its tracer (own panoptic density stream)
        from pagnerf_b200.tracers import PanopticDDensityPackedRFTracer as PanopticPackedRFTracer
    tracer = PanopticPackedRFTracer(raymarch_type=mode, num_steps=int(g["num_steps"]),
                                    bg_color='white' if bool(g["bg_white"]) else 'black',
                                    ray_max_travel=float(g["ray_max_travel"]))
    o = torch.from_numpy(g["o"]).to(DEV).requires_grad_(True)
    d = torch.from_numpy(g["d"]).to(DEV).requires_grad_(True)
    rays = Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0)
    # marcher integers vs oracle
    r_ref = oracle_march(g, mode)
    got = nef.grid.raymarch(rays, level=0, num_samples=int(g["num_steps"]), raymarch_type=mode)
    if mode == 'ray':
        assert torch.equal(got[0].cpu(), r_ref[0]) and torch.equal(got[1].cpu(), r_ref[1])
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    rb = tracer(nef, channels=chans, rays=rays, lod_idx=None, stage='train')
    tol = dict(rtol=2e-3, atol_scale=2e-3) if "tcnn" in name else {}
    for c in chans + ['alpha']:
        assert_close(getattr(rb, c), g["out_" + c], msg=c, **tol)
    assert np.array_equal(rb.hit.cpu().numpy(), g["out_hit"])
    loss = sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c]).to(DEV)).sum() for c in chans)
    loss.backward()
    gtol = dict(rtol=5e-3, atol_scale=5e-3) if "tcnn" in name else dict(rtol=1e-3, atol_scale=2e-4)
    gg = golden_grads(g)
    for k, p in nef.named_parameters():
        if k in gg:
            assert_close(p.grad if p.grad is not None else torch.zeros_like(p), gg[k], msg="grad " + k, **gtol)
    assert_close(o.grad, g["grad_o"], msg="grad origins", **gtol)
    assert_close(d.grad, g["grad_d"], msg="grad dirs", **gtol)


def test_composite_empty_rays_and_properties(cuda_lib):
    from pagnerf_b200 import ops
    N = 9
    ridx = torch.tensor([1, 1, 1, 4, 7, 7], device=DEV)
    off = ops.ray_offsets(ridx, N)
    assert off.tolist() == [0, 0, 3, 3, 3, 4, 4, 4, 6, 6]
    sigma = torch.tensor([0.5, 2.0, 30.0, 1.0, 0.0, 0.0], device=DEV)
    deltas = torch.full((6,), 0.1, device=DEV)
    rgb = torch.rand(6, 3, device=DEV)
    alpha, hit, rgb_o, dep, sem, inst, w = ops.composite(sigma, deltas, None, rgb, None, None, off, True)
    a = alpha[:, 0].cpu()
    assert torch.all(a >= 0) and torch.all(a <= 1 + 1e-6)
    empty = [0, 2, 3, 5, 6, 8]
    assert torch.all(a[empty] == 0) and not hit[empty].any() and torch.all(rgb_o[empty] == 1.0)
    assert hit[1] and hit[4] and not hit[7]        # zero density ray: alpha == 0 -> no hit, white
    assert torch.allclose(rgb_o[7], torch.ones(3, device=DEV))


def test_kaolin_compat_ops(cuda_lib):
    from oracle import spc as ospc
    from pagnerf_b200 import ops
    gen = torch.Generator().manual_seed(0)
    ridx = torch.sort(torch.randint(0, 40, (500,), generator=gen))[0]
    b_ref = ospc.mark_pack_boundaries(ridx)
    b = ops.mark_pack_boundaries(ridx.to(DEV))
    assert torch.equal(b.cpu(), b_ref)
    tau = torch.rand(500, 1, generator=gen).requires_grad_(True)
    x = torch.randn(500, 5, generator=gen).requires_grad_(True)
    _, w_ref = ospc.exponential_integration(None, tau, b_ref)
    s_ref = ospc.sum_reduce(x * w_ref, b_ref)
    gs = torch.randn(s_ref.shape, generator=gen)
    (s_ref * gs).sum().backward()
    tt, xt = tau.detach().to(DEV).requires_grad_(True), x.detach().to(DEV).requires_grad_(True)
    _, w = ops.exponential_integration(None, tt, b)
    s = ops.sum_reduce(xt * w, b)
    assert_close(w, w_ref.detach(), msg="weights")
    assert_close(s, s_ref.detach(), msg="sum_reduce")
    (s * gs.to(DEV)).sum().backward()
    assert_close(tt.grad, tau.grad, rtol=1e-3, msg="grad tau")
    assert_close(xt.grad, x.grad, msg="grad x")


def test_full_size_properties(cuda_lib):
    """BASELINE-size run (16 384 rays x 128 steps, L=24, T=2^18 x2 grids): size-independent properties."""
    import bench
    step = bench.build_workload(torch.device(DEV), n_rays=16384, seed=0)
    out = step.forward_backward()
    rb = out["rb"]
    a = rb.alpha[:, 0]
    assert torch.isfinite(a).all() and (a >= 0).all() and (a <= 1 + 1e-5).all()
    assert torch.isfinite(rb.rgb).all() and (rb.rgb >= -1e-5).all() and (rb.rgb <= 1 + 1e-5).all()
    s = rb.semantics.sum(-1)
    assert torch.allclose(s, a * a, atol=1e-4), "sum_c of composited softmax probabilities == alpha_p * sum w == alpha^2"
    assert torch.allclose(rb.inst_embedding.sum(-1), a * a, atol=1e-4)
    assert (rb.rgb[~rb.hit] == 1.0).all()
    for n, p in step.nef.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
    # packed indices: sorted rays, boundary count == rays with samples
    assert (out["ridx"][1:] >= out["ridx"][:-1]).all()


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) decoders vs the exact fp32 kernels / oracle
# ------------------------------------------------------------------------------------------------
def test_tcgen05_building_blocks(cuda_lib):
    """K-major x K-major, K-major x MN-major and MN-major x MN-major (K = 128 samples) operand images."""
    from pagnerf_b200._lib import call, ptr
    gen = torch.Generator(device=DEV).manual_seed(0)
    cases = [(0, 64, 48, 0), (0, 16, 64, 0), (0, 208, 64, 0), (1, 64, 64, 0), (1, 48, 64, 0), (1, 64, 208, 0), (1, 64, 16, 0),
             (2, 64, 0, 64), (2, 48, 0, 64), (2, 64, 0, 16), (2, 64, 0, 128)]
    for mode, N, K, FA in cases:
        if mode == 0:
            A, B = torch.randn(128, K, device=DEV, generator=gen), torch.randn(N, K, device=DEV, generator=gen); ref = A @ B.t()
        elif mode == 1:
            A, B = torch.randn(128, K, device=DEV, generator=gen), torch.randn(K, N, device=DEV, generator=gen); ref = A @ B
        else:
            A, B = torch.randn(128, FA, device=DEV, generator=gen), torch.randn(128, N, device=DEV, generator=gen); ref = A.t() @ B
        for reps in (1, 2):
            D = torch.zeros(128, N, device=DEV)
            call("pag_tc_gemm_test16", mode, ptr(A), ptr(B), ptr(D), N, K, FA, reps)
            assert_close(D[:ref.shape[0]], ref * reps, rtol=2e-3, atol_scale=2e-3, msg=f"mode {mode} N {N} K {K} FA {FA} reps {reps}")


@pytest.mark.parametrize("name", ["trace_delta_permuto_ray", "trace_nef_tcnn_ray"])
@pytest.mark.parametrize("M", [300, 5000])
def test_tc_decoders_vs_fp32(cuda_lib, name, M):
    """fp16-operand tensor-core decoders (training mode) against the exact fp32 kernels: 2e-3 (north_star fp16 tolerance)."""
    g = load_golden(name)
    nef = build_cuda_nef(g, DEV)
    gen = torch.Generator().manual_seed(M)
    coords = (torch.rand(M, 1, 3, generator=gen) * 2 - 1).to(DEV)
    ray_d = torch.nn.functional.normalize(torch.randn(M, 3, generator=gen), dim=-1).to(DEV)
    chans = {'density', 'rgb', 'semantics', 'inst_embedding'}
    gws = None
    res = {}
    for prec in ('fp32', 'fp16'):
        nef.decoder_precision = prec
        nef.zero_grad(set_to_none=True)
        ct, dt = coords.clone().requires_grad_(True), ray_d.clone().requires_grad_(True)
        out = nef(coords=ct, ray_d=dt, channels=chans)
        if gws is None:      # sorted: set order follows the per-process string hash, which made the drawn weights (and the test) vary per run
            gws = {c: torch.randn(out[c].shape, generator=gen).to(DEV) * 1e-3 for c in sorted(chans)}
        sum((out[c] * gws[c]).sum() for c in chans).backward()
        res[prec] = ({c: out[c].detach() for c in chans}, {k: p.grad.clone() for k, p in nef.named_parameters()}, ct.grad, dt.grad)
    for c in chans:
        assert_close(res['fp16'][0][c], res['fp32'][0][c], msg=c, rtol=2e-3, atol_scale=2e-3)
    # few samples / the golden's small random field: one flipped hidden unit is a visible share of a gradient (measured up to
    # 3.1e-2 relative l2 at M = 300 and 2.2e-2 at M = 5000; at the benchmarked shapes -- 25 k samples, L = 24 -- the worst tensor
    # is at 6e-3, the same as the reference's own autocast step: tests/test_gpu_bench_shapes.py)
    for k in res['fp32'][1]:
        assert_close_norm(res['fp16'][1][k], res['fp32'][1][k], rel_l2=4e-2 if M < 1000 else 2.5e-2, msg="grad " + k)
    # d/d coords passes through every ReLU mask of both MLP chains (measured 3e-2 .. 7e-2 relative l2 at M = 300)
    assert_close_norm(res['fp16'][2], res['fp32'][2], rel_l2=0.1 if M < 1000 else 5e-2, max_frac=0.3, msg="grad coords")
    assert_close_norm(res['fp16'][3], res['fp32'][3], rel_l2=4e-2, max_frac=0.2, msg="grad ray_d")


def test_tc_trace_under_autocast_matches_golden(cuda_lib):
    """Training-mode numerics (autocast -> fp16 coords rounding is bypassed by feeding fp16-exact rays; tensor-core
    decoders) through the tracer vs the reference-made golden at the fp16 tolerance."""
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_ray")
    nef = build_cuda_nef(g, DEV)
    nef.decoder_precision = 'fp16'
    tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=int(g["num_steps"]), bg_color='white')
    o = torch.from_numpy(g["o"]).to(DEV).requires_grad_(True)
    d = torch.from_numpy(g["d"]).to(DEV).requires_grad_(True)
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    rb = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
    for c in chans + ['alpha']:
        assert_close(getattr(rb, c), g["out_" + c], msg=c, rtol=3e-3, atol_scale=3e-3)
    loss = sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c]).to(DEV)).sum() for c in chans)
    loss.backward()
    gg = golden_grads(g)
    for k, p in nef.named_parameters():
        if k in gg:
            assert_close_norm(p.grad, gg[k], msg="grad " + k)


@pytest.mark.parametrize("name,mode", [("trace_delta_permuto_ray", "ray"), ("trace_delta_permuto_voxel", "voxel"), ("trace_nef_tcnn_ray", "ray")])
def test_fused_panoptic_composite_equals_modular(cuda_lib, name, mode):
    """decoder_tc_fused.cu (heads + compositing in one kernel) vs the modular tensor-core path
    (pan_tc kernels -> [M,C] probabilities -> composite kernel): same rounding points, so a tight tolerance."""
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden(name)
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    res = []
    for fused in (True, False):
        nef = build_cuda_nef(g, DEV)
        nef.decoder_precision = 'fp16'
        if not fused:
            nef.fused_panoptic_ok = lambda channels: False
        else:
            assert nef.fused_panoptic_ok(set(chans))
        tracer = PanopticPackedRFTracer(raymarch_type=mode, num_steps=int(g["num_steps"]),
                                        bg_color='white' if bool(g["bg_white"]) else 'black', ray_max_travel=float(g["ray_max_travel"]))
        o = torch.from_numpy(g["o"]).to(DEV).requires_grad_(True)
        d = torch.from_numpy(g["d"]).to(DEV).requires_grad_(True)
        rb = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
        loss = sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c]).to(DEV)).sum() for c in chans)
        loss.backward()
        res.append(({c: getattr(rb, c).detach() for c in chans + ['alpha']}, {k: p.grad.clone() for k, p in nef.named_parameters()}, o.grad, d.grad))
    for c in chans + ['alpha']:
        assert_close(res[0][0][c], res[1][0][c], rtol=1e-3, atol_scale=1e-3, msg=c)
    for k in res[0][1]:
        assert_close_norm(res[0][1][k], res[1][1][k], rel_l2=5e-3, max_frac=2e-2, msg="grad " + k)
    assert_close_norm(res[0][2], res[1][2], rel_l2=5e-3, max_frac=2e-2, msg="grad origins")


@pytest.mark.parametrize("with_pose_grad", [False, True])
@pytest.mark.parametrize("name", ["trace_delta_permuto_ray", "trace_nef_tcnn_ray"])
def test_sync_free_fused_trace_equals_stepwise(cuda_lib, with_pose_grad, name):
    """ops.FusedTraceFn (device-side sample count, worst-case buffers, no host sync) vs the step-by-step plugin path, for a
    permutohedral delta field and for a PanopticNeF on the tcnn-style hash grid (heads on the detached colour features)."""
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden(name)
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    res = []
    for fused in (True, False):
        nef = build_cuda_nef(g, DEV)
        nef.decoder_precision = 'fp16'
        tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=int(g["num_steps"]), bg_color='white')
        tracer.allow_fused = fused
        o = torch.from_numpy(g["o"]).to(DEV).requires_grad_(with_pose_grad)
        d = torch.from_numpy(g["d"]).to(DEV).requires_grad_(with_pose_grad)
        rb = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
        if fused:
            assert torch.is_tensor(tracer.last_num_samples), "fused path keeps the sample count on the device"
        loss = sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c]).to(DEV)).sum() for c in chans)
        loss.backward()
        res.append(({c: getattr(rb, c).detach() for c in chans + ['alpha', 'hit']}, {k: p.grad.clone() for k, p in nef.named_parameters()},
                    o.grad if with_pose_grad else None, d.grad if with_pose_grad else None))
    assert torch.equal(res[0][0]['hit'], res[1][0]['hit'])
    for c in chans + ['alpha']:
        assert_close(res[0][0][c], res[1][0][c], rtol=1e-3, atol_scale=1e-3, msg=c)
    for k in res[0][1]:
        assert_close_norm(res[0][1][k], res[1][1][k], rel_l2=5e-3, max_frac=2e-2, msg="grad " + k)
    if with_pose_grad:
        assert_close_norm(res[0][2], res[1][2], rel_l2=5e-3, max_frac=2e-2, msg="grad origins")
        assert_close_norm(res[0][3], res[1][3], rel_l2=5e-3, max_frac=2e-2, msg="grad dirs")


def test_cuda_graph_replay_matches_eager(cuda_lib):
    """The fused training step is capturable as one CUDA graph (no host sync, device-side sample count and jitter seed):
    replayed gradients == eager gradients for the same jitter stream, and the stream advances between replays."""
    import bench
    from pagnerf_b200.graph import GraphedStep
    dev = torch.device(DEV)
    wl = bench.Workload(dev, n_rays=2048, seed=0, n_batches=2)
    wl.keep_rb = False
    blas = wl.nef.grid.blas
    blas.jitter_seed = 5
    g = GraphedStep(wl.loss_of, wl.dev[0], wl.params, wl.nef)
    seed_before = int(blas.seed_tensor.item())
    assert blas.jitter_seed == seed_before, "host mirror of the device-side jitter seed"
    loss_g = float(g(*wl.dev[1]))
    grads_g = [p.grad.clone() for p in wl.params]
    assert int(blas.seed_tensor.item()) == seed_before + 1 == blas.jitter_seed
    # eager step with the same batch and the same jitter seed (eager traces follow blas.jitter_seed, not the device-side seed)
    blas.jitter_seed = seed_before
    for p in wl.params:
        p.grad = None
    loss_e = wl.loss_of(*wl.dev[1])
    loss_e.backward()
    assert abs(loss_g - float(loss_e)) <= 1e-5 * abs(float(loss_e))
    for a, p in zip(grads_g, wl.params):
        assert_close_norm(a, p.grad, rel_l2=1e-4, max_frac=1e-3, msg="graph vs eager grad")
    l2 = float(g(*wl.dev[1]))
    assert l2 != loss_g, "jitter stream must advance between replays"


def test_fused_trace_all_rays_miss(cuda_lib):
    """Edge case of the sync-free path: the device-side packed-sample count is 0 (every ray misses the octree)."""
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_ray")
    nef = build_cuda_nef(g, DEV)
    nef.decoder_precision = 'fp16'
    tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=16, bg_color='white')
    N = 37
    o = torch.full((N, 3), 5.0, device=DEV)
    d = torch.nn.functional.normalize(torch.ones(N, 3, device=DEV), dim=-1)
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    rb = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
    assert int(tracer.last_num_samples.item()) == 0
    assert torch.all(rb.rgb == 1.0) and torch.all(rb.alpha == 0) and not rb.hit.any()
    assert torch.all(rb.depth == 0) and torch.all(rb.semantics == 0) and torch.all(rb.inst_embedding == 0)
    (rb.rgb.sum() + rb.semantics.sum() + rb.inst_embedding.sum() + rb.depth.sum()).backward()
    for n, p in nef.named_parameters():
        assert p.grad is not None and torch.all(p.grad == 0), n


@pytest.mark.parametrize("N", [1, 31, 1000])
def test_fused_trace_ragged_sizes_vs_stepwise(cuda_lib, N):
    """Ragged ray counts (not multiples of the 128-sample tile / 32-lane warp) through both paths."""
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_ray")
    gen = torch.Generator().manual_seed(N)
    o = torch.stack([torch.rand(N, generator=gen) * 1.2 - 0.6, torch.rand(N, generator=gen) * 1.2 - 0.6, torch.full((N,), 0.9)], 1)
    tgt = torch.stack([torch.rand(N, generator=gen) * 1.8 - 0.9, torch.rand(N, generator=gen) * 1.8 - 0.9, -torch.rand(N, generator=gen) * 0.9], 1)
    d = torch.nn.functional.normalize(tgt - o, dim=-1)
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    outs = []
    for fused in (True, False):
        nef = build_cuda_nef(g, DEV)
        nef.decoder_precision = 'fp16'
        tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=24, bg_color='black')
        tracer.allow_fused = fused
        rb = tracer(nef, channels=chans, rays=Rays(origins=o.to(DEV), dirs=d.to(DEV), dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
        outs.append({c: getattr(rb, c).detach() for c in chans + ['alpha']})
    for c in chans + ['alpha']:
        assert_close(outs[0][c], outs[1][c], rtol=1e-3, atol_scale=1e-3, msg=c)


def test_fused_trace_live_compaction_is_exact(cuda_lib):
    """Dropping the zero-density samples after the density pass (ops.COMPACT_LIVE) must not change outputs or gradients."""
    from pagnerf_b200 import ops
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_ray")
    N = 300
    gen = torch.Generator().manual_seed(7)
    o = torch.stack([torch.rand(N, generator=gen) * 1.2 - 0.6, torch.rand(N, generator=gen) * 1.2 - 0.6, torch.full((N,), 0.9)], 1)
    tgt = torch.stack([torch.rand(N, generator=gen) * 1.8 - 0.9, torch.rand(N, generator=gen) * 1.8 - 0.9, -torch.rand(N, generator=gen) * 0.9], 1)
    d = torch.nn.functional.normalize(tgt - o, dim=-1)
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']

    def run(compact, shift):
        ops.COMPACT_LIVE = compact
        nef = build_cuda_nef(g, DEV)
        nef.decoder_precision = 'fp16'
        with torch.no_grad():      # density pre-activation without its +1 bias: a share of the samples is clamped to 0
            nef.decoder_density.lout.bias[0] = shift
            nef.decoder_density.lout.weight[0] *= 8.0
        tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=24, bg_color='black')
        rb = tracer(nef, channels=chans, rays=Rays(origins=o.to(DEV), dirs=d.to(DEV), dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
        loss = sum((getattr(rb, c).float() ** 2).sum() for c in chans) + rb.alpha.sum()
        loss.backward()
        return ({c: getattr(rb, c).detach() for c in chans + ['alpha']},
                {n: p.grad.detach().clone() for n, p in nef.named_parameters() if p.grad is not None},
                int(ops.FusedTraceFn.last_live_dev.item()))

    try:
        shift, m1 = None, None
        for cand in (0.0, 0.05, -0.05, 0.2, -0.2, 0.5, -0.5):
            _, _, total = run(False, cand)
            o1, g1, m1 = run(True, cand)
            if 0.15 * total < m1 < 0.85 * total:
                shift = cand
                break
        assert shift is not None, "no density-bias shift gave a mix of live and dead samples"
        o0, g0, m0 = run(False, shift)
    finally:
        ops.COMPACT_LIVE = False
    assert 0 < m1 < m0
    for c in chans + ['alpha']:
        assert_close(o1[c], o0[c], rtol=1e-5, atol_scale=1e-6, msg=c)
    assert set(g0) == set(g1)
    for n in g0:
        assert_close(g1[n], g0[n], rtol=1e-4, atol_scale=2e-5, msg=n)   # summation order (atomics, tile grouping) differs


@pytest.mark.parametrize("dd", [False, True])
def test_fused_trace_img16_interchange_equals_f32_rows(cuda_lib, dd):
    """fp16 operand-image interchange between encoders and tensor-core decoders (ops.IMG16: coalesced encoder stores, bulk
    tile copies, fp16 dX images) vs the f32 [M, 2L] row interchange, on the bench field (L = 24, 200 instances): same
    step, same jitter -> outputs at the fp16 tolerance, gradients norm-wise."""
    import bench
    from pagnerf_b200 import ops
    dev = torch.device(DEV)
    res = []
    try:
        for img in (True, False):
            ops.IMG16 = img
            wl = bench.Workload(dev, n_rays=2048, seed=0, n_batches=2, dd=dd)      # dd: DD field + tracer (panoptic density stream)
            wl.keep_rb = True
            blas = wl.nef.grid.blas
            blas.fixed_jitter, blas.jitter_seed = True, 11
            for p in wl.params:
                p.grad = None
            loss = wl.loss_of(*wl.dev[0])
            loss.backward()
            rb = wl.last_rb
            res.append((float(loss), {c: getattr(rb, c).detach().clone() for c in ('rgb', 'depth', 'alpha', 'semantics', 'inst_embedding')},
                        [p.grad.clone() for p in wl.params]))
    finally:
        ops.IMG16 = True
    (l1, o1, g1), (l0, o0, g0) = res
    assert abs(l1 - l0) <= 2e-3 * abs(l0)
    for c in o0:
        assert_close(o1[c], o0[c], rtol=5e-3, atol_scale=5e-3, msg=c)
    for a, b in zip(g1, g0):
        assert torch.isfinite(a).all()
        assert_close_norm(a, b, rel_l2=5e-2, max_frac=0.3, msg="img16 vs f32-row grad")


def test_fused_dd_trace_equals_stepwise(cuda_lib):
    """PanopticDDensityNeF + PanopticDDensityPackedRFTracer: the sync-free fused trace with the panoptic density stream
    (weights carrying gradient: heads backward -> <p, g> per sample -> reverse scan -> ReLU gate -> delta-density head) vs the
    step-by-step path that is pinned by the reference-made golden."""
    from pagnerf_b200.tracers import PanopticDDensityPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_dd_permuto_ray")
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    res = []
    for fused in (True, False):
        nef = build_cuda_nef(g, DEV)
        nef.decoder_precision = 'fp16'
        tracer = PanopticDDensityPackedRFTracer(raymarch_type='ray', num_steps=int(g["num_steps"]), bg_color='white')
        tracer.allow_fused = fused
        o = torch.from_numpy(g["o"]).to(DEV).requires_grad_(True)
        d = torch.from_numpy(g["d"]).to(DEV).requires_grad_(True)
        rb = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
        if fused:
            assert torch.is_tensor(tracer.last_num_samples), "the DD tracer must take the fused path"
        loss = sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c]).to(DEV)).sum() for c in chans)
        loss.backward()
        res.append(({c: getattr(rb, c).detach() for c in chans + ['alpha', 'hit']},
                    {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in nef.named_parameters()}, o.grad, d.grad))
    assert torch.equal(res[0][0]['hit'], res[1][0]['hit'])
    for c in chans + ['alpha']:
        assert_close(res[0][0][c], res[1][0][c], rtol=2e-3, atol_scale=2e-3, msg=c)
    for k in res[0][1]:
        assert_close_norm(res[0][1][k], res[1][1][k], rel_l2=2e-2, max_frac=5e-2, msg="grad " + k)
    assert_close_norm(res[0][2], res[1][2], rel_l2=2e-2, max_frac=5e-2, msg="grad origins")
    assert_close_norm(res[0][3], res[1][3], rel_l2=2e-2, max_frac=5e-2, msg="grad dirs")


@pytest.mark.parametrize("chans", [['rgb'], ['rgb', 'depth'], ['semantics'], ['rgb', 'inst_embedding'], ['depth', 'semantics', 'inst_embedding']])
def test_fused_trace_channel_subsets(cuda_lib, chans):
    """Channel gating of the fused training trace (only the requested heads / outputs are computed) vs the step-by-step path."""
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_ray")
    res = []
    for fused in (True, False):
        nef = build_cuda_nef(g, DEV)
        nef.decoder_precision = 'fp16'
        tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=int(g["num_steps"]), bg_color='white')
        tracer.allow_fused = fused
        o, d = torch.from_numpy(g["o"]).to(DEV), torch.from_numpy(g["d"]).to(DEV)
        rb = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
        if fused:
            assert torch.is_tensor(tracer.last_num_samples)
        sum((getattr(rb, c).float() * torch.from_numpy(g["gw_" + c]).to(DEV)).sum() for c in chans).backward()
        # a parameter "receives gradient" if its .grad is a non-zero tensor (unused heads: None or exact zeros)
        res.append(({c: getattr(rb, c).detach() for c in chans + ['alpha']},
                    {k: p.grad.clone() for k, p in nef.named_parameters() if p.grad is not None and bool((p.grad != 0).any())}))
        assert all(torch.isfinite(v).all() for v in res[-1][1].values())
        for c in ('rgb', 'depth', 'semantics', 'inst_embedding'):
            if c not in chans:
                assert getattr(rb, c, None) is None, f"{c} was not requested"
    for c in chans + ['alpha']:
        assert_close(res[0][0][c], res[1][0][c], rtol=1e-3, atol_scale=1e-3, msg=c)
    assert set(res[0][1]) == set(res[1][1]), "the same parameters receive gradient on both paths"
    for k in res[0][1]:
        assert_close_norm(res[0][1][k], res[1][1][k], rel_l2=5e-3, max_frac=2e-2, msg="grad " + k)


# ------------------------------------------------------------------------------------------------
# sync-free 'voxel' marching (the trainer's mode from epoch 201 on: configs/bup20/best.yaml:34, pc_nerf/trainer.py:362-366)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("max_travel", [None, 0.35])
def test_voxel_device_side_march_bit_exact(cuda_lib, max_travel):
    """Device-side nugget filter + sample emit (no host sync, worst-case buffers) == raytrace -> voxel samples -> max-travel
    filter of the step-by-step path and of the oracle, bit for bit (packed order, ray ids, positions, depths, deltas)."""
    from oracle import spc as ospc, raymarch as orm
    from pagnerf_b200 import ops, spc
    from pagnerf_b200._lib import call, ptr
    level, S, N, seed = 5, 3, 700, 11
    pts = _scene(level, seed=4, frac=0.2)
    octree = spc.unbatched_points_to_octree(torch.from_numpy(pts), level)
    blas = spc.OctreeAS(DEV)
    blas.init(octree)
    o_np, d_np = _rays(N, seed=5)
    o, d = torch.from_numpy(o_np).to(DEV), torch.from_numpy(d_np).to(DEV)
    # reference: step-by-step ops (each bit-exact vs the oracle in test_raymarch_voxel_bit_exact_and_filter)
    ridx, pidx, samples, depths, deltas, boundary, off = ops.raymarch_voxel(blas.octree, blas.prefix, o, d, level, S, seed=seed)
    if max_travel is not None:
        keep = ops.max_travel_mask(ridx, depths, ops.ray_offsets(ridx, N), max_travel)
        assert 0 < int(keep.sum()) < keep.numel(), "the filter must drop something for this test to mean anything"
        deltas = deltas.reshape(depths.shape)[keep].reshape(-1)
        ridx, samples, depths = ridx[keep], samples[keep], depths[keep]
    ref_ridx = ridx.repeat_interleave(S)
    # device-side chain
    Kmax = N * (3 * (1 << level) - 2)
    counts = torch.empty(N, dtype=torch.int32, device=DEV)
    nug_off = torch.empty(N + 1, dtype=torch.int64, device=DEV)
    call("pag_raytrace_count", ptr(blas.octree), ptr(blas.prefix), ptr(o), ptr(d), N, level, ptr(counts), ptr(nug_off))
    assert int(nug_off[-1]) <= Kmax
    nr = torch.empty(Kmax, dtype=torch.int64, device=DEV); npx = torch.empty(Kmax, dtype=torch.int64, device=DEV)
    nd = torch.empty(Kmax, 2, device=DEV)
    call("pag_raytrace_emit", ptr(blas.octree), ptr(blas.prefix), ptr(o), ptr(d), N, level, ptr(nug_off), ptr(nr), ptr(npx), ptr(nd))
    rel = torch.empty(Kmax, dtype=torch.int32, device=DEV)
    offsets = torch.empty(N + 1, dtype=torch.int64, device=DEV)
    call("pag_voxel_filter_count", ptr(nd), ptr(nug_off), N, S, seed, None, float(max_travel or 0.0), int(max_travel is not None),
         ptr(rel), ptr(counts), ptr(offsets))
    M = int(offsets[-1])
    assert M == ref_ridx.shape[0]
    r2 = torch.full((Kmax * S,), -1, dtype=torch.int64, device=DEV)
    s2 = torch.zeros(Kmax * S, 3, device=DEV); dp2 = torch.zeros(Kmax * S, device=DEV); dl2 = torch.zeros(Kmax * S, device=DEV)
    call("pag_voxel_emit_dyn", ptr(o), ptr(d), ptr(nr), ptr(nd), ptr(rel), ptr(nug_off), ptr(offsets), N, Kmax, S, seed, None,
         ptr(r2), ptr(s2), ptr(dp2), ptr(dl2))
    assert torch.equal(r2[:M], ref_ridx)
    assert torch.equal(s2[:M], samples.reshape(-1, 3)), "sample positions bit-exact"
    assert torch.equal(dp2[:M], depths.reshape(-1)) and torch.equal(dl2[:M], deltas.reshape(-1))
    assert torch.equal(offsets, ops.ray_offsets(ridx, N) * S)
    # the one-traversal staged chain the fused trace uses: same counts, same packed samples
    cap = 3 * (1 << level) - 2
    counts3 = torch.empty(N, dtype=torch.int32, device=DEV)
    nug_off3 = torch.empty(N + 1, dtype=torch.int64, device=DEV)
    stage = torch.full((N * cap, 2), float('nan'), device=DEV)
    call("pag_raytrace_stage", ptr(blas.octree), ptr(blas.prefix), ptr(o), ptr(d), N, level, cap, ptr(counts3), ptr(nug_off3), ptr(stage), None)
    assert torch.equal(nug_off3, nug_off)
    # ... and its 64-way split (thread per (ray, child of root, grandchild)): identical rows
    stage64 = torch.full((N * cap, 2), float('nan'), device=DEV)
    slots = torch.empty(N * 64, dtype=torch.int32, device=DEV)
    nug_off4 = torch.empty(N + 1, dtype=torch.int64, device=DEV)
    call("pag_raytrace_stage", ptr(blas.octree), ptr(blas.prefix), ptr(o), ptr(d), N, level, cap, ptr(counts3), ptr(nug_off4), ptr(stage64), ptr(slots))
    assert torch.equal(nug_off4, nug_off)
    assert torch.equal(torch.nan_to_num(stage64, nan=-7.0), torch.nan_to_num(stage, nan=-7.0)), "split traversal == per-ray DFS, bit for bit"
    rel3 = torch.empty(N * cap, dtype=torch.int32, device=DEV)
    offsets3 = torch.empty(N + 1, dtype=torch.int64, device=DEV)
    call("pag_voxel_filter_count_staged", ptr(stage), ptr(nug_off3), N, cap, S, seed, None, float(max_travel or 0.0), int(max_travel is not None),
         ptr(rel3), ptr(counts3), ptr(offsets3))
    assert torch.equal(offsets3, offsets)
    r3 = torch.full((Kmax * S,), -1, dtype=torch.int64, device=DEV)
    s3 = torch.zeros(Kmax * S, 3, device=DEV); dp3 = torch.zeros(Kmax * S, device=DEV); dl3 = torch.zeros(Kmax * S, device=DEV)
    call("pag_voxel_emit_staged", ptr(o), ptr(d), ptr(stage), ptr(rel3), ptr(nug_off3), ptr(offsets3), N, cap, S, seed, None,
         ptr(r3), ptr(s3), ptr(dp3), ptr(dl3))
    assert torch.equal(r3[:M], r2[:M]) and torch.equal(s3[:M], s2[:M]) and torch.equal(dp3[:M], dp2[:M]) and torch.equal(dl3[:M], dl2[:M])


@pytest.mark.parametrize("with_pose_grad", [False, True])
def test_sync_free_fused_trace_equals_stepwise_voxel(cuda_lib, with_pose_grad):
    """ops.FusedTraceFn with 'voxel' marching (nuggets, max-travel filter and samples on the device) vs the step-by-step plugin
    path on the reference-made voxel golden's field, rays and filter threshold; tensor-core decoders on both sides."""
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_voxel")
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    res = []
    for fused in (True, False):
        nef = build_cuda_nef(g, DEV)
        nef.decoder_precision = 'fp16'
        tracer = PanopticPackedRFTracer(raymarch_type='voxel', num_steps=int(g["num_steps"]),
                                        bg_color='white' if bool(g["bg_white"]) else 'black', ray_max_travel=float(g["ray_max_travel"]))
        tracer.allow_fused = fused
        o = torch.from_numpy(g["o"]).to(DEV).requires_grad_(with_pose_grad)
        d = torch.from_numpy(g["d"]).to(DEV).requires_grad_(with_pose_grad)
        rb = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
        assert torch.is_tensor(tracer.last_num_samples) == fused, "fused path keeps the sample count on the device"
        n = int(tracer.last_num_samples)
        loss = sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c]).to(DEV)).sum() for c in chans)
        loss.backward()
        res.append(({c: getattr(rb, c).detach() for c in chans + ['alpha', 'hit']}, {k: p.grad.clone() for k, p in nef.named_parameters()},
                    o.grad if with_pose_grad else None, d.grad if with_pose_grad else None, n))
    assert res[0][4] == res[1][4] > 0, "same packed-sample count"
    assert torch.equal(res[0][0]['hit'], res[1][0]['hit'])
    for c in chans + ['alpha']:
        assert_close(res[0][0][c], res[1][0][c], rtol=1e-3, atol_scale=1e-3, msg=c)
    for k in res[0][1]:
        assert_close_norm(res[0][1][k], res[1][1][k], rel_l2=5e-3, max_frac=2e-2, msg="grad " + k)
    if with_pose_grad:
        assert_close_norm(res[0][2], res[1][2], rel_l2=5e-3, max_frac=2e-2, msg="grad origins")
        assert_close_norm(res[0][3], res[1][3], rel_l2=5e-3, max_frac=2e-2, msg="grad dirs")
    # and against the reference-made golden outputs at the fp16 tolerance
    for c in chans + ['alpha']:
        assert_close(res[0][0][c], g["out_" + c], msg="golden " + c, rtol=3e-3, atol_scale=3e-3)


# ------------------------------------------------------------------------------------------------
# camera-pose transform (BAPipeline.transform_rays, pc_nerf/ba_pipeline.py:85-92) forward + backward
# ------------------------------------------------------------------------------------------------
def _random_poses(n, gen):
    """n view matrices: random rotations (QR) + translations."""
    q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=gen))
    q = q * torch.sign(torch.linalg.det(q))[:, None, None]
    V = torch.eye(4).repeat(n, 1, 1)
    V[:, :3, :3] = q
    V[:, :3, 3] = torch.randn(n, 3, generator=gen) * 0.5
    return V


@pytest.mark.parametrize("B", [1, 37, 4096])
def test_pose_transform_forward_backward(cuda_lib, B):
    from oracle import pose as opose
    from pagnerf_b200 import ops
    gen = torch.Generator().manual_seed(B)
    n_cam, C = 7, 5
    V = _random_poses(n_cam, gen)
    params = opose.params_from_view_matrix(V)
    params = params + 0.05 * torch.randn(n_cam, 9, generator=gen)      # off the orthonormal manifold: Gram-Schmidt must do work
    cam_idx = torch.tensor([3, 0, 6, 3, 1])                              # a camera may repeat inside a batch
    bo = torch.randn(C * B, 3, generator=gen) * 0.1
    bd = torch.nn.functional.normalize(torch.randn(C * B, 3, generator=gen), dim=-1)
    go, gd = torch.randn(C * B, 3, generator=gen), torch.randn(C * B, 3, generator=gen)
    p_ref = params.clone().double().requires_grad_(True)
    o_ref, d_ref = opose.transform_rays(p_ref, cam_idx, bo.double(), bd.double())
    ((o_ref * go.double()).sum() + (d_ref * gd.double()).sum()).backward()
    p = params.clone().to(DEV).requires_grad_(True)
    o, d = ops.pose_transform(p, cam_idx.to(DEV), bo.to(DEV), bd.to(DEV))
    assert_close(o, o_ref.float(), rtol=1e-5, atol_scale=1e-6, msg="world origins")
    assert_close(d, d_ref.float(), rtol=1e-5, atol_scale=1e-6, msg="world dirs")
    ((o * go.to(DEV)).sum() + (d * gd.to(DEV)).sum()).backward()
    assert_close(p.grad, p_ref.grad.float(), rtol=1e-4, atol_scale=1e-5, msg="pose parameter gradient")
    assert float(p.grad[[2, 4, 5]].abs().max()) == 0.0, "cameras that are not in the batch receive no gradient"


def test_ba_pipeline_pose_gradients_through_fused_trace(cuda_lib):
    """BAPipeline (pose table + nef + tracer): d loss / d camera_extrinsics through the fused training trace equals the chain
    rule applied by the oracle to the trace's own d/d origins, d/d dirs; anchor frames keep a zero gradient (:53-62)."""
    from oracle import pose as opose
    from pagnerf_b200.pc_nerf.ba_pipeline import BAPipeline
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_ray")
    nef = build_cuda_nef(g, DEV)
    nef.decoder_precision = 'fp16'
    tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=int(g["num_steps"]), bg_color='white')
    gen = torch.Generator().manual_seed(0)
    N = g["o"].shape[0]
    C = 4
    B = N // C
    V = torch.eye(4).repeat(3, 1, 1)
    V[:, :3, 3] = torch.randn(3, 3, generator=gen) * 0.02
    # camera-space base rays chosen so that the identity-ish poses reproduce the golden's world rays approximately
    bo = torch.from_numpy(g["o"][:C * B]).clone()
    bd = torch.from_numpy(g["d"][:C * B]).clone()
    pipe = BAPipeline(nef, V, tracer, anchor_frame_idxs=[1], near=0.0, far=2.0).to(DEV)
    cam_ids = torch.tensor([0, 1, 2, 0])
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    rb = pipe(channels=chans, rays=Rays(origins=bo.to(DEV), dirs=bd.to(DEV)), cam_ids=cam_ids, lod_idx=None, stage='train')
    loss = sum((getattr(rb, c) * torch.from_numpy(g["gw_" + c][:C * B]).to(DEV)).sum() for c in chans)
    loss.backward()
    gp = pipe.camera_extrinsics.grad
    assert gp is not None and torch.isfinite(gp).all() and float(gp.abs().max()) > 0
    assert float(gp[1].abs().max()) == 0.0, "anchor frame"
    # chain rule check: the same trace on detached world rays gives d/d(o, d); the oracle maps them to the parameters
    rays_w = pipe.transform_rays(Rays(origins=bo.to(DEV), dirs=bd.to(DEV)), cam_ids)
    o = rays_w.origins.detach().requires_grad_(True)
    d = rays_w.dirs.detach().requires_grad_(True)
    nef.zero_grad(set_to_none=True)
    rb2 = tracer(nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=0.0, dist_max=2.0), lod_idx=None, stage='train')
    sum((getattr(rb2, c) * torch.from_numpy(g["gw_" + c][:C * B]).to(DEV)).sum() for c in chans).backward()
    p_ref = pipe.camera_extrinsics.detach().cpu().double().requires_grad_(True)
    o_ref, d_ref = opose.transform_rays(p_ref, cam_ids, bo.double(), bd.double())
    ((o_ref * o.grad.cpu().double()).sum() + (d_ref * d.grad.cpu().double()).sum()).backward()
    ref = p_ref.grad.float()
    ref[1] = 0.0
    assert_close(gp, ref, rtol=1e-3, atol_scale=1e-4, msg="pose gradient through the trace")


def test_fused_adam_matches_torch_adam(cuda_lib):
    """pagnerf_b200.optim.FusedAdam (one multi-tensor launch, device-side step count) vs torch.optim.Adam over the reference's kind of
    parameter groups (per-group lr / weight decay, pc_nerf/trainer.py:229-300): 5 steps, odd sizes, a parameter without gradient."""
    from pagnerf_b200.optim import FusedAdam
    gen = torch.Generator().manual_seed(0)
    shapes = [(64, 48), (64,), (200, 64), (7,), (1,), (24, 1000, 2), (5, 9)]
    ref_p = [torch.randn(s, generator=gen).to(DEV).requires_grad_(True) for s in shapes]
    our_p = [p.detach().clone().requires_grad_(True) for p in ref_p]

    def groups(ps):
        return [dict(params=ps[:3], lr=1e-3), dict(params=ps[3:5], lr=1e-3, weight_decay=0.01),
                dict(params=ps[5:6], lr=1e-1, weight_decay=0.0), dict(params=ps[6:], lr=1e-4)]

    ref = torch.optim.Adam(groups(ref_p), eps=1e-15)
    ours = FusedAdam(groups(our_p), eps=1e-15)
    for it in range(5):
        for k, (a, b) in enumerate(zip(ref_p, our_p)):
            if k == 4 and it < 2:
                a.grad = b.grad = None          # no gradient yet: skipped, its moments start later
                continue
            g = torch.randn(a.shape, generator=gen).to(DEV) * (10.0 ** (k - 3))
            a.grad, b.grad = g.clone(), g.clone()
        ref.step(); ours.step()
    for k, (a, b) in enumerate(zip(ref_p, our_p)):
        if k == 4:
            continue      # started late: torch keeps a per-parameter step count, the fused kernel a per-bucket one (documented)
        assert_close(b, a.detach(), rtol=2e-6, atol_scale=1e-6, msg=f"param {k}")
    st = ours.state[our_p[0]]
    assert set(st) == {'exp_avg', 'exp_avg_sq', 'step'} and int(st['step']) == 5


# ------------------------------------------------------------------------------------------------
# prune(): device-side octree rebuild, checkpoint keys, graph-safe occupancy refresh
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("level,frac", [(1, 0.5), (3, 0.3), (5, 0.05), (7, 0.02), (4, 0.0), (4, 1.0)])
def test_octree_from_mask_matches_oracle(cuda_lib, level, frac):
    """pag_octree_from_mask (dense Morton-ordered mask -> SPC octree / points / prefix / pyramid) == oracle points_to_octree + scan."""
    from oracle import spc as ospc
    from pagnerf_b200 import spc
    n = 8 ** level
    gen = torch.Generator().manual_seed(level)
    mask = torch.rand(n, generator=gen) < frac if 0.0 < frac < 1.0 else torch.full((n,), frac >= 1.0)
    octree, points, pyramid, prefix = spc.octree_from_mask(mask.to(DEV), level)
    codes = torch.nonzero(mask).flatten()
    if codes.numel() == 0:
        assert octree.numel() == 0 and points.numel() == 0 and int(pyramid[1, level + 1]) == 0
        return
    pts = spc.morton_decode(codes, level).numpy()
    ref_oct = ospc.points_to_octree(pts, level)
    ref_pts, ref_pyr, ref_pre = ospc.scan_octree(ref_oct, level)
    assert np.array_equal(octree.cpu().numpy(), ref_oct)
    assert np.array_equal(points.cpu().numpy(), ref_pts)
    assert np.array_equal(prefix.cpu().numpy(), ref_pre)
    assert np.array_equal(pyramid.numpy(), ref_pyr)


def test_prune_rebuilds_octree_and_keeps_checkpoint_keys(cuda_lib):
    """PanopticDeltaNeF.prune() on the GPU (pc_nerf/panoptic_delta_nef.py:63-104): the pruned octree equals the oracle's build
    from the surviving cells, both grids share it, marching works against it, state_dict keys are unchanged and the checkpoint
    round-trips; a cached occupancy bit field is refreshed in place (same address: CUDA graphs stay valid)."""
    from oracle import spc as ospc
    from pagnerf_b200 import spc
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_ray")
    nef = build_cuda_nef(g, DEV)
    level = nef.grid.blas_level
    for gr in (nef.grid, nef.delta_grid):
        gr.blas.init_dense(level)
        gr.blas.to(DEV)
        gr._register_blas_buffers()
    keys0 = set(nef.state_dict().keys())
    bits = nef.grid.blas.level_bits(level)
    addr0 = bits.data_ptr()
    assert int(bits.view(torch.uint8).sum()) == 255 * bits.numel() * 4, "dense octree: every occupancy bit set"
    # a running occupancy that survives the 0.6 decay in about half of the cells (threshold 0.01*512/sqrt(3) = 2.96)
    torch.manual_seed(0)
    nef.grid.occupancy = torch.rand(8 ** level) * 10.0
    with torch.no_grad():      # the field itself contributes no density here: the running occupancy alone decides
        nef.decoder_density.lout.weight[0].zero_()
        nef.decoder_density.lout.bias[0] = -1.0
    nef.prune()
    occ = nef.grid.occupancy
    mask = (occ > (0.01 * 512) / np.sqrt(3)).cpu()
    assert 0 < int(mask.sum()) < mask.numel(), f"prune kept {int(mask.sum())} of {mask.numel()} cells: the test needs a partial octree"
    pts = nef.grid.dense_points.cpu()[mask].numpy()
    ref_oct = ospc.points_to_octree(pts, level)
    ref_pts, ref_pyr, ref_pre = ospc.scan_octree(ref_oct, level)
    for gr in (nef.grid, nef.delta_grid):
        assert np.array_equal(gr.blas.octree.cpu().numpy(), ref_oct)
        assert np.array_equal(gr.blas.points.cpu().numpy(), ref_pts)
        assert np.array_equal(gr.blas.prefix.cpu().numpy(), ref_pre)
        assert np.array_equal(gr.blas.pyramid.numpy(), ref_pyr)
    assert set(nef.state_dict().keys()) == keys0, "prune must not change the checkpoint keys"
    bits2 = nef.grid.blas.level_bits(level)
    assert bits2.data_ptr() == addr0, "occupancy bit field refreshed in place"
    assert int(torch.tensor([bin(int(x) & 0xFFFFFFFF).count('1') for x in bits2.cpu().tolist()]).sum()) == int(mask.sum())
    # the pruned field still traces, and the checkpoint round-trips into a fresh model (octree adopted from the state_dict)
    tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=int(g["num_steps"]), bg_color='white')
    rays = Rays(origins=torch.from_numpy(g["o"]).to(DEV), dirs=torch.from_numpy(g["d"]).to(DEV), dist_min=0.0, dist_max=2.0)
    rb = tracer(nef, channels=['rgb', 'depth', 'semantics', 'inst_embedding'], rays=rays, lod_idx=None, stage='val')
    assert torch.isfinite(rb.rgb).all()
    nef2 = build_cuda_nef(g, DEV)
    nef2.load_state_dict(nef.state_dict())
    assert torch.equal(nef2.grid.blas.octree.cpu(), nef.grid.blas.octree.cpu())
    for gr in (nef.grid, nef.delta_grid, nef2.grid, nef2.delta_grid):
        gr.blas.fixed_jitter, gr.blas.jitter_seed = True, 3
    rb2 = tracer(nef2, channels=['rgb', 'depth', 'semantics', 'inst_embedding'], rays=rays, lod_idx=None, stage='val')
    rb1 = tracer(nef, channels=['rgb', 'depth', 'semantics', 'inst_embedding'], rays=rays, lod_idx=None, stage='val')
    assert torch.equal(rb1.rgb, rb2.rgb) and torch.equal(rb1.inst_embedding, rb2.inst_embedding)


def test_full_size_step_on_a_dirty_allocator(cuda_lib):
    """Every torch.empty() of the fused step is served from NaN-filled memory: a read of a buffer nobody wrote, or of a tensor that
    went back to the caching allocator while another stream was still using it, turns into non-finite gradients.  (Regression:
    a gradient image freed by the panoptic chain was re-used by the colour chain while the delta-grid scatter was still reading.)"""
    import bench
    dev = torch.device(DEV)
    x = torch.full((3_000_000_000 // 4,), float('nan'), device=dev)
    del x
    ref = None
    for rep in range(3):
        wl = bench.Workload(dev, n_rays=16384, seed=0, n_batches=1)
        r = wl.forward_backward()
        torch.cuda.synchronize()
        grads = {n: p.grad for n, p in wl.nef.named_parameters()}
        for n, g in grads.items():
            assert g is not None and torch.isfinite(g).all(), f"rep {rep}: {n}"
        if ref is None:
            ref = {n: g.clone() for n, g in grads.items()}
        else:      # same rays, same parameters, same jitter stream: only the order of the atomics may differ
            for n, g in grads.items():
                assert_close_norm(g, ref[n], rel_l2=1e-4, max_frac=1e-3, msg=f"rep {rep} vs rep 0: {n}")
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------
# device-side instance loss with linear assignment (loss/lin_assignment_things.py)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,R,outlier", [(20, 257, False), (200, 1024, False), (200, 4096, True), (8, 64, False)])
def test_lin_assignment_things_loss_matches_oracle(cuda_lib, C, R, outlier):
    from oracle.losses import lin_assignment_things_loss
    from pagnerf_b200.loss import LinAssignmentThingsLoss
    gen = torch.Generator().manual_seed(C + R)
    B = 3
    p = torch.softmax(torch.randn(B, R, C, generator=gen) * 2.0, -1)
    n_inst = [min(C + 5, 30), 5, 1]                                       # more labels than ids in image 0 when C is small
    gt = torch.stack([torch.randint(0, n + 1, (R,), generator=gen) * 7 for n in n_inst])      # ids 0, 7, 14, ... (0 = no instance)
    gt[2, : R // 2] = 0
    stuff = torch.rand(B, R, generator=gen) < 0.3
    pts = torch.rand(B, R, 3, generator=gen) * 2 - 1 if outlier else None
    pr = p.clone().requires_grad_(True)
    ref, virt_ref = lin_assignment_things_loss(pr, gt, stuff, pts)
    w = torch.rand(B, R, generator=gen)
    (ref * w).sum().backward()
    loss_fn = LinAssignmentThingsLoss(outlier_rejection=outlier)
    pd = p.clone().to(DEV).requires_grad_(True)
    out = loss_fn(pd, gt.to(DEV), stuff.to(DEV), pts.to(DEV) if outlier else None)
    assert torch.equal(loss_fn.last_virtual_labels.cpu().long(), virt_ref), "virtual labels == scipy's assignment"
    assert_close(out, ref.detach(), rtol=1e-5, atol_scale=1e-6, msg="loss")
    (out * w.to(DEV)).sum().backward()
    assert_close(pd.grad, pr.grad, rtol=1e-5, atol_scale=1e-6, msg="grad probabilities")


def test_lin_assignment_no_wrong_pixel_gives_zero_loss(cuda_lib):
    """An image whose arg-max prediction already equals the virtual labels contributes no loss (:84)."""
    from pagnerf_b200.loss import LinAssignmentThingsLoss
    B, R, C = 2, 128, 10
    gt = torch.zeros(B, R, dtype=torch.int64)
    gt[0, :40], gt[0, 40:90] = 3, 9            # two instances in image 0; image 1 is all stuff
    stuff = gt == 0
    p = torch.full((B, R, C), 0.01)
    p[0, :40, 5], p[0, 40:90, 2] = 0.9, 0.9    # confident, consistent ids: the assignment maps 3 -> id 5, 9 -> id 2
    p[0, 90:, 0], p[1, :, 0] = 0.9, 0.9
    p = p / p.sum(-1, keepdim=True)
    out = LinAssignmentThingsLoss()(p.to(DEV), gt.to(DEV), stuff.to(DEV))
    assert float(out.abs().max()) == 0.0
    p[0, 0, 5], p[0, 0, 7] = 0.01, 0.9         # one wrong pixel: the whole image 0 is trained, image 1 still is not
    p = p / p.sum(-1, keepdim=True)
    out = LinAssignmentThingsLoss()(p.to(DEV), gt.to(DEV), stuff.to(DEV))
    assert float(out[0].min()) > 0.0 and float(out[1].abs().max()) == 0.0


def test_fused_panoptic_loss(cuda_lib):
    """pagnerf_b200.loss.panoptic_loss (one launch each way) == the torch formulation of the step's loss (bench.loss_fn), value and
    gradients; channels may be absent; the self-resetting scratch survives repeated calls."""
    import bench, os
    from pagnerf_b200.loss import panoptic_loss
    gen = torch.Generator().manual_seed(0)
    N, Cs, Ci = 3001, 6, 200
    rgb = torch.rand(N, 3, generator=gen); sem = torch.softmax(torch.randn(N, Cs, generator=gen), -1) * torch.rand(N, 1, generator=gen)
    inst = torch.softmax(torch.randn(N, Ci, generator=gen), -1) * torch.rand(N, 1, generator=gen)
    inst[5] = 0.0                                       # a ray without samples: log(0 + 1e-27)
    tr, ts, ti = torch.rand(N, 3, generator=gen), torch.randint(0, Cs, (N,), generator=gen), torch.randint(0, Ci, (N,), generator=gen)
    os.environ["BENCH_TORCH_LOSS"] = "1"
    try:
        a = [t.clone().requires_grad_(True) for t in (rgb, sem, inst)]
        ref = bench.loss_fn(*a, tr, ts, ti)
        (ref * 3.0).backward()
    finally:
        del os.environ["BENCH_TORCH_LOSS"]
    for rep in range(2):
        b = [t.clone().to(DEV).requires_grad_(True) for t in (rgb, sem, inst)]
        out = panoptic_loss(*b, tr.to(DEV), ts.to(DEV), ti.to(DEV), 10.0, 0.1, 1.0, 1e-27)
        (out * 3.0).backward()
        assert abs(float(out) - float(ref)) <= 1e-5 * abs(float(ref))
        for x, y, name in zip(b, a, ("rgb", "sem", "inst")):
            assert_close(x.grad, y.grad, rtol=1e-5, atol_scale=1e-7, msg="grad " + name)
    only = panoptic_loss(None, b[1].detach().requires_grad_(True), None, None, ts.to(DEV), None, 0.0, 1.0, 0.0)
    exp = -torch.log(sem.gather(1, ts[:, None]) + 1e-27).mean()
    assert abs(float(only) - float(exp)) <= 1e-5 * abs(float(exp))


# ------------------------------------------------------------------------------------------------
# register-tiled exact-FP32 inference decoders (csrc/decoder_tiled.cu)
# ------------------------------------------------------------------------------------------------
def _rand_linear(out_f, in_f, gen, scale=1.0):
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * (scale / in_f ** 0.5)
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * 0.1
    return [w.to(DEV), b.to(DEV)]


@pytest.mark.parametrize("M,IN,S", [(1, 48, 1), (127, 48, 1), (128, 12, 1), (1000, 48, 8), (40000, 28, 1)])
def test_tiled_dc_forward_equals_one_sample_per_thread(cuda_lib, M, IN, S):
    """Same op order per output (bias, k ascending, fmaf): sigma of the tiled forward equals the per-thread kernel's bit for bit, rgb to 1e-5."""
    from pagnerf_b200 import ops
    from pagnerf_b200._lib import call, ptr, ptr_array
    gen = torch.Generator().manual_seed(M + IN)
    w = _rand_linear(64, IN, gen) + _rand_linear(16, 64, gen) + _rand_linear(64, 43, gen) + _rand_linear(64, 64, gen) + _rand_linear(3, 64, gen)
    M = (M + S - 1) // S * S
    feats = torch.randn(M, IN, generator=gen).to(DEV)
    lodw = (torch.rand(IN, generator=gen) + 0.5).to(DEV)
    ray_d = torch.nn.functional.normalize(torch.randn(M // S, 3, generator=gen), dim=-1).to(DEV)
    res = []
    for entry in ("pag_decode_dc_fwd", "pag_decode_dc_fwd_tiled"):
        sigma, rgb = torch.full((M,), -7.0, device=DEV), torch.full((M, 3), -7.0, device=DEV)
        call(entry, ptr(feats), ptr(lodw), ptr(ray_d), S, M, IN, ptr_array(w), 64, 27, 1, ptr(sigma), ptr(rgb))
        res.append((sigma, rgb))
    assert torch.equal(res[0][0], res[1][0])
    assert float((res[0][1] - res[1][1]).abs().max()) <= 1e-5      # colour: the tiled kernel builds the view embedding's octaves by angle doubling
    sigma_only = torch.full((M,), -7.0, device=DEV)
    call("pag_decode_dc_fwd_tiled", ptr(feats), None, None, S, M, IN, ptr_array(w), 64, 27, 0, ptr(sigma_only), None)
    call("pag_decode_dc_fwd", ptr(feats), None, ptr(ray_d), S, M, IN, ptr_array(w), 64, 27, 0, ptr(res[0][0]), None)
    assert torch.equal(sigma_only, res[0][0])


@pytest.mark.parametrize("M,IN,Cs,Ci,delta,softmax,temp", [
    (1, 48, 7, 200, True, True, 0.0), (129, 48, 7, 200, True, True, 0.5), (5000, 48, 7, 200, False, True, 0.0),
    (3000, 12, 3, 50, True, False, 2.0), (3000, 12, 3, 51, True, True, 0.0), (777, 48, 0, 200, True, True, 0.0), (777, 28, 16, 0, False, True, 0.0),
    (60000, 48, 7, 208, True, True, 0.0)])
def test_tiled_panoptic_composite_f32(cuda_lib, M, IN, Cs, Ci, delta, softmax, temp):
    """Heads + compositing in one exact-FP32 kernel vs the per-thread FP32 heads ([M,C] probabilities) composited in float64:
    ragged tile tails, rays crossing tile / thread-group boundaries, empty rays, a missing head, no softmax, temperature."""
    from pagnerf_b200 import ops
    gen = torch.Generator().manual_seed(M + Cs + Ci)
    w = (_rand_linear(64, IN, gen) + _rand_linear(max(Cs, 1), 64, gen, 4.0)
         + _rand_linear(64, IN, gen) + _rand_linear(64, 64, gen) + _rand_linear(max(Ci, 1), 64, gen, 4.0))
    feats = torch.randn(M, IN, generator=gen).to(DEV)
    dfeats = torch.randn(M, IN, generator=gen).to(DEV) * 0.3 if delta else None
    lodw = (torch.rand(IN, generator=gen) + 0.5).to(DEV)
    N = max(2, M // 9)
    ridx = torch.sort(torch.randint(0, N, (M,), generator=gen)).values.to(DEV)       # some rays stay empty
    wgt = torch.rand(M, generator=gen).to(DEV)
    alpha = torch.rand(N, 1, generator=gen).to(DEV)
    with torch.no_grad():
        sem, inst = ops.pan_composite_f32(feats, dfeats, lodw, wgt, alpha, ridx, N, Cs, Ci, softmax, softmax, temp, *w)
        ps, pi = ops.DecodePanFn.apply(feats, dfeats, lodw, Cs, Ci, softmax, softmax, temp, False, *w)
    coef = (alpha.reshape(-1)[ridx] * wgt).double()
    for got, p, C in ((sem, ps, Cs), (inst, pi, Ci)):
        if C == 0:
            assert got is None
            continue
        want = torch.zeros(N, C, dtype=torch.float64, device=DEV).index_add_(0, ridx, p.double() * coef[:, None])
        assert_close(got, want.float(), rtol=1e-5, atol_scale=1e-5, msg=f"C={C}")


@pytest.mark.parametrize("name,mode", [("trace_delta_permuto_ray", "ray"), ("trace_delta_permuto_voxel", "voxel"), ("trace_nef_tcnn_ray", "ray")])
def test_tiled_f32_inference_trace_matches_golden(cuda_lib, name, mode):
    """torch.no_grad() + exact-FP32 decoders (the validation / render path, pc_nerf/trainer.py:637-683): the tracer takes the tiled
    kernels; outputs vs the reference-made golden at 1e-4 and vs the stepwise per-thread kernels at 1e-5."""
    from pagnerf_b200 import ops
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden(name)
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    nef = build_cuda_nef(g, DEV)
    nef.decoder_precision = 'fp32'
    tracer = PanopticPackedRFTracer(raymarch_type=mode, num_steps=int(g["num_steps"]),
                                    bg_color='white' if bool(g["bg_white"]) else 'black', ray_max_travel=float(g["ray_max_travel"]))
    rays = Rays(origins=torch.from_numpy(g["o"]).to(DEV), dirs=torch.from_numpy(g["d"]).to(DEV), dist_min=0.0, dist_max=2.0)
    assert not nef.fused_panoptic_ok(set(chans))           # differentiable FP32: the modular kernels (they have a backward)
    res = []
    frac0 = ops.LIVE_COMPACT_FRAC
    try:
        # tiled kernels without / with the live-sample compaction forced on (density pass first, the rest on w != 0 rows), stepwise kernels
        for tiled, frac in ((True, 0.0), (True, 2.0), (False, 0.0)):
            ops.TILED_F32, ops.LIVE_COMPACT_FRAC = tiled, frac
            with torch.no_grad():
                assert nef.fused_panoptic_ok(set(chans)) == tiled
                rb = tracer(nef, channels=chans, rays=rays, lod_idx=None, stage='val')
            res.append({c: getattr(rb, c) for c in chans + ['alpha']})
    finally:
        ops.TILED_F32, ops.LIVE_COMPACT_FRAC = True, frac0
    for c in chans + ['alpha']:
        assert_close(res[0][c], g["out_" + c], msg=c + " vs golden")
        assert_close(res[1][c], g["out_" + c], msg=c + " (compacted) vs golden")
        assert_close(res[0][c], res[2][c], rtol=1e-5, atol_scale=1e-5, msg=c + " vs stepwise")
        assert_close(res[1][c], res[2][c], rtol=1e-5, atol_scale=1e-5, msg=c + " (compacted) vs stepwise")


def test_live_compaction_skips_zero_weight_samples_exactly(cuda_lib):
    """An opaque field (density scaled up until the transmittance underflows to exactly 0 behind the first hits): the no-grad
    render with the zero-weight samples dropped before the colour decoder / delta-grid lookup / heads equals the full render,
    in both decoder precisions, and really runs on fewer samples."""
    from pagnerf_b200 import ops
    from pagnerf_b200.tracers import PanopticPackedRFTracer
    from pagnerf_b200.wisp_compat import Rays
    g = load_golden("trace_delta_permuto_ray")
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    nef = build_cuda_nef(g, DEV)
    with torch.no_grad():
        nef.decoder_density.lout.bias[0] = 3000.0
    tracer = PanopticPackedRFTracer(raymarch_type='ray', num_steps=int(g["num_steps"]), bg_color='white')
    rays = Rays(origins=torch.from_numpy(g["o"]).to(DEV), dirs=torch.from_numpy(g["d"]).to(DEV), dist_min=0.0, dist_max=2.0)
    frac0 = ops.LIVE_COMPACT_FRAC
    seen = []
    orig = ops.pan_composite_f32

    def spy(feats, *a, **k):
        seen.append(feats.shape[0])
        return orig(feats, *a, **k)
    try:
        for prec in ('fp32', 'fp16'):
            nef.decoder_precision = prec
            res = []
            for frac in (0.0, 0.7):
                ops.LIVE_COMPACT_FRAC = frac
                ops.pan_composite_f32 = spy
                with torch.no_grad():
                    rb = tracer(nef, channels=chans, rays=rays, lod_idx=None, stage='val')
                res.append({c: getattr(rb, c) for c in chans + ['alpha']})
            for c in chans + ['alpha']:
                assert_close(res[1][c], res[0][c], rtol=1e-5, atol_scale=1e-5, msg=f"{prec} {c}")
            if prec == 'fp32':
                assert len(seen) == 2 and 0 < seen[1] < 0.5 * seen[0], seen
    finally:
        ops.LIVE_COMPACT_FRAC, ops.pan_composite_f32 = frac0, orig
