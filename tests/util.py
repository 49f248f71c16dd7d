"""Shared builders for the parity tests (oracle side + CUDA side from the same golden parameters)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GOLDEN_CFG = dict(L=6, F=2, hidden=64, num_classes=6, num_instances=20, view_multires=4,
                  coarsest_scale=1.0, finest_scale=0.01, capacity_log_2=10, delta_capacity_log_2=9,
                  codebook_bitwidth=10, base_resolution=16)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_params(g):
    return {k[len("param:"):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith("param:")}


def golden_grads(g):
    return {k[len("grad:"):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith("grad:")}


def _load_decoder(dec, params, name):
    sd = {k[len(name) + 1:]: v for k, v in params.items() if k.startswith(name + ".")}
    dec.load_state_dict(sd)


def build_oracle_field(g):
    """FieldOracle carrying the golden's parameters (permuto+delta or tcnn, decided by the keys present)."""
    from oracle.field import FieldOracle
    from oracle.permuto import PermutoEncodingOracle
    from oracle.hashgrid import TcnnHashGridOracle
    c = GOLDEN_CFG
    p = golden_params(g)

    def permuto(prefix):
        cap = p[prefix + ".embedder.lattice_values"].shape[1]
        enc = PermutoEncodingOracle(cap, c["L"], c["F"], np.geomspace(c["coarsest_scale"], c["finest_scale"], c["L"]))
        enc.load_state_dict({k[len(prefix) + 10:]: v for k, v in p.items() if k.startswith(prefix + ".embedder.")})
        return enc

    if "grid.embedder.lattice_values" in p:
        grid, delta = permuto("grid"), (permuto("delta_grid") if "delta_grid.embedder.lattice_values" in p else None)
    else:
        grid = TcnnHashGridOracle(c["L"], c["F"], c["codebook_bitwidth"], c["base_resolution"], 2.0, out_half=True)
        grid.load_state_dict({"params": p["grid.embedder.params"]})
        delta = None
    dd = any(k.startswith("decoder_delta_density.") for k in p)
    f = FieldOracle(grid, delta, feat_dim=c["L"] * c["F"], hidden_dim=c["hidden"], num_classes=c["num_classes"],
                    num_instances=c["num_instances"], view_multires=c["view_multires"], delta_density=dd)
    for name in ("decoder_density", "decoder_color", "decoder_semantics", "decoder_inst") + (("decoder_delta_density",) if dd else ()):
        _load_decoder(getattr(f, name), p, name)
    return f


def oracle_march(g, raymarch_type):
    """Run the oracle marcher (+ max-travel filter) on the golden's rays; returns torch tensors."""
    from oracle import spc as ospc, raymarch as orm
    level, S = int(g["level"]), int(g["num_steps"])
    pts, pyr, pre = ospc.scan_octree(g["octree"], level)
    if raymarch_type == 'ray':
        ridx, pidx, s, dp, dl, b = orm.raymarch_ray(g["octree"], pre, g["o"], g["d"], level, S, 0.0, 2.0, seed=int(g["jitter_seed"]))
    else:
        ridx, pidx, s, dp, dl, b = orm.raymarch_voxel(g["octree"], pts, pyr, pre, g["o"], g["d"], level, S, seed=int(g["jitter_seed"]))
        keep = orm.max_travel_filter(ridx, dp, float(g["ray_max_travel"]))
        dl = dl.reshape(dp.shape)[keep].reshape(-1, 1)
        b = b.reshape(dp.shape[:2])[keep].reshape(-1)
        ridx, pidx, s, dp = ridx[keep], pidx[keep], s[keep], dp[keep]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return t(ridx).long(), t(pidx).long(), t(s), t(dp), t(dl), t(b)


def build_cuda_nef(g, device):
    """Our plugin classes, loaded from the golden's state_dict (same keys as the reference's nef)."""
    from pagnerf_b200.pc_nerf import PanopticDeltaNeF, PanopticNeF, PanopticDDensityNeF
    c = GOLDEN_CFG
    p = golden_params(g)
    permuto = "grid.embedder.lattice_values" in p
    delta = "delta_grid.embedder.lattice_values" in p
    dd = any(k.startswith("decoder_delta_density.") for k in p)
    kw = dict(grid_type="PermutoGrid" if permuto else "HashGridTinyCudaNN", interpolation_type='linear',
              multiscale_type='cat', feature_dim=c["F"], num_lods=c["L"], base_lod=2, hidden_dim=c["hidden"],
              num_layers=1, view_multires=c["view_multires"], pos_multires=4, embedder_type='positional',
              activation_type='relu', layer_type='none', num_classes=c["num_classes"], num_instances=c["num_instances"],
              sem_num_layers=1, sem_hidden_dim=64, inst_num_layers=2, inst_hidden_dim=64, sem_softmax=True,
              inst_softmax=True, sem_detach=True, inst_detach=True,
              panoptic_features_type='delta' if delta else None, blas_level=int(g["level"]),
              coarsest_scale=c["coarsest_scale"], finest_scale=c["finest_scale"], capacity_log_2=c["capacity_log_2"],
              delta_capacity_log_2=c["delta_capacity_log_2"], codebook_bitwidth=c["codebook_bitwidth"])
    nef = (PanopticDDensityNeF if dd else (PanopticDeltaNeF if delta else PanopticNeF))(**kw)
    grids = [nef.grid] + ([nef.delta_grid] if delta else [])
    for gr in grids:
        if permuto:
            gr.init_from_scales()
        else:
            gr.init_from_resolutions([c["base_resolution"] * 2 ** i for i in range(c["L"])])
        gr.blas_init(torch.from_numpy(g["octree"].copy()))
        gr.blas.fixed_jitter = True
        gr.blas.jitter_seed = int(g["jitter_seed"])
    missing, unexpected = nef.load_state_dict(p, strict=False)
    missing = [k for k in missing if '.blas_' not in k]   # the reference's tcnn grid does not checkpoint its octree
    assert not missing, missing
    return nef.to(device)


def assert_close(a, b, rtol=1e-4, atol_scale=2e-5, msg=""):
    """Elementwise |a-b| <= rtol*|b| + atol_scale*max|b|  (north_star: 1e-4 relative in fp32).  The floor (default 2e-5 of the
    tensor's scale, ~170 fp32 ulps of the largest element) covers the elements that are differences / sums of much larger terms
    -- table gradients accumulated from thousands of atomics in arbitrary order, softmax tails -- whose absolute error is set by
    the large terms, not by their own magnitude; a fifth of the relative tolerance, so it cannot hide a wrong small element
    behind the relative term of a large one."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, f"{msg}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    tol = rtol * b.abs() + atol_scale * (b.abs().max() if b.numel() else 0.0)
    bad = (a - b).abs() > tol
    assert not bad.any(), (f"{msg}: {int(bad.sum())}/{bad.numel()} mismatches, max abs err "
                           f"{float((a - b).abs().max()):.3e}, ref max {float(b.abs().max()):.3e}")


def assert_close_norm(a, b, rel_l2=1.5e-2, max_frac=0.1, msg=""):
    """Reduced-precision comparison: ||a-b||_2 <= rel_l2*||b||_2 and max|a-b| <= max_frac*max|b|.
    Gradients of a ReLU MLP evaluated with fp16-rounded activations differ from the fp32 ones by the few hidden units whose
    pre-activation sign flips under the rounding.  The size of that effect is MEASURED, not assumed: tests/test_gpu_bench_shapes.py
    runs the reference's own autocast numerics (oracle/autocast.py) beside the tensor-core path at the benchmarked shapes -- the
    reference's autocast step is 2e-4 .. 8e-3 relative l2 away from exact fp32 per tensor, the tensor-core kernels 2e-5 .. 6e-3
    (profiles/r02_parity_bench_shapes.md).  The default here (1.5e-2 l2, 10 % max) is for the small goldens (48 rays, a few
    hundred samples), where one flipped unit is a visibly larger share of the total."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, f"{msg}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.isfinite(a).all(), f"{msg}: non-finite values"
    nb = float(b.norm())
    err = float((a - b).norm())
    mx = float((a - b).abs().max()) if a.numel() else 0.0
    assert err <= rel_l2 * nb + 1e-30 and mx <= max_frac * float(b.abs().max()) + 1e-30, (
        f"{msg}: rel l2 err {err / max(nb, 1e-30):.3e} (limit {rel_l2}), max abs err {mx:.3e} vs ref max {float(b.abs().max()):.3e}")
