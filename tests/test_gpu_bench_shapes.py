"""Oracle parity AT THE BENCHMARKED SHAPES (BASELINE config 2: L = 24 levels, T = 2^18 x 2 grids, C_sem = 6,
C_inst = 200 -> the N = 208 MMA, pruned level-7 octree, 128 march steps) for the path bench.py times: the sync-free fused
training trace under autocast (fp16-rounded coordinates, fp16 operand images, tcgen05 decoders, fused heads + compositing).

Three computations of the same step on the same rays and the same jitter stream:
  A  ours            ops.FusedTraceFn through PanopticPackedRFTracer under torch.autocast
  B  oracle, exact   oracle.field.trace_oracle in fp32 on the SAME fp16-rounded coordinates (pos_half)
  C  oracle, amp     the same oracle with oracle/autocast.py: the reference's own autocast numerics (fp16 nn.Linear with
                     fp32 accumulation, fp16 ReLU / sigmoid, fp32 softmax, loss scale 2^16) restated on the CPU
Tolerances: outputs A vs B at north_star's fp16 tolerance 2e-3; every gradient A vs B in relative l2 at
max(2e-3, 1.5 x the error C makes against B) -- i.e. the tensor-core path may be no worse than 1.5 x the reference's own
autocast step, measured here rather than assumed; plus an elementwise check with a floor of the same size.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
N_RAYS = 1024
SEED_JITTER = 5


def _oracle_field(nef):
    """FieldOracle carrying the parameters of a bench.Workload field."""
    from oracle.field import FieldOracle
    from oracle.permuto import PermutoEncodingOracle
    import bench

    def enc(grid):
        e = grid.embedder
        o = PermutoEncodingOracle(e.capacity, e.nr_levels, e.nr_feat, np.geomspace(1.0, 1e-4, e.nr_levels))
        o.load_state_dict({k: v.detach().cpu() for k, v in e.state_dict().items()})
        return o

    dd = hasattr(nef, 'decoder_delta_density')
    f = FieldOracle(enc(nef.grid), enc(nef.delta_grid), feat_dim=bench.L * bench.F, num_classes=bench.C_SEM,
                    num_instances=bench.C_INST, delta_density=dd)
    for name in ('decoder_density', 'decoder_color', 'decoder_semantics', 'decoder_inst') + (('decoder_delta_density',) if dd else ()):
        getattr(f, name).load_state_dict({k: v.detach().cpu() for k, v in getattr(nef, name).state_dict().items()})
    return f


def _rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def _oracle_step(field, o, d, march, gws, chans, amp, dd=False):
    """-> (outputs, {oracle parameter name: grad}, grad origins, grad dirs) of the oracle trace; amp: reference autocast numerics."""
    from oracle import autocast
    from oracle.field import trace_oracle
    ridx, samples, depths, deltas, boundary = march
    # GradScaler semantics (pc_nerf/trainer.py:582 scaler.scale(loss).backward(); scaler.step skips non-finite steps and halves
    # the scale): start from its initial 2^16 and halve until the fp16 backward no longer overflows
    scale = 65536.0 if amp else 1.0
    while True:
        for p in field.parameters():
            p.grad = None
        ot, dt = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
        # sample positions stay attached to the rays (pose optimisation): x = o + t d with the marcher's own depths
        s = ot[ridx][:, None, :] + dt[ridx][:, None, :] * depths.reshape(-1, 1, 1)
        s = samples + (s - s.detach())
        with autocast.emulate_fp16(amp):
            field.pos_half = True
            out = trace_oracle(field, ot, dt, ridx, s, depths, deltas, boundary, chans, dd=dd)
        loss = sum((out[c].float() * gws[c]).sum() for c in chans)
        (loss * scale).backward()
        if all(torch.isfinite(p.grad).all() for p in field.parameters() if p.grad is not None):
            break
        scale *= 0.5
        assert scale >= 1.0
    grads = {k: (p.grad / scale) for k, p in field.named_parameters() if p.grad is not None}
    return {c: out[c].detach().float() for c in chans + ['alpha', 'hit']}, grads, ot.grad / scale, dt.grad / scale


NAME_MAP = {'grid.embedder.lattice_values': 'grid.lattice_values', 'delta_grid.embedder.lattice_values': 'delta_grid.lattice_values'}


@pytest.mark.parametrize("dd", [False, True])
def test_bench_field_fused_trace_vs_oracle(cuda_lib, dd):
    import bench
    from oracle import spc as ospc, raymarch as orm
    from pagnerf_b200.wisp_compat import Rays
    dev = torch.device(DEV)
    wl = bench.Workload(dev, n_rays=N_RAYS, seed=0, n_batches=1, dd=dd)
    for g in (wl.nef.grid, wl.nef.delta_grid):
        g.blas.fixed_jitter, g.blas.jitter_seed = True, SEED_JITTER
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    o_np, d_np = bench.make_rays(N_RAYS, 0, 0)
    gen = torch.Generator().manual_seed(1)
    gws = {'rgb': torch.randn(N_RAYS, 3, generator=gen), 'depth': torch.randn(N_RAYS, 1, generator=gen),
           'semantics': torch.randn(N_RAYS, bench.C_SEM, generator=gen), 'inst_embedding': torch.randn(N_RAYS, bench.C_INST, generator=gen)}
    # ---- A: the benchmarked path --------------------------------------------------------------------------------------
    o = torch.from_numpy(o_np).to(dev).requires_grad_(True)
    d = torch.from_numpy(d_np).to(dev).requires_grad_(True)
    with torch.autocast('cuda', dtype=torch.float16):
        rb = wl.tracer(wl.nef, channels=chans, rays=Rays(origins=o, dirs=d, dist_min=bench.NEAR, dist_max=bench.FAR), lod_idx=None, stage='train')
        assert torch.is_tensor(wl.tracer.last_num_samples), "the fused sync-free trace must be the path taken"
        loss = sum((getattr(rb, c).float() * gws[c].to(dev)).sum() for c in chans)
    loss.backward()
    ours = {c: getattr(rb, c).detach().float().cpu() for c in chans + ['alpha']}
    ours_g = {k: p.grad.detach().cpu() for k, p in wl.nef.named_parameters() if p.grad is not None}
    # ---- oracle march: same octree, same counter-based jitter stream --------------------------------------------------
    octree = ospc.points_to_octree(bench.make_scene(bench.LEVEL, 0), bench.LEVEL)
    _, _, prefix = ospc.scan_octree(octree, bench.LEVEL)
    ridx, pidx, s, dp, dl, b = orm.raymarch_ray(octree, prefix, o_np, d_np, bench.LEVEL, bench.NUM_STEPS, bench.NEAR, bench.FAR, seed=SEED_JITTER)
    assert int(wl.tracer.last_num_samples.item()) == ridx.shape[0], "packed-sample count: bit-field marcher == oracle marcher"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    march = (t(ridx).long(), t(s), t(dp), t(dl), t(b))
    field = _oracle_field(wl.nef)
    ot, dt = torch.from_numpy(o_np), torch.from_numpy(d_np)
    ref, ref_g, ref_go, ref_gd = _oracle_step(field, ot, dt, march, gws, chans, amp=False, dd=dd)
    amp, amp_g, amp_go, amp_gd = _oracle_step(field, ot, dt, march, gws, chans, amp=True, dd=dd)
    assert np.array_equal(rb.hit.cpu().numpy(), ref['hit'].numpy().astype(bool))
    # ---- outputs: north_star fp16 tolerance --------------------------------------------------------------------------
    report = []
    for c in chans + ['alpha']:
        a, r = ours[c], ref[c]
        tol = 2e-3 * r.abs() + 2e-3 * float(r.abs().max())
        bad = (a - r).abs() > tol
        e_ours, e_amp = _rel_l2(a, r), _rel_l2(amp[c], r)
        report.append(f"out {c}: ours {e_ours:.2e} reference-autocast {e_amp:.2e}")
        assert not bad.any(), f"{c}: {int(bad.sum())}/{bad.numel()} beyond 2e-3 (max err {float((a - r).abs().max()):.3e}, ref max {float(r.abs().max()):.3e})"
        assert e_ours <= max(2e-3, 1.5 * e_amp), f"{c}: rel l2 {e_ours:.3e} vs reference autocast {e_amp:.3e}"
    # ---- every gradient: tolerance set from the reference's own autocast error ---------------------------------------
    checked = 0
    for k, g in ours_g.items():
        ko = NAME_MAP.get(k, k)
        if ko not in ref_g:
            continue
        r, am = ref_g[ko].float(), amp_g[ko].float()
        e_ours, e_amp = _rel_l2(g, r), _rel_l2(am, r)
        report.append(f"grad {k}: ours {e_ours:.2e} reference-autocast {e_amp:.2e}")
        lim = max(2e-3, 1.5 * e_amp)
        assert torch.isfinite(g).all(), k
        assert e_ours <= lim, f"grad {k}: rel l2 {e_ours:.3e} > {lim:.3e} (reference autocast {e_amp:.3e})"
        # elementwise, with a floor of the same relative size on the tensor's scale
        floor = lim * float(r.abs().max()) * 4.0
        bad = (g.float() - r).abs() > lim * r.abs() + floor
        assert not bad.any(), f"grad {k}: {int(bad.sum())}/{bad.numel()} elements beyond tolerance"
        checked += 1
    assert checked >= 22, f"only {checked} parameter gradients compared"
    for name, g, r, am in (("origins", o.grad.cpu(), ref_go, amp_go), ("dirs", d.grad.cpu(), ref_gd, amp_gd)):
        e_ours, e_amp = _rel_l2(g, r), _rel_l2(am, r)
        report.append(f"grad {name}: ours {e_ours:.2e} reference-autocast {e_amp:.2e}")
        assert e_ours <= max(2e-3, 1.5 * e_amp), f"grad {name}: rel l2 {e_ours:.3e} vs reference autocast {e_amp:.3e}"
    print("\n".join(report))


def test_permuto_half_coords_bit_exact(cuda_lib):
    """Autocast quirk (grids/permuto_grid.py:65,71): coordinates rounded to fp16 before the lattice.  At the benchmarked
    lattice (L = 24, T = 2^18, finest scale 1e-4) the kernel's internal rounding (pos_half) must give the oracle's
    rem0 / rank / idx on the rounded positions bit for bit, and exactly the features of pre-rounded input."""
    from oracle.permuto import PermutoEncodingOracle
    from pagnerf_b200 import ops
    from pagnerf_b200._lib import call, ptr
    cap, L = 2 ** 18, 24
    enc = PermutoEncodingOracle(cap, L, 2, np.geomspace(1.0, 1e-4, L), seed=3)
    with torch.no_grad():
        enc.lattice_values.mul_(1e4)
    rng = np.random.default_rng(7)
    x = torch.from_numpy(rng.uniform(-1, 1, (4099, 3)).astype(np.float32))
    xr = x.half().float()
    assert not torch.equal(x, xr)
    rem0, rank, idx = enc.indices(x, pos_half=True)
    rem0_u, _, idx_u = enc.indices(x)
    assert (idx != idx_u).any(), "rounding must matter at the fine levels, or this test pins nothing"
    sf, sh, an = enc.scale_factor.to(DEV), enc.random_shift_per_level.to(DEV), enc.anneal_window.to(DEV)
    gi, gr, _ = ops.permuto_indices(xr.to(DEV), cap, sf, sh)
    assert np.array_equal(gi.cpu().numpy().view(np.uint32), idx), "lattice vertex indices on fp16-rounded coordinates"
    assert np.array_equal(gr.cpu().numpy(), rank)
    tb = enc.lattice_values.detach().to(DEV).contiguous()
    M = x.shape[0]
    m_dev = torch.tensor([M], dtype=torch.int64, device=DEV)
    out_h = torch.empty(M, 2 * L, device=DEV)
    out_r = torch.empty(M, 2 * L, device=DEV)
    call("pag_permuto_fwd_dyn", ptr(x.to(DEV)), M, ptr(m_dev), 1, ptr(tb), cap, L, 2, ptr(sf), ptr(sh), ptr(an), ptr(out_h))
    call("pag_permuto_fwd_dyn", ptr(xr.to(DEV)), M, ptr(m_dev), 0, ptr(tb), cap, L, 2, ptr(sf), ptr(sh), ptr(an), ptr(out_r))
    assert torch.equal(out_h, out_r), "in-kernel fp16 rounding == pre-rounded coordinates, bit for bit"
    ref = enc(x, pos_half=True).detach()
    err = (out_h.cpu() - ref).abs()
    assert float((err / (1e-4 * ref.abs() + 1e-4 * ref.abs().max())).max()) <= 1.0, "features on rounded coordinates vs oracle"


def test_bench_field_fp32_inference_vs_oracle(cuda_lib):
    """BASELINE config 5's path at its shapes (L = 24, C_inst = 200): torch.no_grad(), no autocast -> fp32 coordinates, the
    register-tiled exact-FP32 decoders with the heads composited in-kernel (csrc/decoder_tiled.cu), against the exact oracle on
    the same marched samples.  Tolerance: north_star's fp32 1e-4."""
    import bench
    from oracle import spc as ospc, raymarch as orm
    from oracle.field import trace_oracle
    from pagnerf_b200.wisp_compat import Rays
    dev = torch.device(DEV)
    wl = bench.Workload(dev, n_rays=N_RAYS, seed=0, n_batches=1, config=5)
    for g in (wl.nef.grid, wl.nef.delta_grid):
        g.blas.fixed_jitter, g.blas.jitter_seed = True, SEED_JITTER
    chans = ['rgb', 'depth', 'semantics', 'inst_embedding']
    o_np, d_np = bench.make_rays(N_RAYS, 0, 0)
    wl.nef.decoder_precision = 'auto'
    with torch.no_grad():
        assert wl.nef.fused_panoptic_ok(set(chans)) and not wl.nef._use_tc(), "exact-FP32 tiled kernels must be the path taken"
        rb = wl.tracer(wl.nef, channels=chans, rays=Rays(origins=torch.from_numpy(o_np).to(dev), dirs=torch.from_numpy(d_np).to(dev),
                                                          dist_min=bench.NEAR, dist_max=bench.FAR), lod_idx=None, stage='val')
    ours = {c: getattr(rb, c).float().cpu() for c in chans + ['alpha']}
    octree = ospc.points_to_octree(bench.make_scene(bench.LEVEL, 0), bench.LEVEL)
    _, _, prefix = ospc.scan_octree(octree, bench.LEVEL)
    ridx, pidx, s, dp, dl, b = orm.raymarch_ray(octree, prefix, o_np, d_np, bench.LEVEL, bench.NUM_STEPS, bench.NEAR, bench.FAR, seed=SEED_JITTER)
    assert int(wl.tracer.last_num_samples) == ridx.shape[0], "packed-sample count: marcher == oracle marcher"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    field = _oracle_field(wl.nef)
    field.pos_half = False
    with torch.no_grad():
        ref = trace_oracle(field, torch.from_numpy(o_np), torch.from_numpy(d_np), t(ridx).long(), t(s), t(dp), t(dl), t(b), chans)
    assert np.array_equal(rb.hit.cpu().numpy(), ref['hit'].numpy().astype(bool))
    for c in chans + ['alpha']:
        a, r = ours[c], ref[c].float()
        bad = (a - r).abs() > 1e-4 * r.abs() + 1e-4 * float(r.abs().max())
        print(f"out {c}: rel l2 {_rel_l2(a, r):.2e}, max abs err {float((a - r).abs().max()):.2e}")
        assert not bad.any(), f"{c}: {int(bad.sum())}/{bad.numel()} beyond 1e-4 (max err {float((a - r).abs().max()):.3e}, ref max {float(r.abs().max()):.3e})"
