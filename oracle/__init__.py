"""CPU oracle for the PAg-NeRF per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy for integer / bit-exact work, torch-CPU
for the differentiable float work), the algorithms the reference hot path runs:

  * occupancy-octree build / query / ray-trace / ray-march   (oracle.spc, oracle.raymarch)
  * permutohedral-lattice encoding                           (oracle.permuto)
  * instant-ngp hash grids, tcnn flavour and HashNeRF flavour (oracle.hashgrid)
  * BasicDecoder MLPs + positional embedder                  (oracle.decoders)
  * packed exponential integration / segmented sums          (oracle.spc)
  * PanopticNeF / PanopticDeltaNeF forward, PanopticPackedRFTracer.trace
                                                             (oracle.nef, oracle.tracer)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package ``pagnerf_b200``
never does: it fails loudly when its CUDA library is missing.

PARITY STATUS:
  * ``hashgrid.HashEmbedderOracle`` is PINNED: checked bit-for-bit / to 1e-6
    against the reference's own ``grids/hash_grid_torch.py`` imported verbatim
    (tests/golden/make_golden.py, tests/golden/hash_torch_*.npz).
  * the tracer / neural-field glue is PINNED to the reference source text: the
    goldens in tests/golden/trace_*.npz were produced by running the unmodified
    ``tracers/panoptic_packed_rf_tracer.py`` and ``pc_nerf/panoptic_{,delta_}nef.py``
    on top of stub wisp/kaolin modules backed by this oracle.
  * kaolin / kaolin-wisp / permutohedral_encoding / tiny-cuda-nn arithmetic is
    **parity unpinned**: those packages are absent from /root/reference and from
    this image; the restatements follow their published algorithms as recalled
    in SURVEY.md Appendix A and are the *definition* our CUDA path is held to.
"""
