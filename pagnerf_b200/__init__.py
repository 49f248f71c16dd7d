"""pagnerf_b200 -- B200-native (sm_100a) implementation of PAg-NeRF's per-ray hot path.

Plugin surface (same names / signatures as the reference tree):
    pagnerf_b200.grids      Occtree, PermutoGrid, HashGridTinyCudaNN, HashGridTorch   (.raymarch / .interpolate)
    pagnerf_b200.pc_nerf    PanopticNeF, PanopticDeltaNeF                             (nef(channels=..., coords=..., ray_d=...))
    pagnerf_b200.tracers    PanopticPackedRFTracer                                    (.trace / tracer(nef, rays=...))
All arithmetic runs in csrc/*.cu through the C ABI of include/pagnerf_b200.h; importing the package
is cheap, the first op call loads the library and raises if it has not been built.
"""
__version__ = "0.1.0"
