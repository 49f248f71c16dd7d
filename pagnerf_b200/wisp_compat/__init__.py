"""Stand-ins for the kaolin-wisp v0.1.1 *types* the reference's plugin classes inherit from.

The reference (README.md:42) builds its grids / neural fields / tracers on kaolin-wisp base
classes.  When real wisp is importable we re-export its classes so our plugins drop into the
real trainer / app unchanged; otherwise these pure-Python restatements (behaviour as recalled in
SURVEY.md Appendix A.8) provide the same constructor / dispatch contracts.  No arithmetic of the
hot path lives here -- only containers and argument plumbing.
"""
try:  # pragma: no cover - real wisp is not installed in the build image
    from wisp.core import Rays, RenderBuffer                      # noqa: F401
    from wisp.models.nefs import BaseNeuralField                  # noqa: F401
    from wisp.tracers import BaseTracer, PackedRFTracer           # noqa: F401
    from wisp.models.decoders import BasicDecoder                 # noqa: F401
    from wisp.models.embedders import PositionalEmbedder, get_positional_embedder  # noqa: F401
    from wisp.models.activations import get_activation_class      # noqa: F401
    from wisp.models.layers import get_layer_class                # noqa: F401
    from wisp.utils import PerfTimer                              # noqa: F401
    from wisp.models.pipeline import Pipeline                     # noqa: F401
    HAVE_WISP = True
except Exception:  # ImportError or a broken partial install
    from .core import Rays, RenderBuffer                          # noqa: F401
    from .nefs import BaseNeuralField                             # noqa: F401
    from .tracers import BaseTracer, PackedRFTracer               # noqa: F401
    from .modules import (BasicDecoder, PositionalEmbedder, get_positional_embedder,  # noqa: F401
                          get_activation_class, get_layer_class, PerfTimer, Pipeline)
    HAVE_WISP = False
