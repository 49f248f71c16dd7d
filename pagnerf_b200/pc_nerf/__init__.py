"""Neural-field plugin classes with the reference's surface (pc_nerf/panoptic_{,delta_,dd_}nef.py)."""
from .panoptic_nef import PanopticNeF
from .panoptic_delta_nef import PanopticDeltaNeF
from .panoptic_dd_nef import PanopticDDensityNeF

__all__ = ["PanopticNeF", "PanopticDeltaNeF", "PanopticDDensityNeF"]
from .ba_pipeline import BAPipeline  # noqa: F401,E402
