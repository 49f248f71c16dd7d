"""Tracer plugin classes with the reference's surface (tracers/*.py)."""
from .panoptic_packed_rf_tracer import PanopticPackedRFTracer
from .panoptic_dd_packed_rf_tracer import PanopticDDensityPackedRFTracer

__all__ = ["PanopticPackedRFTracer", "PanopticDDensityPackedRFTracer"]
