"""Tracer plugin classes with the reference's surface (tracers/*.py)."""
from .panoptic_packed_rf_tracer import PanopticPackedRFTracer

__all__ = ["PanopticPackedRFTracer"]
