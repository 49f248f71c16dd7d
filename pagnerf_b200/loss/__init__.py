"""Device-side losses with the reference's interfaces (loss/*.py)."""
from .lin_assignment_things import LinAssignmentThingsLoss
from .panoptic_loss import panoptic_loss

__all__ = ["LinAssignmentThingsLoss", "panoptic_loss"]
