"""Device-side losses with the reference's interfaces (loss/*.py)."""
from .lin_assignment_things import LinAssignmentThingsLoss

__all__ = ["LinAssignmentThingsLoss"]
