"""`gpflowSlim.models.gpmc.GPMC` of the reference lives in models/whitened.py here."""
from .whitened import GPMC  # noqa: F401
