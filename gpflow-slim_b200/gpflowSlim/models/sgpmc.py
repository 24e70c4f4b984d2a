"""`gpflowSlim.models.sgpmc.SGPMC` of the reference lives in models/whitened.py here."""
from .whitened import SGPMC  # noqa: F401
