from .lib import CholeskyError, load, handle_for  # noqa: F401
