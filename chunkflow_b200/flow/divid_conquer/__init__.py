from .inferencer import Inferencer  # noqa: F401
