from .base import Chunk  # noqa: F401
