"""``Image.inference(inferencer)`` alias kept from the reference
(chunkflow/chunk/image/base.py:22-24)."""
from chunkflow_b200.chunk.base import Chunk


class Image(Chunk):
    def inference(self, inferencer):
        return inferencer(self)
