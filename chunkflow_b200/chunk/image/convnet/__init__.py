"""Import path named by BASELINE.json's north_star
(``chunkflow.chunk.image.convnet.Inferencer``); in the surveyed checkout the class
lives in ``flow/divid_conquer/inferencer.py`` -- both paths work here."""
from chunkflow_b200.flow.divid_conquer.inferencer import Inferencer  # noqa: F401
