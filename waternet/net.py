"""Drop-in for the reference's ``waternet/net.py``: same names, B200 kernels underneath."""
from waternet_b200.net import ConfidenceMapGenerator, Refiner, WaterNet  # noqa: F401
