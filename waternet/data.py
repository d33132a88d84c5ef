"""Drop-in for the reference's ``waternet/data.py``: same names, B200 kernels underneath."""
from waternet_b200.data import gamma_correction, histeq, transform, white_balance_transform  # noqa: F401
