"""Drop-in for the reference's ``waternet/training_utils.py``."""
from waternet_b200.training_utils import FlipRotate, SyntheticUIEB, UIEBDataset, arr2ten, ten2arr  # noqa: F401
