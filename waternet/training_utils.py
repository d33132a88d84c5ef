"""Drop-in for the reference's ``waternet/training_utils.py``."""
from waternet_b200.training_utils import (FlipRotate, GpuBatchLoader, SyntheticUIEB, UIEBDataset,  # noqa: F401
                                            arr2ten, ten2arr)
