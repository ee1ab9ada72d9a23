import numpy as np  # noqa: F401
import torch  # noqa: F401
from torchrecsys_b200.evaluate.metrics import Metrics  # noqa: F401
