from torchrecsys_b200.model import *  # noqa: F401,F403
from torchrecsys_b200.model import TorchRecSys  # noqa: F401
