from torchrecsys_b200.dataset.dataset import Data, FastDataLoader, ProcessData  # noqa: F401
