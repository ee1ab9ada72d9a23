from torchrecsys_b200.collaborative.mlp import MLP  # noqa: F401
