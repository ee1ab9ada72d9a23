from torchrecsys_b200.collaborative.linear import Linear  # noqa: F401
