from torchrecsys_b200.collaborative.fm import FM  # noqa: F401
