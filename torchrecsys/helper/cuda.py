from torchrecsys_b200.helper.cuda import cpu, gpu  # noqa: F401
