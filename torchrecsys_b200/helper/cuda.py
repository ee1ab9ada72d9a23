"""Device placement shims of the reference (helper/cuda.py:3-16)."""


def gpu(tensor, gpu=False):
    return tensor.cuda() if gpu else tensor


def cpu(tensor):
    return tensor.cpu() if tensor.is_cuda else tensor
