from torchrecsys_b200.helper.loss import hinge_loss  # noqa: F401
