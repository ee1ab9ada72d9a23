"""Drop-in alias: the reference's import paths (``torchrecsys.model.TorchRecSys`` ...) resolve to
torchrecsys_b200, so code and tests written against FrancescoI/torchrecsys run unchanged."""
