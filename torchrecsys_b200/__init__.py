"""torchrecsys_b200 -- B200-native (sm_100a) hot path of TorchRecSys behind the reference API.

Host side mirrors FrancescoI/torchrecsys (`TorchRecSys`, `ProcessData`, `FastDataLoader`, the
`Linear` / `FM` / `MLP` scorer modules, `hinge_loss`, `Metrics`); all arithmetic on the path runs
in hand-written CUDA kernels behind the C ABI of include/trs.h (libtrs_b200.so)."""

__version__ = "0.1.0"
