"""tda_multimodal_b200 -- B200-native (sm_100a) implementation of the per-layer TDA hot path of
Princeton-Applied-Geometry-Topology/tda-multimodal: pairwise distances -> UMAP stages -> Vietoris-Rips
persistence, behind the reference's own call signatures (see shims/ and INTEGRATION.md).

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic of the path runs in
libtda_b200.so (hand-written CUDA, C ABI in include/tda_b200.h).  There is no CPU fallback: importing the
compute modules without the built library, or calling them without a GPU, raises.
"""
__version__ = "0.1.0"
