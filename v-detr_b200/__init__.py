"""vdetr_b200 -- Blackwell-native (sm_100a) implementation of V-DETR's Vertex-RPE decoder hot path.

Layout:
  csrc/               CUDA kernels + the C ABI (include/vdetr_b200.h) -> lib/libvdetr_b200.so
  _C.py               ctypes binding of the C ABI (no torch types cross the boundary: raw pointers + stream)
  ops.py              torch.autograd.Function wrappers around the kernels
  pointnet2_utils.py  drop-in for third_party/pointnet2/pointnet2_utils.py (FPS, gather, ball query, grouping)
  vdetr_transformer.py, helpers.py, model_vdetr.py   drop-ins for the reference's models/*.py

There is no CPU fallback: every op raises if the CUDA library or a CUDA tensor is missing.
"""
__version__ = "0.1"
