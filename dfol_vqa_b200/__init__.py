"""dfol_vqa_b200 -- B200-native differentiable-FOL reasoning path (visual oracle + batched interpreter).

Host side is Python/PyTorch (device memory, streams, torch.distributed); all compute runs in hand-written
sm_100a CUDA kernels behind the C ABI declared in include/dfol_b200.h (libdfol_b200.so, loaded with ctypes).
There is no CPU fallback: importing the compute modules without the built library raises.
"""

__version__ = '0.1.0'
