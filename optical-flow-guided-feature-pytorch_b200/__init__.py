"""B200-native OFF (Optical Flow guided Feature) unit and sub-network.

Drop-in for the OFF path of JoeHEZHAO/Optical-Flow-Guided-Feature-Pytorch
(RGB_OFF.py:596-860, Flow_OFF.py:606-884, basic_ops.py, util.py Sobel filters):
Python/PyTorch host code calling hand-written sm_100a kernels through the C ABI
in include/offk.h (liboffk.so).  No CPU fallback.
"""
__version__ = "0.1.0"
