"""B200-native hot path of the adaptive-depth U-Net (SR + segmentation).

Import this package as ``b200unet`` (the repo-root shim ``b200unet.py`` maps that
name onto this directory, whose mandated name is not a Python identifier).

Layout:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/b200_unet.h)
  _ffi.py          ctypes binding of libb200unet.so
  ops.py           operator wrappers (torch tensors -> C-ABI calls)
  keras/           Keras-shaped layer / Model / optimizer / callback API of the reference
  shared/          mirror of the reference's shared/custom_layers.py interface
"""
__version__ = "0.1.0"
