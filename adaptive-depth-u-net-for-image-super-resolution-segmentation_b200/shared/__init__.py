"""Mirrors of the reference's ``shared/`` package (custom_layers, pipeline) plus the segmentation data feed."""
# custom_layers subclasses keras.layers.Layer and keras.engine lowers the custom layers: importing the keras package
# first settles that cycle whichever of the two a caller imports first (``from b200unet.shared.custom_layers import ...``
# is the reference's own first import, Super_resolution/code/train_adaptive_unet.py:27-34).
from .. import keras as _keras  # noqa: F401,E402
