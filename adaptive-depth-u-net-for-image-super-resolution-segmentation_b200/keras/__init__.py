"""Keras-shaped API surface of the reference, backed by the sm_100a kernels.

    from b200unet.keras import Input, Model, layers as L, mixed_precision
    from b200unet.keras.callbacks import EarlyStopping, ModelCheckpoint, BackupAndRestore
"""
from . import callbacks, layers, losses, optimizers
from .layers import Input, clear_session
from .model import History, Model, set_random_seed


class mixed_precision:  # namespace, as keras.mixed_precision
    from .model import global_policy, set_global_policy
    set_global_policy = staticmethod(set_global_policy)
    global_policy = staticmethod(global_policy)


__all__ = ["Input", "Model", "layers", "losses", "optimizers", "callbacks", "mixed_precision", "clear_session",
           "set_random_seed", "History"]
