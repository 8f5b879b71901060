"""Model builders with the reference's names and signatures.

  * ``conv_block`` / ``build_super_resolution_unet`` / ``build_losses_and_metrics``
      /root/reference/Super_resolution/code/train_adaptive_unet.py:200-210, :217-287, :294-373
  * ``build_vanilla_super_resolution_unet`` (the reference's second ``build_super_resolution_unet``)
      /root/reference/Super_resolution/code/u-net-vinillia.py:128-167
  * ``seg_conv_block`` / ``build_adaptive_depth_unet``
      /root/reference/Segmenation/code/train_adaptive_unet.py:325-332, :335-362
  * ``encoder_block`` / ``decoder_block`` / ``build_unet``
      /root/reference/Segmenation/code/unet_vinillia.py:60-63, :66-69, :72-91

They only compose the Keras-shaped layers of ``b200unet.keras``; the recorded graph is lowered to
the sm_100a kernels by ``keras/engine.py``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

from .keras import Input, Model
from .keras import layers as L
from .keras import losses as LS
from .shared.custom_layers import (ClippedResidualAdd, ResizeByScale, ResizeToMatch, custom_depth_from_scale,
                                   estimate_bottleneck_size)


def conv_block(inputs, nf: int):
    """Two rounds of Conv3x3(nf, same, bias) -> LayerNormalization(channels) -> ReLU."""
    x = inputs
    for _ in range(2):
        x = L.Conv2D(nf, 3, padding="same", use_bias=True)(x)
        x = L.LayerNormalization(axis=-1)(x)
        x = L.Activation("relu")(x)
    return x


def build_super_resolution_unet(scale: float, base_channels: int = 64, residual_head_channels: int = 64,
                                depth_override: Optional[int] = None, input_size: int = 256,
                                max_depth: int = 7) -> Tuple[Model, Dict[str, object]]:
    """Adaptive-depth SR U-Net: encoder (conv_block + ResizeByScale) x depth, bottleneck, decoder
    (ResizeToMatch + Conv3x3/ReLU + concat[up, skip] + conv_block) x depth, residual head
    (conv_block -> zero-initialised 1x1 conv -> clip(input + residual))."""
    depth = depth_override if depth_override is not None else custom_depth_from_scale(
        scale, max_depth=max_depth, base_resolution=input_size)
    shrink = ResizeByScale(scale, name="enc_down")
    grow = ResizeToMatch(name="dec_up")
    inputs = Input(shape=(input_size, input_size, 3), name="low_res_input")

    skips, x, width = [], inputs, base_channels
    for _ in range(depth):
        skips.append(conv_block(x, width))
        x = shrink(skips[-1])
        width *= 2
    x = conv_block(x, width)
    for skip in skips[::-1]:
        width //= 2
        x = grow([x, skip])
        x = L.Conv2D(width, 3, padding="same", activation="relu")(x)
        x = L.Concatenate()([x, skip])
        x = conv_block(x, width)

    head = conv_block(x, residual_head_channels)
    residual = L.Conv2D(3, 1, padding="same", kernel_initializer="zeros", bias_initializer="zeros",
                        name="residual_rgb")(head)
    outputs = ClippedResidualAdd(name="enhanced_rgb")([inputs, residual])
    model = Model(inputs, outputs, name=f"U-Net_SR_scale{scale:.2f}_depth{depth}")
    info = {"scale": scale, "depth": depth, "bottleneck_size": estimate_bottleneck_size(input_size, scale, depth),
            "base_channels": base_channels, "max_depth": max_depth}
    return model, info


def build_losses_and_metrics(loss_name: str):
    """("charbonnier" | "l1") -> (loss, [psnr]).  The reference's third option, "combined", needs
    ImageNet VGG19 weights downloaded at run time and is out of scope."""
    key = loss_name.lower()
    if key in ("charbonnier", "l1"):
        return LS.SRLoss(key), [LS.PSNRMetric()]
    if key == "combined":
        raise NotImplementedError("loss 'combined' needs downloaded VGG19 ImageNet weights (no network here); "
                                  "use 'charbonnier' or 'l1'")
    raise ValueError(f"Unknown loss '{loss_name}'. Expected one of: 'charbonnier', 'l1', 'combined'.")


def build_vanilla_super_resolution_unet(input_shape=(256, 256, 3), base_channels: int = 64, depth: int = 4) -> Model:
    """The fixed-depth baseline of ``u-net-vinillia.py``: BatchNorm conv blocks, MaxPool2D down, bilinear
    UpSampling2D + Conv3x3/ReLU up, concat[up, skip], 3-channel sigmoid head (``base_channels`` / ``depth`` are
    64 / 4 in the reference, exposed here for small test nets).  Train it with ``"mse"`` (the MSE term of its
    ``combined`` loss; the SSIM and VGG19 terms need downloaded ImageNet weights -- out of scope)."""
    inputs = Input(shape=tuple(input_shape), name="low_res_input")
    x, skips, nf = inputs, [], base_channels
    for _ in range(depth):
        skip = seg_conv_block(x, nf)
        skips.append((nf, skip))
        x = L.MaxPooling2D(pool_size=(2, 2))(skip)
        nf *= 2
    x = seg_conv_block(x, nf)
    for nf, skip in skips[::-1]:
        x = L.UpSampling2D(size=(2, 2), interpolation="bilinear")(x)
        x = L.Conv2D(nf, 3, padding="same", activation="relu")(x)
        x = L.Concatenate()([x, skip])
        x = seg_conv_block(x, nf)
    outputs = L.Conv2D(3, 1, padding="same", activation="sigmoid", name="enhanced_rgb")(x)
    h, w = input_shape[0], input_shape[1]
    return Model(inputs, outputs, name=f"U-Net_SR_{h}x{w}")


# --------------------------------------------------------------------------- segmentation
def seg_conv_block(inputs, filters: int):
    """Two rounds of Conv3x3(filters, same, bias) -> BatchNormalization -> ReLU."""
    x = inputs
    for _ in range(2):
        x = L.Conv2D(filters, 3, padding="same", use_bias=True)(x)
        x = L.BatchNormalization()(x)
        x = L.Activation("relu")(x)
    return x


def build_adaptive_depth_unet(input_size: int, base_channels: int, depth: int) -> Model:
    """BatchNorm U-Net with MaxPooling2D down, bilinear UpSampling2D up (no channel-reducing conv,
    so the decoder block sees 3f channels) and a 1-channel sigmoid head."""
    inputs = Input(shape=(input_size, input_size, 3), name="isic_image")
    x, skips, f = inputs, [], base_channels
    for _ in range(depth):
        x = seg_conv_block(x, f)
        skips.append((f, x))
        x = L.MaxPooling2D(pool_size=(2, 2))(x)
        f *= 2
    x = seg_conv_block(x, f)
    for f, skip in skips[::-1]:
        x = L.UpSampling2D(size=(2, 2), interpolation="bilinear")(x)
        x = L.Concatenate()([x, skip])
        x = seg_conv_block(x, f)
    outputs = L.Conv2D(1, 1, activation="sigmoid", name="lesion_mask")(x)
    return Model(inputs=inputs, outputs=outputs, name=f"adaptive_unet_depth{depth}_c{base_channels}")


def encoder_block(x, nf: int):
    skip = conv_block(x, nf)
    return L.MaxPooling2D(2)(skip), skip


def decoder_block(x, skip, nf: int):
    x = L.Conv2DTranspose(nf, 2, strides=2, padding="same")(x)
    x = L.Concatenate()([x, skip])
    return conv_block(x, nf)


def build_unet(input_size: int, num_classes: int = 1, base_channels: int = 32, depth: int = 4) -> Model:
    """LayerNorm U-Net with MaxPooling2D / Conv2DTranspose(2, strides 2); sigmoid head for one
    class, softmax head otherwise."""
    inputs = Input(shape=(input_size, input_size, 3), name="images")
    x, skips, nf = inputs, [], base_channels
    for _ in range(depth):
        x, skip = encoder_block(x, nf)
        skips.append(skip)
        nf *= 2
    x = conv_block(x, nf)
    for skip in skips[::-1]:
        nf //= 2
        x = decoder_block(x, skip, nf)
    act = "sigmoid" if num_classes == 1 else "softmax"
    outputs = L.Conv2D(num_classes, 1, activation=act, name="mask_logits")(x)
    return Model(inputs, outputs, name="unet_isic_baseline")
