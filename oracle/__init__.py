"""CPU oracle for the adaptive-depth U-Net hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy + torch-CPU, fp32/fp64), the arithmetic
of the reference's encoder/decoder convolution stack so that the CUDA kernels of
the product can be checked against it.  It is *not* part of the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and then only as the checker (or as the
timed CPU baseline), never as the thing that ships.

PARITY UNPINNED (op level).  The reference is pure TensorFlow 2.16.1 / Keras
3.3.3 Python (``Super_resolution/requirement.txt:4,8``); neither package is
installable in this image (no wheel, no network) and the reference repo holds no
tests, golden vectors or fixtures for this path.  The arithmetic therefore lives
in an absent third-party dependency and is restated here from its published
semantics.  That statement covers the arithmetic INSIDE TensorFlow/Keras: Conv2D,
LayerNormalization, BatchNormalization, tf.image.resize (ScaleAndTranslate), MaxPooling2D,
Conv2DTranspose, Adam, tf.image.ssim.  What IS pinned against the reference:
  * its 15 ``model.summary()`` dumps (layer shapes, parameter totals, the ``ceil`` size
    chains) -- ``tests/test_oracle_pins.py``;
  * everything the reference wrote itself, by RUNNING its code here behind stand-ins for the
    absent packages (``tests/test_reference_boundary_cpu.py``): the losses and metrics
    (Charbonnier / L1 / PSNR, Dice / IoU / the hybrid losses, the baseline's whole-batch Dice,
    BT.601 luma), the custom layers' ``call`` code (clip-add, the float32 size rule), the model
    graphs its builders construct, the depth rules, splits, command lines;
  * the OpenCV patch pipeline (``cv_resize_np.py``), against ``cv2`` and fixtures made by the
    reference's own ``shared/pipeline.py`` -- ``tests/test_pipeline_cpu.py``.
Every function cites the reference call site (file:line) whose behaviour it restates.
"""
