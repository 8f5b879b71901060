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
semantics.  The only machine-checkable pins the reference offers are its 15
``model.summary()`` dumps (layer shapes, parameter totals and the ``ceil`` size
chains), which ``tests/test_oracle_pins.py`` checks.  Every function cites the
reference call site (file:line) whose behaviour it restates.
"""
