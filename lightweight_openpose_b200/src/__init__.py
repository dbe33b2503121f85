"""Drop-in mirror of the reference's ``src`` package for the inference hot path:
``lightweight_openpose``, ``net_factory`` and ``pose_decode`` keep the reference's
function names, argument meaning and return types, and run on hand-written
sm_100a kernels through the C ABI in ``include/lwop.h``."""
