"""pymhealth_b200 -- B200-native (sm_100a) drop-in for the sliding-window / spectral /
location feature hot path of callumstew/pymhealth (package ``mhealth``).

The sub-packages mirror the reference's module layout for that path only:
``util`` (window drivers), ``generic`` (reducers), ``fft``, ``heart.hrv`` (PSD band reducers),
``location`` and ``inertial`` -- numpy arrays in, numpy arrays out, CUDA kernels underneath,
no CPU fallback.
"""
__version__ = "0.1.0"
