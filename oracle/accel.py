"""Accelerometer pre-stage of the reference, restated.  TEST INFRASTRUCTURE ONLY.

Reference: ``src/mhealth/inertial/accelerometer.py`` -- ``roll`` (:13-25), ``pitch`` (:44-56), ``magnitude`` (:198-225),
``magnitude_dot`` (:236-259).  numba evaluates these elementwise on arrays; the numpy forms below produce the same
types [pinned by tests/golden/ref_extra.npz]: magnitude keeps the float type of its input, roll / pitch take arctan2 in
the input type and widen to float64 through the ``* 180 / np.pi`` scaling.
"""
import numpy as np


def roll(y, z):
    """accelerometer.py:25 -- ``np.arctan2(y, z) * 180/np.pi``."""
    return np.arctan2(y, z).astype(np.float64) * 180 / np.pi


def pitch(x, y, z):
    """accelerometer.py:56 -- ``np.arctan2(-x, np.sqrt(y*y + z*z)) * 180/np.pi``."""
    return np.arctan2(-x, np.sqrt(y * y + z * z)).astype(np.float64) * 180 / np.pi


def magnitude(x, y, z):
    """accelerometer.py:225 -- ``np.sqrt(x**2 + y**2 + z**2)`` in the input type."""
    return np.sqrt(x ** 2 + y ** 2 + z ** 2)


def magnitude_dot(x, y, z):
    """accelerometer.py:259 -- ``np.sqrt(np.dot(x, x) + np.dot(y, y) + np.dot(z, z))``."""
    return float(np.sqrt(np.dot(x, x) + np.dot(y, y) + np.dot(z, z)))
