"""i2c_b200 -- B200-native batched Gaussian input inference for control (host side of the C-ABI)."""
from . import _capi as capi
from ._capi import I2cError, lib
from . import batched
from .batched import BatchedI2c, gauss_hermite, quadrature, rollout
from . import envs
from .mpc import BatchedPartiallyObservedMpc

__all__ = ["BatchedI2c", "BatchedPartiallyObservedMpc", "quadrature", "gauss_hermite", "rollout", "I2cError", "lib", "capi", "envs"]
