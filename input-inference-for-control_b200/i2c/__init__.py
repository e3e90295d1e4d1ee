"""Drop-in mirror of the reference's ``i2c`` package for the hot path.

Same module / class / method names as JoeMWatson/input-inference-for-control so that callers written against the
reference (``from i2c.i2c import I2cGraph`` ...) run on the CUDA path: every sweep is executed by
``libi2c_b200.so`` through ``i2c_b200.BatchedI2c`` with one problem (B = 1).  There is no NumPy implementation
of the sweeps in this package -- without the library or a GPU the constructors raise.
"""
