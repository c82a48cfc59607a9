"""B200-native observation hot path of FactoryTaskInsertionTactile
(osheraz/IsaacGymInsertion): batched allsight tactile renderer + external-camera
point-cloud pipeline as hand-written sm_100a kernels behind a C-ABI.

Importing the package does not touch the GPU; the kernels are loaded on first use
and there is no CPU fallback (see `_lib.load`).
"""
__version__ = "0.1.0"
