"""omp_amg_b200 -- B200-native AMG hierarchy construction behind the reference's interface.

The product is the CUDA shared library ``libomp_amg_b200.so`` (C ABI in
``include/omp_amg_b200.h``); this package is the thin host-side mirror of the reference's
entry points (``amg_setup`` / ``amg_export`` of amg_setup.h, ``crs_setup`` / ``crs_solve`` /
``crs_stats`` / ``crs_free`` of crs.h) over ctypes.  There is no CPU implementation: if the
library or a CUDA device is missing, calls raise.
"""
from .api import (AmgError, Hierarchy, amg_setup, amg_setup_from_dump, crs_setup, crs_solve,  # noqa: F401
                  crs_stats, crs_free, lib, lib_path, build_info, device_count)
from . import matrices  # noqa: F401

__all__ = ["AmgError", "Hierarchy", "amg_setup", "amg_setup_from_dump", "crs_setup", "crs_solve",
           "crs_stats", "crs_free", "lib", "lib_path", "build_info", "device_count", "matrices"]
