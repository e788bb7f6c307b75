"""non-decimated_wavelets_b200 -- B200-native multi-level non-decimated wavelet dec/rec (1-D..4-D)
behind the API of arg-min-x/Non-Decimated_Wavelets.

The directory name follows the reference repo and is not a Python identifier; import it with
`importlib.import_module("non-decimated_wavelets_b200")` or through the root-level alias
`import nddwt_b200`.
"""
from .api import (FilterSpec, harr_nddwt_2D, harr_nddwt_4D, nd_dwt_1D, nd_dwt_2D, nd_dwt_3D, nd_dwt_4D,
                  nd_dwt_mex, to_device, to_host, wave_filters)
from ._lib import LIB_PATH, NddwtError, Plan, lib

__all__ = ["wave_filters", "nd_dwt_1D", "nd_dwt_2D", "nd_dwt_3D", "nd_dwt_4D", "harr_nddwt_2D",
           "harr_nddwt_4D", "nd_dwt_mex", "FilterSpec", "to_device", "to_host", "Plan", "NddwtError", "lib",
           "LIB_PATH"]
