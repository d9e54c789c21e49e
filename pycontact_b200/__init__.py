"""pycontact_b200 -- B200-native implementation of PyContact's per-frame
contact-analysis hot path (neighbour search -> contact score -> H-bond test ->
accumulation), behind the reference's own analysis API.

Importing the package never touches the GPU; the CUDA library is loaded on first
use and its absence is a hard error (there is no CPU fallback).
"""
__version__ = "0.1.0"
