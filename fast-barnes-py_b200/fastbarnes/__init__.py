# -*- coding: utf-8 -*-
"""
fastbarnes (B200-native): drop-in for the optimized-convolution path of MeteoSwiss/fast-barnes-py.
See interpolation.py, interpolationS2.py and include/fastbarnes_b200.h.
"""
__version__ = '2.0.0+b200.1'
