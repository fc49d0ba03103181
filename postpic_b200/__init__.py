"""postpic_b200 -- B200-native particle-to-grid deposition for postpic.

A from-scratch sm_100a implementation of postpic's one native hot path (the weighted
histogram / histogram2d / histogram3d with particle shapes 0-3), behind the same
Python signatures.  See DESIGN.md and INTEGRATION.md.
"""
from .particles import histogramdd, particleshapes, MultiSpecies, Field, ArraySource

__all__ = ['histogramdd', 'particleshapes', 'MultiSpecies', 'Field', 'ArraySource']
__version__ = '0.1.0'
