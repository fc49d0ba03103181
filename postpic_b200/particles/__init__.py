"""Particle-to-grid deposition: the `postpic.particles` names that sit on the hot path
(/root/reference/postpic/particles/__init__.py:27-32)."""
from ._routines import histogramdd, particleshapes
from .particles import MultiSpecies, Field, ArraySource

__all__ = ['histogramdd', 'particleshapes', 'MultiSpecies', 'Field', 'ArraySource']
