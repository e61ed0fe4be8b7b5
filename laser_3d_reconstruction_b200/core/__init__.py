"""Core module (reference core/__init__.py:6-7)."""
from .laser_extractor import FastStegerExtractor, SimpleLaserExtractor
from .reconstruction import Reconstructor

__all__ = ["SimpleLaserExtractor", "FastStegerExtractor", "Reconstructor"]
