"""Importable alias for the hyphenated package directory `html5-canvas-raytracer_b200/`."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("html5-canvas-raytracer_b200")
