"""Import-only stub (reference Spiking_submodules.py:5 imports it and never calls it)."""
