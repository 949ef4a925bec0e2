"""Import-only stub (reference Spiking_modules.py:14 imports it and never calls it)."""
