"""Parity and boundary tests of goofer_b200 (see conftest.py for the gpu / reference markers)."""
