"""Gives the tests access to helper functions of tests/golden/make_golden.py (TEST INFRASTRUCTURE)."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
oracle_votes = _mod.oracle_votes
