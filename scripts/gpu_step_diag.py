"""Runs every parity check of tests/parity_checks.py on the GPU, printing one JSON line per check."""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "parity_checks.py"),
               run_name="__main__")
