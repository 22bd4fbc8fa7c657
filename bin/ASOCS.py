#!/usr/bin/env python
"""Drop-in for the reference's ASOCS.py command line:  ASOCS.py <ini>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soc_b200.asocs import main  # noqa: E402

if __name__ == "__main__":
    main(sys.argv)
