#!/usr/bin/env python
"""drop-in for the reference's scripts/get_degePrimer.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiprime_b200.degeprimer import main  # noqa: E402

if __name__ == "__main__":
    main()
