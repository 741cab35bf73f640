#!/usr/bin/env python
"""drop-in for the reference's scripts/extract_PCR_product.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiprime_b200.pcr_product import main  # noqa: E402

if __name__ == "__main__":
    main()
