#!/usr/bin/env python3
"""Console entry point with the reference's script name (setup.py:88-92 installs ttl_track.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracktolearn_b200.runners.ttl_track import main  # noqa: E402

if __name__ == '__main__':
    main()
