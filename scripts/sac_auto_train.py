#!/usr/bin/env python
"""Entry point of the training runner (reference: setup.py console script sac_auto_train.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tracktolearn_b200.trainers.sac_auto_train import main  # noqa: E402

if __name__ == '__main__':
    main()
