"""``python src/experiments/peclr_training.py <reference flags>`` -- same entrypoint as the reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from peclr_b200.peclr_training import main  # noqa: E402

if __name__ == "__main__":
    main()
