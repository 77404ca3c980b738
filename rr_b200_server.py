"""`python -m rr_b200_server --config config/config.yaml` — gateway launcher (see rr_b200.server)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from rr_b200.server import main  # noqa: E402

if __name__ == "__main__":
    main()
