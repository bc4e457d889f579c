"""Re-export of bisinger_b200/synthetic.py (seeded synthetic weights and inputs) under the name the oracle-side scripts and the
tests have always used.  The generator lives in the package so that bench.py's product arm and the tools do not import anything from
``oracle/`` (which is test infrastructure: only tests/, smoke() and bench.py's CPU legs may touch it)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from bisinger_b200.synthetic import *  # noqa: E402,F401,F403
from bisinger_b200.synthetic import DIFFNET_CONFIG, HIFIGAN_CONFIG, SPEC_MAX, SPEC_MIN  # noqa: E402,F401
