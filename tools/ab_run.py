"""Run a script of this repo against another build of the library (A/B on one box):
    ITR_LIB=path/to/libitrails_b200.so python tools/ab_run.py bench.py --no-cpu-baseline"""
import os, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from itrails_b200 import _lib
if os.environ.get("ITR_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["ITR_LIB"])
sys.argv = sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
