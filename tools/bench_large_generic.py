"""tools/bench_configs.py `large` with every plan forced onto the generic kernel (A/B for the multi-warp-frame kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meyda_b200 as mb
from meyda_b200 import _capi
_orig = mb.Plan.__init__
def _init(self, *a, **k):
    k["flags"] = k.get("flags", 0) | _capi.MB_FLAG_GENERIC_KERNEL
    _orig(self, *a, **k)
mb.Plan.__init__ = _init
sys.argv = [sys.argv[0], "large"]
exec(open(os.path.join(ROOT, "tools", "bench_configs.py")).read())
