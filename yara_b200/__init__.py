"""yara_b200 -- the B200-native byte-scanning hot path of YARA 4.2.1 behind the
unchanged libyara API.

The product is native: `libyrb200.so` (CUDA engine + C ABI, include/yara_b200.h)
and `libyara_b200.so` (drop-in libyara whose block scan runs on the GPU).  This
package only holds ctypes bindings over those two libraries for benchmarks,
tests and scripting, plus the synthetic ruleset/corpus generators.  There is no
CPU implementation of the path in here; importing works without a GPU, calling
a scan without one raises.
"""
from .api import Engine, LibYara, DeviceBuffer, PinnedBuffer, lib_dir  # noqa: F401
