"""h264bsd_b200 -- B200-native H.264 Baseline macroblock-reconstruction engine behind the C API of
oneam/h264bsd.  The product is libh264bsd_b200.so (csrc/); these modules are thin ctypes mirrors."""
from .decoder import H264bsdDecoder, decode_stream  # noqa: F401
from .batch import Batch, ParsedStream  # noqa: F401
