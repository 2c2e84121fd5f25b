"""Empty stand-in for `onnxruntime` (absent in this image). Test infrastructure only."""
__version__ = '0.0.0'
