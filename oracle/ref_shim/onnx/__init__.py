"""Empty stand-in for `onnx` (absent in this image); the reference's torch branch
of GenericNNetWrapper.predict is used instead. Test infrastructure only."""
