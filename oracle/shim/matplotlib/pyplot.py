def __getattr__(name):
    raise AttributeError(f"matplotlib shim: pyplot.{name} is not available (render paths are out of scope)")
