"""See skimage/__init__.py (shim)."""


def disk(*a, **k):
    raise NotImplementedError("skimage is shimmed: only needed by the reference's Visualizer")
