"""See matplotlib/__init__.py (shim)."""


def get_cmap(*a, **k):
    raise NotImplementedError("matplotlib is shimmed: only needed by the reference's Visualizer")
