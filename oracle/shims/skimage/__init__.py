"""TEST INFRASTRUCTURE shim: LFG/modules/util.py:18 imports skimage.draw.disk for its Visualizer class only."""
