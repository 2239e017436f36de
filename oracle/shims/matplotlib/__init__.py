"""TEST INFRASTRUCTURE shim: LFG/modules/util.py:17 imports matplotlib.pyplot at module level for its Visualizer class only;
the decode path (Generator.forward_with_flow) never touches it.  Absent from this image; an empty module is enough."""
