"""Oracle scaffolding (test infrastructure, never shipped): stand-in for the un-vendored
third-party package `einops-exts==0.0.4` (reference requirements.txt:29), which the reference
UNet imports at video_flow_diffusion_multiGPU_v0_crema_plus_faceemb_ca_multi_test.py:18 and
calls at :616 and :683.  Only `rearrange_many` is used."""
from einops import rearrange


def rearrange_many(tensors, pattern, **kwargs):
    return tuple(rearrange(t, pattern, **kwargs) for t in tensors)
