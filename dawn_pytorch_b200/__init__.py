"""B200-native (sm_100a) implementation of DAWN's per-step denoising UNet behind the reference's
Python module interface (DynamicNfUnet3D / DynamicNfGaussianDiffusion)."""
from .unet import DynamicNfUnet3D, Unet3D  # noqa: F401
from .diffusion import DynamicNfGaussianDiffusion, GaussianDiffusion  # noqa: F401
from .lfg import Generator as LfgGenerator  # noqa: F401
from .flow_diffusion import Face_loc_Encoder, FlowDiffusion  # noqa: F401
