"""B200-native 3D-DDPM sampling hot path (drop-in for guided_diffusion on that path)."""
from . import _native, dist_util, ensemble, gaussian_diffusion, respace, script_util, slab, unet, volume  # noqa: F401
from .script_util import (  # noqa: F401
    add_dict_to_argparser, args_to_dict, create_gaussian_diffusion, create_model_and_diffusion,
    model_and_diffusion_defaults, sr_create_model, sr_create_model_and_diffusion, sr_model_and_diffusion_defaults,
    str2bool,
)
