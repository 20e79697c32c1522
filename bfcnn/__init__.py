"""`bfcnn` drop-in alias: `import bfcnn; bfcnn.load_model(name)(uint8[N,H,W,3])` as in the
reference README (README.md:128-151), served by the B200-native package."""
from blind_image_denoising_b200 import *  # noqa: F401,F403
from blind_image_denoising_b200 import (CONFIGS_DICT, configs, load_default_denoiser,  # noqa: F401
                                        load_denoiser_model, load_model, models)
from blind_image_denoising_b200 import __version__  # noqa: F401
