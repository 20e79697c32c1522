"""`bfcnn` drop-in alias: `import bfcnn; bfcnn.load_model(name)(uint8[N,H,W,3])` as in the
reference README (README.md:128-151), served by the B200-native package.  Every name of the reference's
`bfcnn.__all__` (bfcnn/__init__.py:129-143) that lies on the hot path is importable from here:
models, configs, train_loop, load_model, load_image, model_builder, schedule_builder, optimizer_builder,
load_denoiser_model, load_default_denoiser (export_model and the pyramid builders are out of scope)."""
from blind_image_denoising_b200 import *  # noqa: F401,F403
from blind_image_denoising_b200 import (CONFIGS_DICT, configs, dataset_builder, load_config,  # noqa: F401
                                        load_default_denoiser, load_denoiser_model, load_image, load_model,
                                        loss_function_builder, model_builder, models, optimizer_builder,
                                        schedule_builder, train_loop)
from blind_image_denoising_b200 import __version__  # noqa: F401
