"""tqdne_b200 -- B200-native (sm_100a) EDM waveform-sampling engine behind the tqdne module API.

Hot path: Heun sampler -> conditional UNet denoiser -> latent decoder -> log-spectrogram inverse.
Everything numerical runs in hand-written CUDA kernels behind the C-ABI in include/tqdne_b200.h;
there is no CPU fallback.  See DESIGN.md / INTEGRATION.md.
"""

__version__ = "0.1.0"

from .architectures import (  # noqa: F401
    get_1d_autoencoder_configs,
    get_1d_unet_config,
    get_2d_autoencoder_configs,
    get_2d_unet_config,
)
from .autoencoder import LightningAutoencoder  # noqa: F401
from .blocks import Decoder, Encoder  # noqa: F401
from .classifier import LithningClassifier  # noqa: F401
from .edm import EDM, LightningEDM  # noqa: F401
from .representation import Identity, LogSpectrogram, MovingAverageEnvelope, Normalization  # noqa: F401
from .unet import UNetModel  # noqa: F401
