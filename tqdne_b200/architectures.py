"""Named architectures (same dictionaries as tqdne/architectures.py:1-79)."""


def _autoencoder_pair(config, dims, kernel):
    base = dict(model_channels=64, channel_mult=(1, 2, 4), attention_resolutions=(), num_res_blocks=2, dims=dims,
                conv_kernel_size=kernel, dropout=0.1)
    enc = base | dict(in_channels=config.channels, out_channels=config.latent_channels * 2)
    dec = base | dict(in_channels=config.latent_channels, out_channels=config.channels)
    return enc, dec


def get_1d_autoencoder_configs(config):
    return _autoencoder_pair(config, 1, 5)


def get_2d_autoencoder_configs(config):
    return _autoencoder_pair(config, 2, 3)


def _unet(config, in_channels, out_channels, dims, kernel, model_channels, **extra):
    return dict(in_channels=in_channels, out_channels=out_channels, cond_features=len(config.features_keys), dims=dims,
                conv_kernel_size=kernel, model_channels=model_channels, channel_mult=(1, 2, 4, 4),
                attention_resolutions=(8,), num_res_blocks=2, num_heads=4, dropout=0.1, flash_attention=False, **extra)


def get_1d_unet_config(config, in_channels, out_channels):
    return _unet(config, in_channels, out_channels, 1, 5, 64)


def get_2d_unet_config(config, in_channels, out_channels, model_channels=128, use_causal_mask=False):
    return _unet(config, in_channels, out_channels, 2, 3, model_channels, use_causal_mask=use_causal_mask)
