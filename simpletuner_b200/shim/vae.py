"""Seam B8: VAE latent encode.  `load_vae` ends with `self.post_vae_load_setup()` (common.py:2747-2757) and
`VAECache.encode_images` reaches the VAE through `encode_cache_batch -> encode_with_vae(vae, samples)`
(common.py:2766-2779), accepting an object with `.latent_dist.sample()` (caching/vae.py:1331-1343)."""
from __future__ import annotations

import logging

logger = logging.getLogger("simpletuner_b200.shim")

_CFG = ("in_channels", "latent_channels", "block_out_channels", "layers_per_block", "norm_num_groups", "use_quant_conv",
        "scaling_factor", "shift_factor", "mid_block_add_attention")


def build_b200_vae(ref_vae, device):
    """B200 AutoencoderKL (encoder half) carrying the reference VAE's encoder / quant_conv weights."""
    from ..vae.autoencoder import AutoencoderKL

    c = ref_vae.config
    get = (lambda k, d=None: c.get(k, d)) if isinstance(c, dict) else (lambda k, d=None: getattr(c, k, d))
    kw = {k: get(k) for k in _CFG if get(k, None) is not None or k == "shift_factor"}
    vae = AutoencoderKL(**kw)
    enc = {k: v for k, v in ref_vae.state_dict().items() if k.startswith(("encoder.", "quant_conv."))}
    vae.load_state_dict(enc, strict=True)
    return vae.to(device)


class B200VAEMixin:
    _b200_vae = None

    def post_vae_load_setup(self):
        parent = getattr(super(), "post_vae_load_setup", None)
        if callable(parent):
            parent()
        self._b200_vae = None
        try:
            self._b200_vae = build_b200_vae(self.vae, self.accelerator.device)
        except (NotImplementedError, RuntimeError, KeyError, TypeError) as exc:   # unsupported VAE flavour -> reference path
            logger.warning("libstb200 VAE encode not used: %s", exc)

    def encode_with_vae(self, vae, samples):
        if self._b200_vae is None or vae is not getattr(self, "vae", None):
            return super().encode_with_vae(vae, samples)
        return self._b200_vae.encode(samples)
