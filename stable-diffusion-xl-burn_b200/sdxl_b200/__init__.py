"""sdxl_b200 — host-side mirror of the reference's diffusion sampling surface over libsdxl_b200.so."""
from .config import SDXL_BASE, SDXL_REFINER, TINY, TINY_REFINER, SDXL_VAE, TINY_VAE, UNetConfig, VaeConfig, block_program  # noqa: F401
from .weights import alphas_cumprod, build_pack, n_params, synth_weights, unet_tensor_specs, vae_decoder_tensor_specs  # noqa: F401
from ._lib import LIB_PATH, PROTOTYPES, SdxlError, SdxlLibraryMissing, load  # noqa: F401
from .engine import Conditioning, Context, Diffuser, LatentDecoder, ddim_timesteps  # noqa: F401
