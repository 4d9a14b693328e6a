"""fitv2_b200 — B200-native (sm_100a) FiTv2 denoising hot path behind the reference's FiT interface."""
from .model import FiT
from .sampler import EulerCFGSampler, euler_cfg_sample, make_grid, pack_images_uint8
from .transport import Sampler, Transport, create_transport
from .checkpoint import init_from_ckpt
from ._lib import FitV2Error

__all__ = ["FiT", "EulerCFGSampler", "euler_cfg_sample", "make_grid", "pack_images_uint8", "FitV2Error", "Sampler", "Transport",
           "create_transport", "init_from_ckpt"]
