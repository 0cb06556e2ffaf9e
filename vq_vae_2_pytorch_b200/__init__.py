"""B200-native implementation of the VQ-VAE-2 `Quantize` hot path (drop-in for
alehdaghi/vq-vae-2-pytorch `vqvae.Quantize`).  See DESIGN.md / INTEGRATION.md."""
from .quantize import Quantize, row_layout  # noqa: F401
from . import distributed  # noqa: F401
from . import _native  # noqa: F401
from .egress import CodeEgress, unpack_codes  # noqa: F401
from .trainer_glue import DeferredMetrics, ddp_wrap, replicas_identical  # noqa: F401

__all__ = ["Quantize", "row_layout", "distributed", "CodeEgress", "unpack_codes", "DeferredMetrics", "ddp_wrap", "replicas_identical"]
