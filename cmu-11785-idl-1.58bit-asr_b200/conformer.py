"""Compatibility name: the model lives in ``asr_model.py``; ``onebit_b200.conformer`` re-exports it."""
from .asr_model import *  # noqa: F401,F403
from .asr_model import (ConformerASR, ConformerBlock, ConformerEncoder, ConvModule, Conv2dSubsampling,  # noqa: F401
                        FeedForwardModule, LayerNorm, MHSA, ModelDims, RelPositionalEncoding, TransformerDecoder, swish)
