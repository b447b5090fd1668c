"""onebit_b200 - B200-native quantised linear layer (drop-in for the reference's onebit_asr/quant.py).

Import name: ``onebit_b200`` (the repo-root shim ``onebit_b200.py`` maps it onto this directory, whose
name is not a Python identifier).  ``install_as_reference_quant()`` registers the module as ``quant`` so
the reference's ``conformer.py`` (which does ``from quant import QuantizedLinear``, conformer.py:12)
picks this layer up unchanged.
"""
from . import _cabi, asr_model, attention, conformer, dp, inference, norm, routes, training
from .asr_model import ConformerASR
from .inference import PackedQuantizedLinear, ctc_greedy_decode, pack_model_for_inference
from .quant import BitLinear, QuantizedLinear, act_quant_int8, install_as_reference_quant, quantize_weight

__all__ = ["ConformerASR", "PackedQuantizedLinear", "pack_model_for_inference", "ctc_greedy_decode", "inference", "conformer", "training", "dp", "QuantizedLinear", "BitLinear", "quantize_weight", "act_quant_int8", "install_as_reference_quant", "_cabi"]
