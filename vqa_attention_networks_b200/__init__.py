"""B200-native (sm_100a) implementation of the MFB/MFH fusion and co-attention hot path of
klory/vqa-attention-networks, behind the reference's own nn.Module interface.

    from vqa_attention_networks_b200 import MFB, MHBCoAtt        # same ctor / forward as the reference

The kernels live in ``csrc/`` behind the C ABI declared in ``include/vqa_b200.h``; see DESIGN.md.
"""
from .hieCoAtten import HieCoAtten  # noqa: F401
from .mfb import MFB  # noqa: F401
from .mhb_coAtt import MHB, MHBCoAtt  # noqa: F401
from .modules import Attention_1, Attention_2, Attention_layer, Nonlinear_layer  # noqa: F401

__all__ = ["MFB", "MHBCoAtt", "MHB", "HieCoAtten", "Attention_layer", "Attention_1", "Attention_2", "Nonlinear_layer"]
