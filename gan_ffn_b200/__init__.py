"""gan_ffn_b200 -- B200 (sm_100a) implementation of the GAN-FFN fusion hot path.

Drop-in for the hot-path classes of the reference's ``model.py`` (same names, constructors,
``forward`` signatures and ``state_dict`` keys); every number is computed by the hand-written
CUDA kernels in ``libganffn.so`` (C ABI: ``include/ganffn.h``).  There is no CPU fallback.
"""
from .model import (AcousticDiscriminator, AcousticGenerator, BCELoss, GAN_FFN, GAN_FFN_DialogueRNN, MaskedNLLLoss,
                    PositionalEncoding,
                    TextDiscriminator, TextGenerator, VisualDiscriminator, VisualGenerator)
from .optim import FusedAdam
from .functional import frozen_parameters, manual_seed, set_deterministic
from .checkpoint import load_reference_state, strip_data_parallel

__all__ = ["AcousticGenerator", "VisualGenerator", "TextGenerator", "AcousticDiscriminator", "VisualDiscriminator",
           "TextDiscriminator", "GAN_FFN", "GAN_FFN_DialogueRNN", "MaskedNLLLoss", "BCELoss", "PositionalEncoding", "FusedAdam",
           "manual_seed", "set_deterministic", "frozen_parameters", "load_reference_state", "strip_data_parallel"]
