"""Loading the reference's checkpoints into the B200 modules.

The reference saves *whole module objects* (``torch.save(model, path)``, train_IEMOCAP.py:427-438).  When it trained on a
GPU the six networks were ``nn.DataParallel`` instances (train_IEMOCAP.py:587-593), so every ``state_dict`` key carries a
``module.`` segment, and ``GAN_FFN`` holds the *wrapped* generators (:629-635): its keys read
``acoustic_generator.module.transformer_encoder.layers.0...``.  This module removes those segments -- at any depth --
so that such a checkpoint loads strictly into the unwrapped modules of this package (whose keys are the reference's
own, see tests/test_oracle_golden.py).
"""
from __future__ import annotations

import re
from typing import Dict, Mapping, Union

import torch

_DP_SEGMENT = re.compile(r"(^|\.)module\.")


def strip_data_parallel(state: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``state_dict`` with every ``nn.DataParallel`` ``module.`` segment removed (prefix or nested)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in state.items():
        nk = k
        while True:
            stripped = _DP_SEGMENT.sub(r"\1", nk, count=1)
            if stripped == nk:
                break
            nk = stripped
        if nk in out:
            raise KeyError(f"stripping 'module.' makes two keys collide on {nk!r}")
        out[nk] = v
    return out


def load_reference_state(target: torch.nn.Module, source: Union[torch.nn.Module, Mapping[str, torch.Tensor]],
                         strict: bool = True):
    """Loads a reference network / ``GAN_FFN`` (a module or its ``state_dict``; plain or ``nn.DataParallel``-wrapped at
    any level) into ``target``.  For a whole-module pickle: ``src = torch.load(path, weights_only=False)`` with the
    reference's ``model.py`` importable (the pickle names ``model.<Class>``), then ``load_reference_state(ours, src)``."""
    state = source.state_dict() if isinstance(source, torch.nn.Module) else source
    return target.load_state_dict(strip_data_parallel(state), strict=strict)
