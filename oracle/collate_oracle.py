"""CPU restatement of the reference's ``collate_fn`` (dataloader.py:55-58) -- TEST INFRASTRUCTURE ONLY.

Checker for ``gan_ffn_b200.pipeline.collate_on_device`` (tests/test_pipeline.py); never imported by the product path.
Pinned by construction: it *is* the reference's expression -- ``pad_sequence`` over the first four columns of the items
(seq-major), ``pad_sequence(..., batch_first=True)`` over umask and label -- applied to items shaped like
``IEMOCAPDataset.__getitem__`` returns them (dataloader.py:41-51)."""
from typing import Sequence

import torch
from torch.nn.utils.rnn import pad_sequence


def collate_reference(items: Sequence[Sequence[torch.Tensor]]):
    """Returns a ``gan_ffn_b200.synthetic.Batch`` holding what ``collate_fn`` returns for the six tensors of a batch."""
    from gan_ffn_b200.synthetic import Batch
    cols = [pad_sequence([it[k] for it in items]) if k < 4 else pad_sequence([it[k] for it in items], True) for k in range(6)]
    return Batch(cols[0], cols[1], cols[2], cols[3], cols[4], cols[5], [int(it[0].shape[0]) for it in items])
