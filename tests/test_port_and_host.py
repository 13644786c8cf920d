"""CPU: the stock-torch port used as the timed CPU baseline, the C-ABI surface, and host logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import helpers as H
from helpers import O


def test_reference_port_matches_reference_fixture_and_shares_state_dict_keys():
    from oracle import reference_port as RP
    G = H.golden()
    ns, ffn = H.build_nets("cpu")
    pn, pffn = RP.build()
    batch = H.golden_batch()
    for k in H.NET_ORDER:
        pn[k].load_state_dict(ns[k].state_dict(), strict=True)      # identical keys both ways
        pn[k].eval()
        with torch.no_grad():
            y = pn[k](H.net_inputs(batch)[k])
        H.assert_close(y, G[f"{k}/out"], f"port {k}", atol_frac=1e-6)
    pffn.load_state_dict(ffn.state_dict(), strict=True)
    pffn.eval()
    with torch.no_grad():
        lp = pffn(batch.acoustic, batch.visual, batch.text)[0]
    H.assert_close(lp, G["ffn/log_prob"], "port log_prob", atol_frac=1e-6)
    loss = RP.PortMaskedNLLLoss(torch.tensor(H.synthetic.IEMOCAP_LOSS_WEIGHTS))(
        lp.transpose(0, 1).contiguous().view(-1, 6), batch.label.view(-1), batch.umask)
    H.assert_close(loss.item(), G["ffn/loss"], "port loss")


def test_library_exports_every_declared_symbol():
    from gan_ffn_b200 import _lib
    protos = _lib.parse_header()
    text = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(ganffn_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    assert declared == set(protos), declared ^ set(protos)
    assert len(protos) >= 30
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), f"{name} declared in include/ganffn.h but not exported by libganffn.so"
    L = _lib.lib()
    assert L.cdll.ganffn_version() >= 100


def test_host_side_shape_preconditions_without_a_gpu():
    from gan_ffn_b200._lib import lib
    L = lib()
    n = L.query("ganffn_net_stash_floats", 0, 94, 32, 100, 100, 10, 2048, 8, 512, 100)
    assert n > 94 * 32 * 2048 * 8
    assert L.query("ganffn_net_scratch_floats", 1, 94, 32, 512, 100, 10, 2048, 8, 64, 16) > 0
    with pytest.raises(ValueError, match="110"):
        L.query("ganffn_net_stash_floats", 0, 111, 32, 100, 100, 10, 2048, 8, 512, 100)
    with pytest.raises(ValueError, match="nhead"):
        L.query("ganffn_net_stash_floats", 0, 10, 2, 100, 100, 7, 2048, 8, 512, 100)
    with pytest.raises(ValueError, match="visual discriminator"):
        L.query("ganffn_net_stash_floats", 0, 10, 2, 512, 100, 10, 2048, 8, 512, 100)


def test_product_path_refuses_cpu_tensors():
    import gan_ffn_b200 as GB
    g = GB.TextGenerator(100)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.zeros(5, 2, 100))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GB.BCELoss()(torch.rand(4, 1), torch.ones(4, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GB.MaskedNLLLoss()(torch.zeros(4, 6), torch.zeros(4, dtype=torch.long), torch.ones(1, 4))


def test_synthetic_batches_have_loader_shapes_and_padding():
    b = H.synthetic.make_batch(n_dialogues=5, lengths=[3, 11, 7, 1, 11])
    assert b.text.shape == (11, 5, 100) and b.visual.shape == (11, 5, 512) and b.acoustic.shape == (11, 5, 100)
    assert b.qmask.shape == (11, 5, 2) and b.umask.shape == (5, 11) and b.label.shape == (5, 11)
    assert b.label.dtype == torch.int64 and float(b.text.min()) >= 0 and float(b.text.max()) < 1
    assert b.umask.sum().item() == 33 == b.real_utterances and b.padded_slots == 55
    assert float(b.text[3:, 0].abs().max()) == 0 and float(b.visual[1:, 3].abs().max()) == 0
    assert int(b.label[0, 3:].abs().max()) == 0
    assert torch.equal(b.qmask.sum(-1), b.umask.t())
    with pytest.raises(ValueError, match="110"):
        H.synthetic.make_batch(n_dialogues=1, lengths=[111])
    full = H.synthetic.make_batch(n_dialogues=32, seq_len=94)
    assert full.padded_slots == 3008 and full.h2d_bytes() == 3008 * (712 * 4 + 2 * 4 + 4 + 8)


def test_shard_indices_cover_each_dialogue_once():
    from gan_ffn_b200 import parallel
    for n, w in ((32, 8), (33, 4), (5, 8), (16700, 8)):
        seen = sum((parallel.shard_indices(n, w, r) for r in range(w)), [])
        assert seen == list(range(n))
        sizes = [len(parallel.shard_indices(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    b = H.synthetic.make_batch(n_dialogues=6, lengths=[4, 9, 2, 9, 5, 1])
    sh = parallel.shard_batch(b, 4, 1)
    assert sh.seq_len == 9 and sh.lengths == [2, 9] and torch.equal(sh.text, b.text[:, 2:4])


def test_data_parallel_wrapped_reference_checkpoints_load():
    """The reference trains on GPU with every network wrapped in nn.DataParallel and GAN_FFN holding the wrapped
    generators (train_IEMOCAP.py:587-593, :629-635), and pickles whole modules (:427-438): state_dict keys carry
    `module.` segments at the top level and nested.  load_reference_state strips them and loads strictly."""
    import torch.nn as nn
    import gan_ffn_b200 as GB
    from gan_ffn_b200 import model as M
    from oracle import reference_port as RP
    pnets, _ = RP.build(seed=11)
    wrapped = {k: nn.DataParallel(v) for k, v in pnets.items()}                      # :587-593
    pffn = RP.PortGANFFN(wrapped["acoustic_gen"], wrapped["visual_gen"], wrapped["text_gen"], 6, dropout=0.6)   # :629-635
    keys = list(pffn.state_dict())
    assert any(k.startswith("acoustic_generator.module.") for k in keys)
    assert all(k.startswith("module.") for k in wrapped["text_disc"].state_dict())

    torch.manual_seed(5)
    ours = {"acoustic_gen": M.AcousticGenerator(100), "visual_gen": M.VisualGenerator(100), "text_gen": M.TextGenerator(100),
            "text_disc": M.TextDiscriminator(100), "visual_disc": M.VisualDiscriminator(100)}
    ffn = M.GAN_FFN(ours["acoustic_gen"], ours["visual_gen"], ours["text_gen"], n_classes=6, dropout=0.6)
    with pytest.raises(RuntimeError):
        ours["text_disc"].load_state_dict(wrapped["text_disc"].state_dict())         # the prefixed keys do not load as is
    for k in ("text_disc", "visual_disc"):
        res = GB.load_reference_state(ours[k], wrapped[k])                           # a wrapped module ...
        assert not res.missing_keys and not res.unexpected_keys
    res = GB.load_reference_state(ffn, pffn.state_dict())                            # ... or its state_dict, nested wrappers
    assert not res.missing_keys and not res.unexpected_keys
    for k in ("acoustic_gen", "visual_gen", "text_gen", "text_disc", "visual_disc"):
        for (n1, p1), (n2, p2) in zip(sorted(ours[k].state_dict().items()), sorted(pnets[k].state_dict().items())):
            assert n1 == n2 and torch.equal(p1, p2), (k, n1, n2)
    assert torch.equal(ffn.fc.weight, pffn.fc.weight)
    # a plain (unwrapped) checkpoint goes through the same call
    GB.load_reference_state(ours["text_gen"], pnets["text_gen"].state_dict())
