"""Input pipeline: packed host batches, device-side collate, one-batch-ahead prefetch -- against a
CPU restatement of the reference's collate_fn (dataloader.py:55-58, pad_sequence)."""
import pytest
import torch

from gan_ffn_b200 import pipeline, synthetic
from oracle.collate_oracle import collate_reference


def _items(lengths, seed=0, n_classes=6):
    """Per-dialogue tuples shaped like IEMOCAPDataset.__getitem__ (dataloader.py:41-51)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in lengths:
        spk = torch.randint(0, 2, (n,), generator=g)
        out.append((torch.rand(n, 100, generator=g), torch.rand(n, 512, generator=g), torch.rand(n, 100, generator=g),
                    torch.nn.functional.one_hot(spk, 2).float(), torch.ones(n), torch.randint(0, n_classes, (n,), generator=g), f"vid{n}"))
    return out


@pytest.mark.parametrize("lengths", [[5, 1, 9], [110], [3, 3, 3, 3], [1], [17, 94, 2, 40, 40]])
def test_pack_is_the_inverse_of_the_reference_collate(lengths):
    items = _items(lengths, seed=len(lengths))
    pb = pipeline.pack_dialogues(items, pin=False)
    assert pb.lengths_host == lengths and pb.seq_len == max(lengths) and pb.n_dialogues == len(lengths)
    assert pb.node_off.tolist() == [sum(lengths[:i]) for i in range(len(lengths) + 1)]
    assert pb.text.shape == (sum(lengths), 100) and pb.visual.shape == (sum(lengths), 512)
    ref = collate_reference(items)
    back = pipeline.pack_batch(ref, pin=False)
    for a, b in zip(pb.tensors(), back.tensors()):
        assert torch.equal(a, b)
    # the packed batch carries only the real utterances over the bus
    assert pb.h2d_bytes() <= ref.h2d_bytes() + 8 * (len(lengths) + 1) + 4 * len(lengths)
    if min(lengths) < max(lengths):
        assert pb.h2d_bytes() < ref.h2d_bytes()


def test_pack_rejects_an_empty_dialogue():
    items = _items([4, 2])
    items.append(tuple(t[:0] if torch.is_tensor(t) else t for t in items[0]))
    with pytest.raises(ValueError):
        pipeline.pack_dialogues(items, pin=False)


def _same(a: synthetic.Batch, b: synthetic.Batch):
    for name in ("text", "visual", "acoustic", "qmask", "umask", "label"):
        x, y = getattr(a, name).cpu(), getattr(b, name).cpu()
        assert x.shape == y.shape and x.dtype == y.dtype, (name, x.shape, y.shape, x.dtype, y.dtype)
        assert torch.equal(x, y), name
    assert a.lengths == b.lengths


@pytest.mark.gpu
@pytest.mark.parametrize("lengths", [[5, 1, 9], [110], [3, 3, 3, 3], [1], [17, 94, 2, 40, 40], list(range(1, 33))])
def test_device_collate_is_bit_exact(lengths):
    items = _items(lengths, seed=3 + len(lengths))
    ref = collate_reference(items)
    got = pipeline.collate_on_device(pipeline.pack_dialogues(items).to("cuda"))
    _same(got, ref)
    # a longer (global) pad length, as a data-parallel shard gets it
    S = min(110, max(lengths) + 7)
    got = pipeline.collate_on_device(pipeline.pack_dialogues(items).to("cuda"), seq_len=S)
    assert got.seq_len == S
    _same(synthetic.Batch(got.text[:ref.seq_len], got.visual[:ref.seq_len], got.acoustic[:ref.seq_len], got.qmask[:ref.seq_len],
                          got.umask[:, :ref.seq_len], got.label[:, :ref.seq_len], got.lengths), ref)
    assert float(got.text[ref.seq_len:].abs().sum()) == 0.0 and float(got.umask[:, ref.seq_len:].sum()) == 0.0
    with pytest.raises(ValueError):
        pipeline.collate_on_device(pipeline.pack_dialogues(items).to("cuda"), seq_len=max(lengths) - 1)


@pytest.mark.gpu
def test_prefetcher_yields_the_loader_batches_in_order_and_feeds_a_train_step():
    import gan_ffn_b200 as G
    from gan_ffn_b200 import train
    all_lengths = [[6, 3, 8, 2], [12, 12, 1, 7], [4, 9, 9, 9], [2, 2, 5, 30], [10, 1, 1, 3]]
    batches = [_items(ls, seed=50 + k) for k, ls in enumerate(all_lengths)]
    packed = [pipeline.pack_dialogues(it) for it in batches]
    seen = 0
    for k, got in enumerate(pipeline.DevicePrefetcher(packed, "cuda")):
        _same(got, collate_reference(batches[k]))
        seen += 1
    assert seen == len(batches)
    assert list(pipeline.DevicePrefetcher([], "cuda")) == []
    # the batches are what the trainers consume: one stage-2 step straight from the prefetcher
    nets, ffn = train.build_networks(device="cuda")
    cls = train.ClassifierTrainer(ffn, torch.tensor(synthetic.IEMOCAP_LOSS_WEIGHTS, device="cuda"))
    for got in pipeline.DevicePrefetcher(packed[:2], "cuda"):
        loss, pred, labels = cls.step(got, train=True)
        assert torch.isfinite(loss) and pred.shape == labels.shape
