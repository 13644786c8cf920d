"""Dialogue-sharded inference scoring (BASELINE configs 3/4): the host-side plan on CPU, and on the GPU that dealing
whole batches to ranks reproduces the unsharded scores exactly (a batch keeps its pad length wherever it is scored)
while re-batching a dialogue with a different pad length does change its output."""
import pytest
import torch

from gan_ffn_b200 import scoring, synthetic


def test_plan_and_shards_cover_every_dialogue_once():
    lengths = synthetic.ragged_lengths(101, 10, 110, seed=5)
    for sort in (True, False):
        plan = scoring.plan_batches(lengths, batch_size=8, sort=sort)
        assert sorted(i for b in plan for i in b) == list(range(101))
        assert all(len(b) == 8 for b in plan[:-1]) and len(plan[-1]) == 101 % 8
        for world in (1, 2, 3, 8):
            shards = [scoring.shard_batches(plan, world, r) for r in range(world)]
            assert sorted(i for s in shards for b in s for i in b) == list(range(101))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    sorted_slots = scoring.padded_slots(scoring.plan_batches(lengths, 8, True), lengths)
    loader_slots = scoring.padded_slots(scoring.plan_batches(lengths, 8, False), lengths)
    assert sum(lengths) <= sorted_slots < loader_slots


def test_synthetic_corpus_is_addressable_by_index():
    corpus = scoring.SyntheticDialogues([5, 9, 3, 7], n_classes=7)
    a = corpus.batch([1, 2])
    b = corpus.batch([2, 3, 1])
    assert a.seq_len == 9 and b.seq_len == 9 and a.lengths == [9, 3]
    assert torch.equal(a.text[:, 0], b.text[:, 2]) and torch.equal(a.visual[:3, 1], b.visual[:3, 0])
    assert torch.count_nonzero(a.text[3:, 1]) == 0 and a.umask[1].tolist() == [1.0] * 3 + [0.0] * 6
    assert int(a.label.max()) < 7


@pytest.mark.gpu
def test_sharded_scoring_equals_unsharded_scoring():
    from gan_ffn_b200 import train
    nets, ffn = train.build_networks(n_classes=7, device="cuda")
    lengths = synthetic.ragged_lengths(12, 4, 33, seed=9)                 # MELD-shaped: short dialogues, 7 classes
    corpus = scoring.SyntheticDialogues(lengths, n_classes=7)
    plan = scoring.plan_batches(lengths, batch_size=4)
    ref = {}
    for b in plan:
        out = scoring.score_batch(ffn, corpus.batch(b).to("cuda"))
        for k, i in enumerate(b):
            ref[i] = out["log_prob"][:, k].clone()
    for world in (2, 3):
        got = {}
        for r in range(world):
            for b in scoring.shard_batches(plan, world, r):
                out = scoring.score_batch(ffn, corpus.batch(b).to("cuda"))
                for k, i in enumerate(b):
                    got[i] = out["log_prob"][:, k]
        assert sorted(got) == sorted(ref)
        for i in ref:
            assert torch.equal(got[i], ref[i]), f"dialogue {i} scored differently on a shard"
    # the pad length is part of the result: the same dialogue in a longer batch gives different numbers
    i_short, i_long = plan[0][0], plan[-1][-1]
    mixed = scoring.score_batch(ffn, corpus.batch([i_short, i_long]).to("cuda"))["log_prob"][:lengths[i_short], 0]
    assert not torch.allclose(mixed, ref[i_short][:lengths[i_short]], rtol=1e-3, atol=1e-4)


@pytest.mark.gpu
def test_graphed_scorer_equals_eager_scoring():
    from gan_ffn_b200 import train
    nets, ffn = train.build_networks(n_classes=6, device="cuda")
    lengths = [9, 9, 7, 9, 12, 12, 10, 12]
    corpus = scoring.SyntheticDialogues(lengths)
    scorer = scoring.GraphedScorer(ffn)
    # two shapes; the small one is recorded before the large one is first seen (the workspace grows in between: a
    # recorded graph must keep working), then both are revisited
    batches = [[0, 1, 2, 3], [3, 2, 1, 0], [2, 3, 0, 1], [4, 5, 6, 7], [7, 6, 5, 4], [0, 1, 2, 3], [4, 5, 6, 7]]
    for idx in batches:
        b = corpus.batch(idx).to("cuda")
        ref = scoring.score_batch(ffn, b)["log_prob"].clone()
        got = scorer(b)["log_prob"]
        assert torch.equal(got, ref), f"graph replay differs from eager scoring for batch {idx}"
    assert len(scorer._graphs) == 2
