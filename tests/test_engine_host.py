"""CPU: host logic of the engine -- the tiling of the K x U pair grid into bounded blocks (what streams a job
through a fixed activation buffer, replacing the reference's groups-of-50 loop, model.py:769-780) and the
frame arithmetic of the variants."""
import pytest
import torch
from hypothesis import given, settings, strategies as st

from enhance_cb_whisper_b200.engine import KWSEngine, PackedWeights


def engine(variant="LE"):
    return KWSEngine(PackedWeights(variant=variant, C=2, D=64, P=64))


@settings(max_examples=200, deadline=None)
@given(K=st.integers(1, 300), U=st.integers(1, 40), max_pairs=st.integers(1, 700))
def test_pair_chunks_cover_every_pair_exactly_once(K, U, max_pairs):
    seen = torch.zeros(K, U, dtype=torch.int32)
    for k0, k1, u0, u1 in engine().pair_chunks(K, U, 150, 1500, max_pairs):
        assert 0 <= k0 < k1 <= K and 0 <= u0 < u1 <= U
        assert (k1 - k0) * (u1 - u0) <= max(max_pairs, 1)
        seen[k0:k1, u0:u1] += 1
    assert int(seen.min()) == 1 and int(seen.max()) == 1


def test_pair_chunks_are_keyword_major():
    """A block is a run of keywords against a few utterances: the utterance tiles (the larger operand) stay
    L2-resident while keyword rows stream."""
    blocks = list(engine().pair_chunks(1000, 256, 150, 1500, 1184))
    assert all(u1 - u0 == 1 for _, _, u0, u1 in blocks)  # 1184 pairs < 1000 keywords x 2 utterances
    assert blocks[0] == (0, 1000, 0, 1)
    blocks = list(engine().pair_chunks(100, 64, 150, 1500, 1184))
    assert blocks[0] == (0, 100, 0, 11)


@pytest.mark.parametrize("variant,T,exp", [("L", 150, 150), ("LE", 1500, 1500), ("LEF", 150, 75), ("LEF", 151, 76),
                                           ("LEF", 1, 1)])
def test_out_frames(variant, T, exp):
    assert engine(variant).out_frames(T) == exp


def test_compress_rejects_wrong_width_and_layer_count():
    from enhance_cb_whisper_b200 import ops

    e = engine("L")
    with pytest.raises(ops.KWSError):
        e.compress(torch.zeros(1, 2, 4, 32), None)  # D != model embedding_dim
    with pytest.raises(ops.KWSError):
        e.compress(torch.zeros(1, 2, 4, 64), None, layer_idx=[0])  # needs C indices


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(min_value=0, max_value=150), min_size=0, max_size=300), st.integers(min_value=1, max_value=8))
def test_length_balanced_shards_properties(lens, world):
    """SURVEY 8e load balance: a partition of the vocabulary, shard sizes within one keyword, total valid frames
    within the longest keyword of each other, ascending ids inside a shard (the top-k tie-break order survives)."""
    import torch

    from enhance_cb_whisper_b200 import parallel

    lt = torch.tensor(lens, dtype=torch.int64)
    shards = parallel.length_balanced_shards(lt, world)
    assert len(shards) == world
    allids = torch.cat(shards) if shards else torch.empty(0, dtype=torch.int64)
    assert sorted(allids.tolist()) == list(range(len(lens)))
    sizes = [int(s.numel()) for s in shards]
    assert max(sizes) - min(sizes) <= 1
    for s in shards:
        assert s.tolist() == sorted(s.tolist())
    if lens:
        tot = [int(lt[s].sum()) for s in shards]
        assert max(tot) - min(tot) <= max(lens)
