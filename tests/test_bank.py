"""Keyword-bank loader: on-disk *.bin format, padding / layer-selection semantics (CPU) and the streamed
compression into the resident operand bank (GPU)."""
import os

import pytest
import torch

from oracle import kws_oracle as O


def _write_bins(folder, tensors):
    width = len(str(len(tensors) - 1))
    for i, t in enumerate(tensors):
        if t is not None:
            with open(os.path.join(folder, str(i).zfill(width) + ".bin"), "wb") as f:
                torch.save(t.clone(), f)  # src/utils.py:199-201


def _ragged(n, layers, D, seed, lens):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.functional.normalize(torch.randn(layers, t, D, generator=g), dim=-1) for t in lens[:n]]


def test_iter_bin_dir_order_and_ghosts(tmp_path, built_lib):
    from enhance_cb_whisper_b200 import bank

    items = _ragged(4, 12, 16, 1, [5, 9, 3, 7])
    items[2] = None  # ghost keyword: no file (dataset.py:700-729)
    _write_bins(str(tmp_path), items)
    got = list(bank.iter_bin_dir(str(tmp_path), n_items=4))
    assert [g is None for g in got] == [False, False, True, False]
    for a, b in zip(got, items):
        if b is not None:
            assert torch.equal(a, b)
    assert len(list(bank.iter_bin_dir(str(tmp_path)))) == 4  # highest index present + 1


def test_pad_item_matches_reference_semantics(built_lib):
    from enhance_cb_whisper_b200 import bank

    hs = _ragged(1, 12, 8, 2, [30])[0]
    for n_frames in (22, 30, 41):
        got, T = bank.pad_item(hs, n_frames, 3)
        exp, mask = O.pad_frames(O.select_layers(hs, 3), n_frames)
        assert torch.equal(got, exp)
        assert T == int(mask[0].sum())


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["L", "LE", "LEF"])
def test_streamed_bank_equals_padded_batch(built_lib, cuda_dev, variant):
    """build_keyword_bank (ragged items, chunks of 3, a ghost) == compress() of the padded batch, bit for bit;
    score_bank == score on the padded tensors."""
    import enhance_cb_whisper_b200 as kb
    from enhance_cb_whisper_b200 import bank

    torch.manual_seed(3)
    C, D, Tk, Tu = 3, 128, 22, 70
    m = kb.KWSModelB200(n_layers=C, embedding_dim=D, proj_mlp_units=64, learn_features=variant != "L",
                        proj_mlp=variant != "L", frames_conv=variant == "LEF", resnet_version="resnet-18",
                        features_size=(Tk, Tu)).to(cuda_dev).eval()
    items = _ragged(7, 12, D, 4, [5, 22, 30, 9, 1, 17, 12])
    items[3] = None
    b = bank.build_keyword_bank(m, items, Tk, cuda_dev, chunk=3)
    assert b.K == 7 and b.hotword_mask.tolist() == [1, 1, 1, 0, 1, 1, 1]
    assert b.lengths.tolist() == [5, 22, 22, 0, 1, 17, 12]
    padded = torch.stack([bank.pad_item(t if t is not None else torch.zeros(12, 1, D), Tk, C)[0] for t in items])
    mask = (torch.arange(Tk)[None] < b.lengths.cpu()[:, None]).float()
    if variant == "LEF":
        mask = O.pooled_mask(mask)
    mask = mask[:, None, :].expand(-1, C, -1).contiguous()
    eng = m.prepare(cuda_dev)
    ref = eng.compress(padded.to(cuda_dev), mask.to(cuda_dev), list(range(C)))
    assert torch.equal(b.kwd_n, ref)
    g = torch.Generator().manual_seed(9)
    utt = torch.nn.functional.normalize(torch.randn(2, C, Tu, D, generator=g), dim=-1)
    um = torch.ones(2, C, (Tu + 1) // 2 if variant == "LEF" else Tu)
    sc, det, lg = bank.score_bank(m, b, utt.to(cuda_dev), um.to(cuda_dev))
    sc2, det2, lg2 = m.score(padded.to(cuda_dev), utt.to(cuda_dev), mask.to(cuda_dev), um.to(cuda_dev),
                             hotword_mask=b.hotword_mask)
    assert torch.equal(lg, lg2) and torch.equal(sc, sc2) and torch.equal(det, det2)
    assert float(sc[3].abs().max()) == 0.0  # ghost keyword scores 0
