"""GPU: device-side batch assembly (tgan_batch_next / tgan_batch_gather through tgan_b200.data.DeviceMusicDataset)
against the oracle's restatement of the reference iterators (oracle/data_oracle.py, pinned to the reference's own
batches by tests/test_data_golden.py) and against those golden batches directly.  Bit-exact: ids, pad fill, reset flags,
token counts, epoch wrap / reshuffle, the one-pass end, degenerate sequence lengths."""
import os

import numpy as np
import pytest
import torch

import data_oracle as DO
from golden_util import GOLD

pytestmark = pytest.mark.gpu


def _ds(seqs, pad):
    from tgan_b200.data import DeviceMusicDataset
    return DeviceMusicDataset({"train": seqs, "valid": seqs, "test": seqs}, pad, "cuda")


def _same(got, want, has_target=True):
    assert np.array_equal(got[0].cpu().numpy(), want[0])
    if has_target:
        assert np.array_equal(got[1].cpu().numpy(), want[1])
        g2 = got[2].cpu().numpy() if isinstance(got[2], torch.Tensor) else np.asarray(got[2])
        assert np.array_equal(g2, np.asarray(want[2]))
        assert got[3] == want[3]
    else:
        assert got[1] == want[1]


def test_device_iterators_reproduce_the_reference_golden_batches():
    z = np.load(os.path.join(GOLD, "batches_tiny.npz"))
    seqs = DO.ragged_corpus(int(z["seed"]), int(z["n_seq"]))
    B, bptt, pad = int(z["B"]), int(z["bptt"]), int(z["pad_id"])
    ds = _ds(seqs, pad)

    def check(tag, it, has_target=True):
        n, k = int(z[f"{tag}.n"]), 0
        for item in it:
            if k >= n:
                break
            want = (z[f"{tag}.data{k}"],) + ((z[f"{tag}.target{k}"], z[f"{tag}.reset{k}"], int(z[f"{tag}.ntok{k}"]))
                                             if has_target else (int(z[f"{tag}.ntok{k}"]),))
            _same(item, want, has_target)
            k += 1
        assert k == n, (tag, k, n)

    check("train", ds.get_iterator(B, bptt, "cuda", "train", True, seed=7)())
    once = list(ds.get_iterator(B, bptt, "cuda", "train", False)())
    assert len(once) == int(z["once.n"])
    check("once", iter(once))
    ev = list(ds.eval_iterator(B, bptt, "cuda", "valid")())
    assert len(ev) == int(z["eval.n"])
    check("eval", iter(ev))
    check("eval_r1", ds.eval_iterator(B, bptt, "cuda", "valid", local_rank=1, world_size=2)())
    np.random.seed(99)
    check("dis", ds.get_dis_iterator(B, bptt, "cuda", "train", True, seed=5)(), has_target=False)


def test_device_train_iterator_at_training_size_matches_oracle():
    """experiment_baseline shapes: 512 columns x 128 tokens over a 1500-sequence corpus, through an epoch wrap."""
    seqs = DO.ragged_corpus(3, 1500, lo=1, hi=400)
    pad, B, bptt = 1, 512, 128
    ds = _ds(seqs, pad)
    it_dev = ds.get_iterator(B, bptt, "cuda", "train", True, seed=11)()
    it_ref = DO.train_iterator(seqs, pad, B, bptt, True, seed=11)
    resets = 0
    for k in range(12):
        got, want = next(it_dev), next(it_ref)
        _same(got, want)
        resets += int(want[2].sum())
    assert resets > B  # the run crossed sequence boundaries (and the reshuffle) many times
    np.random.seed(5)
    it_dev = ds.get_dis_iterator(B, bptt, "cuda", "train", True, seed=2)()
    got = [next(it_dev) for _ in range(3)]
    np.random.seed(5)
    it_ref = DO.dis_iterator(seqs, pad, B, bptt, True, seed=2)
    for g in got:
        _same(g, next(it_ref), has_target=False)
