"""CPU: the oracle's restatement of the reference's batch iterators (oracle/data_oracle.py) against the batches the
UNMODIFIED reference MusicDataset iterators produced (tests/golden/batches_tiny.npz, oracle/make_goldens.py)."""
import os

import numpy as np

import data_oracle as DO
from golden_util import GOLD


def load():
    z = np.load(os.path.join(GOLD, "batches_tiny.npz"))
    seqs = DO.ragged_corpus(int(z["seed"]), int(z["n_seq"]))
    return z, seqs, int(z["B"]), int(z["bptt"]), int(z["pad_id"])


def check(z, tag, it, has_target=True):
    n = int(z[f"{tag}.n"])
    k = 0
    for item in it:
        if k >= n:
            break
        assert np.array_equal(item[0], z[f"{tag}.data{k}"]), (tag, k)
        if has_target:
            assert np.array_equal(item[1], z[f"{tag}.target{k}"]), (tag, k)
            assert np.array_equal(np.asarray(item[2]), z[f"{tag}.reset{k}"]), (tag, k)
            assert item[3] == int(z[f"{tag}.ntok{k}"]), (tag, k)
        else:
            assert item[1] == int(z[f"{tag}.ntok{k}"]), (tag, k)
        k += 1
    assert k == n, (tag, k, n)


def test_oracle_train_iterator_matches_reference():
    z, seqs, B, bptt, pad = load()
    check(z, "train", DO.train_iterator(seqs, pad, B, bptt, True, seed=7))
    once = list(DO.train_iterator(seqs, pad, B, bptt, False))
    assert len(once) == int(z["once.n"])  # the one-pass iterator ends by itself
    check(z, "once", iter(once))


def test_oracle_eval_iterator_matches_reference():
    z, seqs, B, bptt, pad = load()
    ev = list(DO.eval_iterator(seqs, pad, B, bptt))
    assert len(ev) == int(z["eval.n"])
    check(z, "eval", iter(ev))
    ev1 = list(DO.eval_iterator(seqs, pad, B, bptt, local_rank=1, world_size=2))
    assert len(ev1) == int(z["eval_r1.n"])
    check(z, "eval_r1", iter(ev1))


def test_oracle_dis_iterator_matches_reference():
    z, seqs, B, bptt, pad = load()
    np.random.seed(99)
    check(z, "dis", DO.dis_iterator(seqs, pad, B, bptt, True, seed=5), has_target=False)
