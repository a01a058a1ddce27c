"""CPU restatement of the reference's batch iterators (numpy).  TEST INFRASTRUCTURE ONLY -- imported by tests/ as the
checker of the device-side batch assembly (tgan_batch_next / tgan_batch_gather), never by the product path.

Follows ``MusicDataset.get_iterator`` (model/data_utils.py:206-304), ``get_dis_iterator`` (:307-368) and
``eval_iterator`` (:370-434) with ``TRAIN.random_crop`` and ``TRAIN.append_note_status`` off (every shipped config).
``seqs`` are the split's sequences WITH their start token (data_utils.py:121-141).

Parity status: **pinned** -- ``tests/golden/batches_tiny.npz`` holds the batches the UNMODIFIED reference iterators
produce on a seeded ragged corpus (``oracle/make_goldens.py::run_batches_case``); ``tests/test_data_golden.py``
checks this restatement against every one of them.
"""
from __future__ import annotations

import numpy as np


def train_iterator(seqs, pad_id, batch_size, bptt, do_shuffle=True, seed=None, max_batches=None):
    """Yields (data [bptt, B] int64, target, reset_mem [B] bool, batch_token_num) -- data_utils.py:228-302."""
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    total = len(seqs)
    perm = np.arange(total)
    if do_shuffle:
        rng = np.random.RandomState(seed)                                         # :231-233
        rng.shuffle(perm)
    assert batch_size < total                                                     # :234
    tracker = [(i, 0) for i in range(batch_size)]                                 # :235
    next_idx = batch_size
    produced = 0
    while max_batches is None or produced < max_batches:
        data = np.full((bptt, batch_size), pad_id, dtype=np.int64)                # :246-247
        target = np.full((bptt, batch_size), pad_id, dtype=np.int64)
        reset = np.zeros(batch_size, dtype=bool)
        ntok = 0
        for i in range(batch_size):
            idx, pos = tracker[i]
            while idx < total:
                sid = perm[idx]
                L = lens[sid]
                if pos + 1 >= L:                                                  # :256-262
                    idx, pos = next_idx, 0
                    tracker[i] = (idx, pos)
                    next_idx += 1
                    reset[i] = True
                    continue
                n_new = min(L - 1 - pos, bptt)                                    # :272-277
                data[:n_new, i] = seqs[sid][pos:pos + n_new]
                target[:n_new, i] = seqs[sid][pos + 1:pos + 1 + n_new]
                ntok += n_new
                tracker[i] = (idx, pos + n_new)
                break
        if ntok == 0:                                                             # :285-293
            if not do_shuffle:
                return
            rng.shuffle(perm)
            tracker = [(i, 0) for i in range(batch_size)]
            next_idx = batch_size
            continue
        produced += 1
        yield data, target, reset, int(ntok)


def dis_iterator(seqs, pad_id, batch_size, bptt, do_shuffle=True, seed=None, max_batches=None, randint=None):
    """Yields (data [bptt, B] int64, batch_token_num) -- data_utils.py:325-366.  ``randint(lo, hi)`` stands for the
    global ``np.random.randint`` the reference draws the chunk offsets from (:349)."""
    randint = randint or np.random.randint
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    total = len(seqs)
    perm = np.arange(total)
    if do_shuffle:
        rng = np.random.RandomState(seed)
        rng.shuffle(perm)
    assert batch_size < total
    tracker = [(i, 0) for i in range(batch_size)]
    next_idx = batch_size
    produced = 0
    while max_batches is None or produced < max_batches:
        data = np.full((bptt, batch_size), pad_id, dtype=np.int64)
        ntok = 0
        for i in range(batch_size):
            idx, pos = tracker[i]
            while idx < total:
                sid = perm[idx]
                L = lens[sid]
                if bptt > L:                                                      # :343-347
                    idx, pos = next_idx, 0
                    tracker[i] = (idx, pos)
                    next_idx += 1
                    continue
                pos = randint(0, L - bptt + 1)                                    # :349-353
                data[:bptt, i] = seqs[sid][pos:pos + bptt]
                ntok += bptt
                tracker[i] = (idx, pos + bptt)
                break
        if ntok == 0:
            if not do_shuffle:
                return
            rng.shuffle(perm)
            tracker = [(i, 0) for i in range(batch_size)]
            next_idx = batch_size
            continue
        produced += 1
        yield data, int(ntok)


def eval_iterator(seqs, pad_id, batch_size, bptt, local_rank=0, world_size=0):
    """Yields (data, target, reset_all_mem, batch_token_num) -- data_utils.py:370-432."""
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    if world_size > 0:                                                            # :382-391
        n = len(seqs)
        b = n // world_size * local_rank
        e = n if local_rank == world_size - 1 else n // world_size * (local_rank + 1)
        seqs, lens = seqs[b:e], lens[b:e]
    total = len(seqs)
    for bb in range(0, total, batch_size):
        reset_all = True
        be = min(bb + batch_size, total)
        for sb in range(0, int(max(lens[bb:be])) - 1, bptt):                       # :402-405
            data = np.full((bptt, batch_size), pad_id, dtype=np.int64)
            target = np.full((bptt, batch_size), pad_id, dtype=np.int64)
            ntok = 0
            for i in range(bb, be):
                if lens[i] > sb + 1:
                    n_new = min(sb + bptt, lens[i] - 1) - sb                      # :410-413
                    data[:n_new, i - bb] = seqs[i][sb:sb + n_new]
                    target[:n_new, i - bb] = seqs[i][sb + 1:sb + n_new + 1]
                    ntok += n_new
            yield data, target, reset_all, int(ntok)
            reset_all = False


def ragged_corpus(seed, n_seq, vocab=310, lo=1, hi=90):
    """Seeded corpus of ragged sequences WITH their start token 0 (<S>), including degenerate lengths 1, 2 and 3 (the
    fixture generator and the tests rebuild the same corpus from the seed)."""
    rng = np.random.RandomState(seed)
    lens = rng.randint(lo, hi, size=n_seq)
    lens[:4] = [1, 2, 3, hi + 20]
    return [np.concatenate([[0], rng.randint(2, vocab, size=int(L) - 1)]).astype(np.int64) for L in lens]
