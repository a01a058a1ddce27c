"""Device-side batch assembly: the reference's ``MusicDataset`` iterators over an HBM-resident corpus.

Reference: ``MusicDataset.get_iterator`` (model/data_utils.py:206-304), ``get_dis_iterator`` (:307-368) and
``eval_iterator`` (:370-434).  There the host walks ``batch_size`` column trackers per batch, copies token spans out of
per-sequence CPU tensors into ``[bptt, batch]`` LongTensors and ships them over PCIe (``data.to(device)``); here each
split is uploaded once (``DeviceSplit``: one int32 token array with every sequence's start token, offsets, lengths) and
a batch is one kernel launch (``tgan_batch_next``: tracker walk + gather; ``tgan_batch_gather`` for plans the host owns:
the closed-form eval plan and the dis iterator's ``np.random.randint`` offsets, whose draw order is part of the
reference's behaviour).  The iterators yield exactly what the reference's yield -- same tensors, same ``reset_mem``,
same ``batch_token_num`` (a Python int: one 4-byte device read per batch, as train.py consumes it on the host,
train.py:870-906) -- so ``train.py``'s loops run on them unchanged.  ``TRAIN.random_crop`` / ``append_note_status`` are
off in every shipped config and raise here.  CUDA only: there is no host fallback.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import lib as L


class DeviceSplit:
    """One split of the corpus in HBM.  ``seqs``: sequences WITH their start token (data_utils.py:121-141)."""

    def __init__(self, seqs: Sequence, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.TganError("device-side batch assembly runs on CUDA only (no host fallback)")
        arrs = [np.asarray(s.cpu() if isinstance(s, torch.Tensor) else s, dtype=np.int64).reshape(-1) for s in seqs]
        self.n_seq = len(arrs)
        self.lens = np.array([len(a) for a in arrs], dtype=np.int32)
        self.offs = np.zeros(self.n_seq, dtype=np.int64)
        if self.n_seq > 1:
            self.offs[1:] = np.cumsum(self.lens[:-1].astype(np.int64))
        flat = np.concatenate(arrs + [np.zeros(1, dtype=np.int64)]) if arrs else np.zeros(1, dtype=np.int64)
        if flat.max(initial=0) >= 2 ** 31 or flat.min(initial=0) < 0:
            raise L.TganError("token ids must fit int32")
        self.corpus = torch.from_numpy(flat.astype(np.int32)).to(self.device)  # (+1 guard token: target reads p + 1)
        self.seq_off = torch.from_numpy(self.offs).to(self.device)
        self.seq_len = torch.from_numpy(self.lens).to(self.device)


class DeviceMusicDataset:
    """The iterator surface of the reference's ``MusicDataset`` over HBM-resident splits.

    ``DeviceMusicDataset.from_reference(dataset, device)`` wraps an already loaded reference ``MusicDataset`` (its
    ``train_data`` / ``valid_data`` / ``test_data`` lists and ``vocab``); the constructor takes the splits directly."""

    def __init__(self, splits: dict, pad_id: int, device, random_crop: bool = False, append_note_status: bool = False):
        if random_crop or append_note_status:
            raise NotImplementedError("random_crop / append_note_status are off in every shipped config and not "
                                      "part of the device-side iterators")
        self.device = torch.device(device)
        self.pad_id = int(pad_id)
        self.splits = {k: DeviceSplit(v, device) for k, v in splits.items()}

    @classmethod
    def from_reference(cls, dataset, device):
        cfg = dataset.cfg
        return cls({"train": dataset.train_data, "valid": dataset.valid_data, "test": dataset.test_data},
                   dataset.vocab.pad_id, device, random_crop=cfg.TRAIN.random_crop,
                   append_note_status=cfg.TRAIN.append_note_status)

    def _split(self, split, allowed):
        if split not in allowed:
            raise NotImplementedError(split) if "train" in allowed else ValueError(split)
        return self.splits[split]

    # ------------------------------------------------------------------------------------------------ training
    def get_iterator(self, batch_size, bptt, device=None, split="train", do_shuffle=True, seed=None):
        """data_utils.py:206-304 -> a callable returning the generator of
        ``(data [bptt, B] int64, target, reset_mem [B] bool, batch_token_num, None)``."""
        sp = self._split(split, ("train", "valid", "test"))
        dev, pad = self.device, self.pad_id

        def iterator():
            perm = np.arange(sp.n_seq)
            if do_shuffle:
                rng = np.random.RandomState(seed)
                rng.shuffle(perm)
            assert batch_size < sp.n_seq
            perm_dev = torch.from_numpy(perm.astype(np.int32)).to(dev)
            fresh = torch.cat([torch.arange(batch_size, dtype=torch.int32), torch.zeros(batch_size, dtype=torch.int32),
                               torch.tensor([batch_size], dtype=torch.int32)]).to(dev)
            tracker = fresh.clone()
            ntok_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            ntok_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            while True:
                data = torch.empty(bptt, batch_size, dtype=torch.int64, device=dev)
                target = torch.empty(bptt, batch_size, dtype=torch.int64, device=dev)
                reset = torch.empty(batch_size, dtype=torch.uint8, device=dev)
                L.batch_next(sp.corpus, sp.seq_off, sp.seq_len, perm_dev, sp.n_seq, tracker, data, target, reset,
                             ntok_dev, bptt, batch_size, pad)
                ntok_host.copy_(ntok_dev, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                ntok = int(ntok_host[0])
                if ntok == 0:  # the permutation is used up (:285-293)
                    if not do_shuffle:
                        return
                    rng.shuffle(perm)
                    perm_dev.copy_(torch.from_numpy(perm.astype(np.int32)))
                    tracker.copy_(fresh)
                    continue
                yield data, target, reset.bool(), ntok, None

        return iterator

    # ------------------------------------------------------------------------------------------------ discriminator
    def get_dis_iterator(self, batch_size, bptt, device=None, split="train", do_shuffle=True, seed=None):
        """data_utils.py:307-368 -> generator of ``(data [bptt, B] int64, batch_token_num)``.  The chunk offsets come
        from the global ``np.random.randint`` in the reference's column order (:349), so the (tiny) plan is made on the
        host exactly as the reference makes it; the token gather runs on the device."""
        sp = self._split(split, ("train", "valid", "test"))
        dev, pad = self.device, self.pad_id

        def iterator():
            perm = np.arange(sp.n_seq)
            if do_shuffle:
                rng = np.random.RandomState(seed)
                rng.shuffle(perm)
            assert batch_size < sp.n_seq
            tracker = [(i, 0) for i in range(batch_size)]
            next_idx = batch_size
            src = torch.zeros(batch_size, dtype=torch.int64).pin_memory()
            n_new = torch.zeros(batch_size, dtype=torch.int32).pin_memory()
            while True:
                src.fill_(-1)
                n_new.zero_()
                ntok = 0
                for i in range(batch_size):
                    idx, pos = tracker[i]
                    while idx < sp.n_seq:
                        sid = perm[idx]
                        length = int(sp.lens[sid])
                        if bptt > length:
                            idx, pos = next_idx, 0
                            tracker[i] = (idx, pos)
                            next_idx += 1
                            continue
                        pos = np.random.randint(0, length - bptt + 1)
                        src[i] = int(sp.offs[sid]) + pos
                        n_new[i] = bptt
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
                data = torch.empty(bptt, batch_size, dtype=torch.int64, device=dev)
                src_d, n_d = src.to(dev, non_blocking=True), n_new.to(dev, non_blocking=True)
                L.batch_gather(sp.corpus, src_d, n_d, data, None, bptt, batch_size, pad)
                torch.cuda.current_stream().synchronize()  # the pinned plan buffers are rewritten for the next batch
                yield data, ntok

        return iterator

    # ------------------------------------------------------------------------------------------------ evaluation
    def eval_iterator(self, batch_size, bptt, device=None, split="valid", local_rank=0, world_size=0):
        """data_utils.py:370-434 -> generator of ``(data, target, reset_all_mem, batch_token_num, None)``; the plan is a
        closed form of (batch_begin, seq_begin), the gather runs on the device."""
        sp = self._split(split, ("valid", "test"))
        dev, pad = self.device, self.pad_id
        lo, hi = 0, sp.n_seq
        if world_size > 0:
            lo = sp.n_seq // world_size * local_rank
            hi = sp.n_seq if local_rank == world_size - 1 else sp.n_seq // world_size * (local_rank + 1)
        lens, offs = sp.lens[lo:hi].astype(np.int64), sp.offs[lo:hi]
        total = hi - lo

        def iterator():
            for bb in range(0, total, batch_size):
                reset_all = True
                be = min(bb + batch_size, total)
                cl, co = lens[bb:be], offs[bb:be]
                for sb in range(0, int(cl.max()) - 1, bptt):
                    n_new = np.zeros(batch_size, dtype=np.int32)
                    src = np.full(batch_size, -1, dtype=np.int64)
                    live = cl > sb + 1
                    n_new[:be - bb] = np.where(live, np.minimum(sb + bptt, cl - 1) - sb, 0)
                    src[:be - bb] = np.where(live, co + sb, -1)
                    data = torch.empty(bptt, batch_size, dtype=torch.int64, device=dev)
                    target = torch.empty(bptt, batch_size, dtype=torch.int64, device=dev)
                    L.batch_gather(sp.corpus, torch.from_numpy(src).to(dev), torch.from_numpy(n_new).to(dev), data, target,
                                   bptt, batch_size, pad)
                    yield data, target, reset_all, int(n_new.sum()), None
                    reset_all = False

        return iterator
