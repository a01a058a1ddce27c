"""Prepare baseline/_ref/ (git-ignored, but shipped to the GPU box by gpurun) for tests/test_dropin_scripts_gpu.py:

  baseline/_ref/model/      a verbatim copy of /root/reference/model (the UNMODIFIED train.py / generate.py and everything
                            they import); never committed -- the reference's sources stay out of the repo's history
  baseline/_ref/data/       a synthetic corpus in the layout MusicDataset expects (data_utils.py:101-175): vocab.txt =
                            the reference's performance_vocab.txt, {train,valid,test}/*.npy = equal-length int arrays of
                            MAESTRO-vocabulary ids
  baseline/_ref/bert/       a seed-initialised BertForMaskedLM checkpoint standing in for ../BERT/checkpoint-1969000
                            (not shipped by the reference, experiment_spanbert.yml:72)
  baseline/_ref/cfg/*.yml   small experiment / inference files for the reference's own yacs schemas

Run in the build container (where /root/reference exists): python tools/vendor_reference.py.  __graft_entry__.build() calls it."""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")

TRAIN_YML = """
MODEL:
  num_layers: 2
  num_heads: 4
  units: 64
  inner_size: 128
  dropout: 0.1
  attention_dropout: 0.1
TRAIN:
  load_from_previous: "Null"
  batch_size: 8
  batch_chunk: 2
  tgt_length: 32
  mem_length: 64
  lr: 0.002
  lr_min: 0.0001
  scheduler: inv_sqrt
  warmup_step: 4
  clip: 1.0
  max_step: 12
  log_interval: 4
  eval_interval: 6
  optim: {optim}
EVALUATE:
  batch_size: 1
  tgt_length: 32
  mem_length: 128
DATASET:
  trim_padding: True
DISCRIMINATOR:
  freeze_discriminator: False
  type: "{dtype}"
  batch_chunk: 2
  start_iter: 2
  dis_loss_freq: 3
  gen_loss_freq: 3
  tgt_len: 16
  mem_len: 16
  beta_max: 100.0
  dis_steps: 1
  gen_loss_factor: 1
  dis_loss_factor: 1
  sample_chunks_mem: 2
  context_len: 5
  adapt: 'exp'
  gen_lr: 0.002
  gen_scheduler: inv_sqrt
  gen_warmup_step: 4
  dis_lr: 0.002
  dis_scheduler: inv_sqrt
  dis_warmup_step: 4
  BERT:
    model_path: "{bert}"
    model_type: "bert_lm"
    freeze_layers: ['0', '1']
    loss_type: 'wgan-gp'
  CNN:
    loss_type: 'rsgan'
"""


def main():
    if not os.path.isdir(os.path.join(REF, "model")):
        print("vendor_reference: /root/reference not present, nothing to do")
        return 0
    os.makedirs(DST, exist_ok=True)
    model_dst = os.path.join(DST, "model")
    if os.path.isdir(model_dst):
        shutil.rmtree(model_dst)
    shutil.copytree(os.path.join(REF, "model"), model_dst, ignore=shutil.ignore_patterns("__pycache__"))
    # ---- synthetic corpus
    data = os.path.join(DST, "data")
    if os.path.isdir(data):
        shutil.rmtree(data)
    for split in ("train", "valid", "test"):
        os.makedirs(os.path.join(data, split))
    shutil.copy(os.path.join(REF, "data", "performance_vocab.txt"), os.path.join(data, "vocab.txt"))
    rng = np.random.RandomState(1111)
    for split, n in (("train", 24), ("valid", 4), ("test", 4)):
        for i in range(n):  # equal lengths: np.array(list of arrays) in load_cache_data must stay rectangular
            np.save(os.path.join(data, split, f"{i:03d}.npy"), rng.randint(2, 310, size=600).astype(np.int64))
    # ---- stand-in BERT checkpoint
    import torch
    from transformers import BertConfig, BertForMaskedLM
    torch.manual_seed(0)
    bert = os.path.join(DST, "bert")
    cfg = BertConfig(vocab_size=311, hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=128,
                     max_position_embeddings=64, type_vocab_size=2, hidden_act="gelu", layer_norm_eps=1e-12)
    BertForMaskedLM(cfg).save_pretrained(bert)
    # ---- configs
    cdir = os.path.join(DST, "cfg")
    os.makedirs(cdir, exist_ok=True)
    for name, dtype, optim in (("train_bert.yml", "bert", "adam"), ("train_cnn.yml", "cnn", "lamb")):
        open(os.path.join(cdir, name), "w").write(TRAIN_YML.format(dtype=dtype, optim=optim, bert=bert))
    print("vendor_reference: wrote", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
