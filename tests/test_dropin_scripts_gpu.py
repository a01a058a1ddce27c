"""GPU: the reference's UNMODIFIED train.py and generate.py run on top of the drop-in modules.

baseline/_ref/ (git-ignored; built by tools/vendor_reference.py in the build container and shipped to the GPU box with
the repo snapshot) holds a verbatim copy of the reference's model/ directory, a synthetic MAESTRO-vocabulary corpus in
the layout MusicDataset expects, a stand-in BERT checkpoint and small experiment files.  The scripts are started exactly
as their README does -- cwd = model/, `python train.py --data_dir .. --work_dir .. --cfg ..` -- with ONE addition:
PYTHONPATH=<repo>/transformer-gan_b200/compat, which resolves mem_transformer / transformer_gan / discriminator /
utils.proj_adaptive_softmax to this package and covers the environment gaps (yacs, nltk, transformers.AdamW).
What the run exercises: MusicDataset iterators -> TransformerGAN(...) under DistributedDataParallel -> 12 optimizer steps
with batch_chunk 2 (MLE every step through torch.optim.Adam / lamb.Lamb -- updates the engine must notice --,
discriminator + generator updates every 3rd step incl. WGAN-GP, requires_grad toggling of train.py:940-1000), evaluate()
with its same_length / length switch, checkpointing (pickled vocab), then generate.py loading that checkpoint and
decoding with its `debug: True` incremental == full memory self-check (atol 1e-4 -> that run uses the fp32 mode)."""
import glob
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
COMPAT = os.path.join(ROOT, "transformer-gan_b200", "compat")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "model", "train.py")),
                               reason="baseline/_ref not built (tools/vendor_reference.py needs /root/reference)")


def _env(port, **extra):
    env = dict(os.environ, PYTHONPATH=COMPAT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1",
               LOCAL_RANK="0", PYTHONUNBUFFERED="1")
    env.update(extra)
    return env


def _run(cmd, env, timeout=900):
    out = subprocess.run(cmd, cwd=os.path.join(REF, "model"), env=env, capture_output=True, text=True, timeout=timeout)
    return out.returncode, out.stdout, out.stderr


@needs_ref
@pytest.mark.parametrize("cfg_name", ["train_bert.yml", "train_cnn.yml"])
def test_train_py_then_generate_py_run_unmodified(cfg_name, tmp_path):
    work = str(tmp_path / "work")
    rc, so, se = _run([sys.executable, "train.py", "--data_dir", os.path.join(REF, "data"), "--work_dir", work, "--cfg",
                       os.path.join(REF, "cfg", cfg_name)], _env(29741, CUDA_VISIBLE_DEVICES="0"))
    assert rc == 0, (so[-3000:], se[-6000:])
    runs = glob.glob(os.path.join(work, "*"))
    assert len(runs) == 1
    ckpts = sorted(os.path.basename(p) for p in glob.glob(os.path.join(runs[0], "*.pt")))
    assert ckpts, os.listdir(runs[0])
    log = open(glob.glob(os.path.join(runs[0], "*.log"))[0]).read() if glob.glob(os.path.join(runs[0], "*.log")) else so + se
    assert "nan" not in log.lower().replace("nanoseconds", ""), log[-2000:]
    # ---- generate.py on that checkpoint; debug: True runs the script's own incremental == full-sequence memory check
    ckpt = "checkpoint_last.pt" if "checkpoint_last.pt" in ckpts else ckpts[0]
    inf = tmp_path / "inference.yml"
    inf.write_text(f"""
EVENT:
  vocab_file_path: '{os.path.join(REF, "data", "vocab.txt")}'
MODEL:
  model_directory: '{runs[0]}'
  memory_length: 100
  checkpoint_name: '{ckpt}'
  debug: True
SAMPLING:
  technique: 'topk'
  threshold: 32.0
  temperature: 0.95
INPUT:
  time_extension: False
  conditional_input_melody: 'Null'
  exclude_bos_token: True
  num_midi_files: 1
  num_empty_tokens_to_ignore: 0
OUTPUT:
  output_txt_directory: '{tmp_path / "out"}'
GENERATION:
  generation_length: 48
  duration_based: False
""")
    rc, so, se = _run([sys.executable, "generate.py", "--inference_config", str(inf)],
                      _env(29742, CUDA_VISIBLE_DEVICES="0", TGAN_B200_DTYPE="fp32"))
    assert rc == 0 and "Mem same" in so, (so[-3000:], se[-6000:])
    toks = open(tmp_path / "out" / "0.txt").read().split()
    assert len(toks) == 48


@needs_ref
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_train_py_under_two_rank_ddp(tmp_path):
    work = str(tmp_path / "work")
    env = dict(os.environ, PYTHONPATH=COMPAT, PYTHONUNBUFFERED="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29743", "train.py", "--data_dir", os.path.join(REF, "data"), "--work_dir", work, "--cfg",
           os.path.join(REF, "cfg", "train_bert.yml")]
    rc, so, se = _run(cmd, env)
    assert rc == 0, (so[-3000:], se[-6000:])
