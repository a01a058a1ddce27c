"""Times the GAN phases of experiment_spanbert.yml (transformer_gan.py:232-533) on one GPU: one "dis_loss" call and one
"gen_loss" call of the drop-in TransformerGAN on a [128, B] batch of synthetic MAESTRO-vocab tokens
(123 Gumbel-softmax sampling steps, BERT 5x768 discriminator with seeded random weights, WGAN-GP).
Usage: python tools/gan_bench.py [B] [reps] [graphs 0/1] [phases, e.g. gen_loss]"""
import json, os, sys, tempfile, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transformer-gan_b200"))
import torch


class Vocab:
    vec_len = 0
    def __len__(self):
        return 310


def make_cfg(bert_dir, batch_chunk=1):
    ns = types.SimpleNamespace
    return ns(MODEL=ns(num_layers=6, num_heads=10, units=500, inner_size=1000, dropout=0.1, attention_dropout=0.1,
                       tie_embedding=True, tie_proj=False, pre_lnorm=False, same_length=False, clamp_len=-1),
              TRAIN=ns(tgt_length=128, mem_length=1024, pad_type="model", replace_start_with_pad=False,
                       append_note_status=False),
              DISCRIMINATOR=ns(type="bert", tgt_len=128, mem_len=128, context_len=5, sample_chunks_mem=2,
                               truncate_backprop=False, backprop_outside=True, gen_loss_factor=1.0, dis_loss_factor=1.0,
                               batch_chunk=batch_chunk,
                               BERT=ns(model_path=bert_dir, loss_type="wgan-gp", model_type="bert_lm", random_weights=True,
                                       freeze_layers=["0", "1", "2", "3", "4"]),
                               CNN=ns(embed_dim=64, hidden_dim=64, num_rep=64, init="uniform", loss_type="rsgan")),
              PPO=ns(dis_D_type="bert", dis_D_num_rep=1, clip_param=0.4))


def build(device, seed=0):
    import transformer_gan as TG
    d = tempfile.mkdtemp(prefix="tgan_bert_")
    json.dump(dict(vocab_size=311, hidden_size=768, num_hidden_layers=5, num_attention_heads=12, intermediate_size=3072,
                   max_position_embeddings=512, type_vocab_size=2, hidden_act="gelu", hidden_dropout_prob=0.1,
                   attention_probs_dropout_prob=0.1, layer_norm_eps=1e-12, model_type="bert"),
              open(os.path.join(d, "config.json"), "w"))
    torch.manual_seed(seed)
    model = TG.TransformerGAN(make_cfg(d), Vocab())
    g = torch.Generator().manual_seed(1111)
    for name, p in model.generator.named_parameters():
        if name.endswith("layer_norm.weight"):
            p.data.copy_(1.0 + 0.01 * torch.randn(p.shape, generator=g))
        elif name.endswith("bias") and "r_" not in name:
            p.data.zero_()
        else:
            p.data.copy_(0.01 * torch.randn(p.shape, generator=g))
    return model.to(device).train()


def time_phase(model, data, phase, reps, warm=1):
    ts = []
    for _ in range(reps + warm):
        model.zero_grad(set_to_none=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model(data, None, None, phase)
        v = float(out[phase])
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts = ts[warm:]
    return sorted(ts)[len(ts) // 2], v


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    graphs = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
    from tgan_b200 import lib as L
    dev = torch.device("cuda", 0)
    model = build(dev)
    model.temperature = 1.0
    model.use_cuda_graphs = graphs
    g = torch.Generator().manual_seed(7)
    data = torch.randint(2, 310, (128, B), generator=g).to(dev)
    phases = sys.argv[4].split(",") if len(sys.argv) > 4 else ("dis_loss", "gen_loss")
    for phase in phases:
        n0 = L.launch_count()
        warm = 2 if graphs else 1  # graphs: first call eager (lazy init), second call captures
        t, v = time_phase(model, data, phase, reps, warm)
        n = (L.launch_count() - n0) // (reps + warm)
        print(f"{phase}: graphs={int(graphs)} B={B}  {t*1e3:9.1f} ms per call  ({B / t:8.1f} sequences/s, {n} library launches, value {v:.4f})", flush=True)
