"""Loss / schedule helpers of the GAN step with the reference's names (model/utils/helpers.py:62-141).

Only the functions ``transformer_gan.py`` and ``train.py`` import for the adversarial phase are provided:
``get_fixed_temperature`` (Gumbel temperature schedule, helpers.py:62-82) and ``get_losses`` (adversarial objectives,
helpers.py:85-141).  They act on a handful of scalars / ``[batch]`` vectors of discriminator logits, so they stay
host-side torch expressions; the heavy arithmetic of the path is in the CUDA library.
"""
import math

import torch
import torch.nn.functional as F


def get_fixed_temperature(temper, i, N, adapt):
    """Inverse-temperature schedule beta(i) in [1, temper]; the caller uses tau = 1 / beta (train.py:862-868)."""
    if adapt == "no":
        return 1.0
    if adapt == "lin":
        return 1 + i / (N - 1) * (temper - 1)
    if adapt == "exp":
        return temper ** (i / N)
    if adapt == "log":
        return 1 + (temper - 1) / math.log(N) * math.log(i + 1)
    if adapt == "sigmoid":
        return (temper - 1) / (1 + math.exp((N / 2 - i) * 20 / N)) + 1
    if adapt == "quad":
        return (temper - 1) / (N - 1) ** 2 * i ** 2 + 1
    if adapt == "sqrt":
        return (temper - 1) / math.sqrt(N - 1) * math.sqrt(i) + 1
    raise ValueError(f"unknown temperature policy {adapt!r}")


def get_losses(d_out_real, d_out_fake, loss_type="JS"):
    """-> (g_loss, d_loss) for discriminator logits on real / generated samples (helpers.py:85-141)."""
    ones_r, ones_f, zeros_f = torch.ones_like(d_out_real), torch.ones_like(d_out_fake), torch.zeros_like(d_out_fake)
    bce = F.binary_cross_entropy_with_logits
    if loss_type in ("standard", "JS", "KL"):
        d_loss_fake = bce(d_out_fake, zeros_f)
        d_loss = bce(d_out_real, ones_r) + d_loss_fake
        if loss_type == "standard":
            g_loss = bce(d_out_fake, ones_f)
        elif loss_type == "JS":
            g_loss = -d_loss_fake
        else:
            g_loss = torch.mean(-d_out_fake)
    elif loss_type == "hinge":
        d_loss = torch.mean(F.relu(1.0 - d_out_real)) + torch.mean(F.relu(1.0 + d_out_fake))
        g_loss = -torch.mean(d_out_fake)
    elif "wgan" in loss_type:  # 'wgan' / 'wgan-gp'
        d_loss_fake = d_out_fake.mean()
        d_loss = -d_out_real.mean() + d_loss_fake
        g_loss = -d_loss_fake
    elif loss_type == "tv":
        d_loss = torch.mean(torch.tanh(d_out_fake) - torch.tanh(d_out_real))
        g_loss = torch.mean(-torch.tanh(d_out_fake))
    elif "rsgan" in loss_type:  # 'rsgan' / 'rsgan-gp'
        d_loss = bce(d_out_real - d_out_fake, ones_r)
        g_loss = bce(d_out_fake - d_out_real, ones_f)
    elif "ppo" in loss_type:  # 'ppo' / 'ppo-gp'
        with torch.no_grad():
            W = d_out_fake.shape[0] * F.softmax(d_out_fake.detach(), dim=0)
        d_loss = torch.mean(W * d_out_fake - d_out_real)
        g_loss = -torch.mean(d_out_fake)
    else:
        raise NotImplementedError(f"Divergence {loss_type!r} is not implemented")
    return g_loss, d_loss


def truncated_normal_(tensor, mean=0.0, std=1.0):
    """In-place N(mean, std) truncated to two standard deviations (helpers.py: used by CNN discriminator init)."""
    with torch.no_grad():
        torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=mean - 2 * std, b=mean + 2 * std)
    return tensor
