"""Host-side mirror of one iteration of the reference training loop (train_gan.py:141-179):
warm-up LR, `set_grad` toggling, D step (G forward without grad, loss_D_fn, backward, Adam), G step
(G forward, loss_G_fn through the frozen D, backward, Adam).  Used by bench.py, smoke() and the parity
tests; the reference's own `train_gan.py` drives the same plug-ins through contrad_b200.dropin."""
import torch


def update_warmup(optimizer, cur_step, warmup, lr):
    """train_gan.py:88-93."""
    if warmup > 0:
        lr_w = min(1., (cur_step + 1) / warmup) * lr
        for group in optimizer.param_groups:
            group["lr"] = lr_w


def set_grad(model, flag=True):
    """utils.py:125-127."""
    for p in model.parameters():
        p.requires_grad = flag


def sample_generator(G, num_samples, enable_grad=True):
    """train_gan.py:96-100."""
    latent = G.sample_latent(num_samples)
    with torch.set_grad_enabled(enable_grad):
        return G(latent)


def grad_norm(model):
    sq = [p.grad.double().pow(2).sum() for p in model.parameters() if p.grad is not None]
    return torch.stack(sq).sum().sqrt() if sq else torch.zeros((), dtype=torch.float64)


def train_step(P, opt, train_fn, models, optimizers, images, step, use_warmup=True, record_grad_norms=False):
    """One step (n_critic = 1).  Returns a dict of 0-dim tensors (no host sync inside)."""
    generator, discriminator = models
    opt_G, opt_D = optimizers
    generator.train()
    discriminator.train()
    if use_warmup:
        update_warmup(opt_G, step, opt["warmup"], opt["lr"])
        update_warmup(opt_D, step, opt["warmup"], opt.get("lr_d", opt["lr"]))
    out = {}
    set_grad(generator, False)
    set_grad(discriminator, True)
    gen_images = sample_generator(generator, images.size(0), enable_grad=False)
    d_loss, aux = train_fn["D"](P, discriminator, opt, images, gen_images)
    loss = d_loss + aux["penalty"]
    opt_D.zero_grad()
    loss.backward()
    if record_grad_norms:
        out["d_grad_norm"] = grad_norm(discriminator)
    opt_D.step()
    out.update(d_loss=d_loss.detach(), d_penalty=aux["penalty"].detach(), d_real=aux["d_real"].detach(),
               d_gen=aux["d_gen"].detach())

    set_grad(generator, True)
    set_grad(discriminator, False)
    gen_images = sample_generator(generator, images.size(0))
    g_loss = train_fn["G"](P, discriminator, opt, images, gen_images)
    opt_G.zero_grad()
    g_loss.backward()
    if record_grad_norms:
        out["g_grad_norm"] = grad_norm(generator)
    opt_G.step()
    out["g_loss"] = g_loss.detach()
    return out
