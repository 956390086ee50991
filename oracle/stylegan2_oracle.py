"""TEST INFRASTRUCTURE - CPU oracle of the StyleGAN2 side of the ContraD hot path.

A plain-PyTorch fp32 *restatement* of the reference algorithm on explicit state_dicts (keys and shapes of the
reference modules), NCHW like the reference.  Every function cites the reference lines it follows.  Pinned on
fixtures produced by running the UNMODIFIED reference in the build container (tests/golden/make_golden_sg2.py ->
tests/golden/stylegan2_small.pt; replayed by tests/test_oracle_golden.py).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product never does.
"""
import math

import torch
import torch.nn.functional as F

SMALL32 = {4: 512, 8: 512, 16: 256, 32: 128}                          # discriminator.py:193-199 / generator.py:161-167


def channels_for(size, channel_multiplier=2, small32=False):
    """discriminator.py:193-211, generator.py:161-179."""
    if small32:
        return dict(SMALL32)
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: int(256 * channel_multiplier), 128: int(128 * channel_multiplier),
            256: int(64 * channel_multiplier), 512: int(32 * channel_multiplier), 1024: int(16 * channel_multiplier)}


# ------------------------------------------------------------------------------------------------ ops
def make_kernel(k=(1, 3, 3, 1)):
    """layers.py:23-31."""
    k = torch.tensor(k, dtype=torch.float32)
    k = k[None, :] * k[:, None]
    return k / k.sum()


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """op/upfirdn2d.py:159-200 restated: zero-stuffing, (possibly negative) padding, correlation with the flipped
    kernel, decimation.  x: [N,C,H,W]."""
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    p0, p1 = pad
    z = x.new_zeros(n, c, h * up, w * up)
    z[:, :, ::up, ::up] = x
    z = F.pad(z, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    z = z[:, :, max(-p0, 0):z.shape[2] - max(-p1, 0), max(-p0, 0):z.shape[3] - max(-p1, 0)]
    wk = torch.flip(kernel, [0, 1]).view(1, 1, kh, kw)
    out = F.conv2d(z.reshape(n * c, 1, z.shape[2], z.shape[3]), wk)
    out = out.reshape(n, c, out.shape[2], out.shape[3])
    return out[:, :, ::down, ::down]


def fused_lrelu(x, bias, slope=0.2, scale=2 ** 0.5):
    """op/fused_act.py:86-94."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    return F.leaky_relu(x + bias.view(*shape), slope) * scale


def equal_conv(x, weight, stride=1, padding=0):
    """layers.py:96-121 (bias=False in every ConvLayer)."""
    scale = 1 / math.sqrt(weight.shape[1] * weight.shape[2] ** 2)
    return F.conv2d(x, weight * scale, stride=stride, padding=padding)


def equal_linear(x, weight, bias, lr_mul=1.0, bias_init=0.0, activation=False):
    """layers.py:132-154."""
    scale = (1 / math.sqrt(weight.shape[1])) * lr_mul
    b = bias * lr_mul + bias_init
    if activation:
        return fused_lrelu(F.linear(x, weight * scale), b)
    return F.linear(x, weight * scale, b)


def conv_layer(sd, prefix, x, ksize, downsample=False, activate=True):
    """layers.py:174-198: [Blur] + EqualConv2d + [FusedLeakyReLU]; children indexed like the nn.Sequential."""
    i = 0
    if downsample:
        p = (4 - 2) + (ksize - 1)
        x = upfirdn2d(x, sd[prefix + ".0.kernel"], pad=((p + 1) // 2, p // 2))
        i = 1
        x = equal_conv(x, sd["%s.%d.weight" % (prefix, i)], stride=2, padding=0)
    else:
        x = equal_conv(x, sd["%s.%d.weight" % (prefix, i)], stride=1, padding=ksize // 2)
    if activate:
        x = fused_lrelu(x, sd["%s.%d.bias" % (prefix, i + 1)])
    return x


def minibatch_stddev(x, group=4):
    """discriminator.py:22-33."""
    b, c, h, w = x.shape
    g = min(b, group)
    s = x.view(g, -1, 1, c, h, w)
    s = torch.sqrt(s.var(0, unbiased=False) + 1e-8)
    s = s.mean([2, 3, 4], keepdim=True).mean(2)
    return torch.cat([x, s.repeat(g, 1, h, w)], 1)


# ------------------------------------------------------------------------------------------------ discriminator
def d_penultimate(sd, x, size):
    """ResidualDiscriminatorP.penultimate (discriminator.py:228-235): features [B, 8192] in (c,h,w) order."""
    out = x * 2.0 - 1.0
    out = conv_layer(sd, "layers.0", out, 1)                                       # FromRGB
    n_blocks = int(math.log(size, 2)) - 2
    for i in range(1, n_blocks + 1):                                               # ResBlock, discriminator.py:60-76
        p = "layers.%d" % i
        o = conv_layer(sd, p + ".conv1", out, 3)
        o = conv_layer(sd, p + ".conv2", o, 3, downsample=True)
        s = conv_layer(sd, p + ".skip", out, 1, downsample=True, activate=False)
        out = (o + s) / math.sqrt(2)
    out = minibatch_stddev(out)
    out = conv_layer(sd, "last_conv", out, 3)
    return out.reshape(out.shape[0], -1)


def _mlp(sd, prefix, names, feat):
    h = F.leaky_relu(F.linear(feat, sd["%s.%s.weight" % (prefix, names[0])], sd["%s.%s.bias" % (prefix, names[0])]), 0.1)
    return F.linear(h, sd["%s.%s.weight" % (prefix, names[1])], sd["%s.%s.bias" % (prefix, names[1])])


def d_forward(sd, x, size, sg_linear=False):
    """BaseDiscriminator.forward (models/gan/base.py:107-150) -> (d [B,1], projection, projection2)."""
    feat = d_penultimate(sd, x, size)
    feat_d = feat.detach() if sg_linear else feat
    d = _mlp(sd, "linear", ("l1", "l2"), feat_d)
    p1 = _mlp(sd, "projection", ("0", "2"), feat)
    p2 = _mlp(sd, "projection2", ("0", "2"), feat)
    return d + (p1.mean() + p2.mean()) * 0.0, p1, p2


def r1_penalty(sd, x, size):
    """G_D.forward(return_r1_loss=True) (train_stylegan2_contraD.py:129-136) / r1_loss (train_stylegan2.py:106-113):
    per-sample squared norm of d D(x) / d x, differentiable w.r.t. the parameters."""
    x = x.detach().requires_grad_(True)
    d, _, _ = d_forward(sd, x, size)
    (g,) = torch.autograd.grad(d.sum(), x, create_graph=True, retain_graph=True)
    return g.pow(2).reshape(g.shape[0], -1).sum(1)


# ------------------------------------------------------------------------------------------------ generator
def modulated_conv(sd, prefix, x, style, ksize, demodulate=True, upsample=False):
    """ModulatedConv2d.forward (generator.py:52-82), per-sample weights + grouped convolution like the reference."""
    weight = sd[prefix + ".weight"]                                                # [1, Cout, Cin, k, k]
    b, cin, h, w = x.shape
    cout = weight.shape[1]
    scale = 1 / math.sqrt(cin * ksize ** 2)
    s = equal_linear(style, sd[prefix + ".modulation.weight"], sd[prefix + ".modulation.bias"], bias_init=1.0)
    wt = scale * weight * s.view(b, 1, cin, 1, 1)
    if demodulate:
        wt = wt * torch.rsqrt(wt.pow(2).sum([2, 3, 4]) + 1e-8).view(b, cout, 1, 1, 1)
    x = x.reshape(1, b * cin, h, w)
    if upsample:
        wt = wt.transpose(1, 2).reshape(b * cin, cout, ksize, ksize)
        out = F.conv_transpose2d(x, wt, padding=0, stride=2, groups=b)
        out = out.view(b, cout, out.shape[2], out.shape[3])
        p = (4 - 2) - (ksize - 1)
        out = upfirdn2d(out, sd[prefix + ".blur.kernel"], pad=((p + 1) // 2 + 1, p // 2 + 1))
    else:
        out = F.conv2d(x, wt.view(b * cout, cin, ksize, ksize), padding=ksize // 2, groups=b)
        out = out.view(b, cout, out.shape[2], out.shape[3])
    return out


def style_layer(sd, prefix, x, style, noise, upsample=False):
    """StyleLayer.forward (generator.py:119-124): modulated conv, NoiseInjection (:91-94), FusedLeakyReLU."""
    out = modulated_conv(sd, prefix + ".conv", x, style, 3, upsample=upsample)
    out = out + sd[prefix + ".noise.weight"] * noise
    return fused_lrelu(out, sd[prefix + ".activate.bias"])


def to_rgb(sd, prefix, x, style, skip=None):
    """ToRGB.forward (generator.py:137-149)."""
    out = modulated_conv(sd, prefix + ".conv", x, style, 1, demodulate=False) + sd[prefix + ".bias"]
    if skip is not None:
        out = out + upfirdn2d(skip, sd[prefix + ".upsample.kernel"], up=2, pad=(2, 1))
    return out


def g_mapping(sd, z, n_mlp=8, lr_mlp=0.01):
    """Generator.style (generator.py:154-160): PixelNorm (layers.py:15-20) + 8 EqualLinear(fused lrelu)."""
    h = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(1, n_mlp + 1):
        h = equal_linear(h, sd["style.%d.weight" % i], sd["style.%d.bias" % i], lr_mul=lr_mlp, activation=True)
    return h


def n_latent_for(size):
    return int(math.log(size, 2)) * 2 - 2


def g_forward(sd, z, size, noises, z_mix=None, mix_layer=None):
    """Generator.forward (generator.py:233-290) in train mode with the random draws made explicit:
    noises: list of [B,1,H,W] per StyleLayer; z_mix / mix_layer: the style-mixing latent and the per-sample first
    layer index that uses it (n_latent = no mixing), generator.py:252-266.  Returns the image in [0,1] (unclamped)."""
    n_latent = n_latent_for(size)
    latent = g_mapping(sd, z)
    latents = latent.unsqueeze(1).repeat(1, n_latent, 1)
    if z_mix is not None:
        latent_mix = g_mapping(sd, z_mix).unsqueeze(1)
        mask = (torch.arange(n_latent)[None] < mix_layer.cpu().unsqueeze(1)).float().unsqueeze(-1).to(latents.device)
        latents = latents * mask + latent_mix * (1 - mask)
    b = z.shape[0]
    out = sd["input.const"].repeat(b, 1, 1, 1)
    out = style_layer(sd, "conv1", out, latents[:, 0], noises[0])
    skip = to_rgb(sd, "to_rgb1", out, latents[:, 1])
    idx = 1
    for j in range(int(math.log(size, 2)) - 2):
        out = style_layer(sd, "layers.%d" % (2 * j), out, latents[:, idx], noises[1 + 2 * j], upsample=True)
        out = style_layer(sd, "layers.%d" % (2 * j + 1), out, latents[:, idx + 1], noises[2 + 2 * j])
        skip = to_rgb(sd, "to_rgbs.%d" % j, out, latents[:, idx + 2], skip)
        idx += 2
    return 0.5 * skip + 0.5


def noise_shapes(size, batch):
    shapes = [(batch, 1, 4, 4)]
    for i in range(3, int(math.log(size, 2)) + 1):
        shapes += [(batch, 1, 2 ** i, 2 ** i)] * 2
    return shapes


# ------------------------------------------------------------------------------------------------ states
def make_d_state(size=32, small32=True, channel_multiplier=2, d_hidden=512, d_project=128, generator=None):
    """Random ResidualDiscriminatorP state_dict with the reference's keys, shapes and initial distributions
    (EqualConv2d: N(0,1) layers.py:101-103; FusedLeakyReLU bias 0; nn.Linear default init for the heads)."""
    ch = channels_for(size, channel_multiplier, small32)
    rn = lambda *s: torch.randn(*s, generator=generator)
    sd = {"layers.0.0.weight": rn(ch[size], 3, 1, 1), "layers.0.1.bias": torch.zeros(ch[size])}
    cin = ch[size]
    log_size = int(math.log(size, 2))
    for bi, i in enumerate(range(log_size, 2, -1), start=1):
        cout = ch[2 ** (i - 1)]
        p = "layers.%d" % bi
        sd[p + ".conv1.0.weight"] = rn(cin, cin, 3, 3)
        sd[p + ".conv1.1.bias"] = torch.zeros(cin)
        sd[p + ".conv2.0.kernel"] = make_kernel()
        sd[p + ".conv2.1.weight"] = rn(cout, cin, 3, 3)
        sd[p + ".conv2.2.bias"] = torch.zeros(cout)
        sd[p + ".skip.0.kernel"] = make_kernel()
        sd[p + ".skip.1.weight"] = rn(cout, cin, 1, 1)
        cin = cout
    sd["last_conv.0.weight"] = rn(ch[4], cin + 1, 3, 3)
    sd["last_conv.1.bias"] = torch.zeros(ch[4])
    nfeat = ch[4] * 16

    def lin(name, fin, fout):
        bound = 1 / math.sqrt(fin)
        sd[name + ".weight"] = (torch.rand(fout, fin, generator=generator) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(fout, generator=generator) * 2 - 1) * bound

    lin("linear.l1", nfeat, d_hidden); lin("linear.l2", d_hidden, 1)
    lin("projection.0", nfeat, d_hidden); lin("projection.2", d_hidden, d_project)
    lin("projection2.0", nfeat, d_hidden); lin("projection2.2", d_hidden, d_project)
    return sd


def make_g_state(size=32, small32=True, channel_multiplier=2, style_dim=512, n_mlp=8, lr_mlp=0.01, generator=None):
    """Random Generator state_dict (generator.py:152-227)."""
    ch = channels_for(size, channel_multiplier, small32)
    rn = lambda *s: torch.randn(*s, generator=generator)
    sd = {}
    for i in range(1, n_mlp + 1):
        sd["style.%d.weight" % i] = rn(style_dim, style_dim) / lr_mlp
        sd["style.%d.bias" % i] = torch.zeros(style_dim)
    sd["input.const"] = rn(1, ch[4], 4, 4)

    def modconv(p, cin, cout, k, upsample=False):
        sd[p + ".weight"] = rn(1, cout, cin, k, k)
        sd[p + ".modulation.weight"] = rn(cin, style_dim)
        sd[p + ".modulation.bias"] = torch.zeros(cin)
        if upsample:
            sd[p + ".blur.kernel"] = make_kernel() * 4

    def style(p, cin, cout, upsample=False):
        modconv(p + ".conv", cin, cout, 3, upsample)
        sd[p + ".noise.weight"] = torch.zeros(1)
        sd[p + ".activate.bias"] = torch.zeros(cout)

    def rgb(p, cin, upsample=True):
        if upsample:
            sd[p + ".upsample.kernel"] = make_kernel() * 4
        modconv(p + ".conv", cin, 3, 1)
        sd[p + ".bias"] = torch.zeros(1, 3, 1, 1)

    style("conv1", ch[4], ch[4])
    rgb("to_rgb1", ch[4], upsample=False)
    cin = ch[4]
    for j, i in enumerate(range(3, int(math.log(size, 2)) + 1)):
        cout = ch[2 ** i]
        style("layers.%d" % (2 * j), cin, cout, upsample=True)
        style("layers.%d" % (2 * j + 1), cout, cout)
        rgb("to_rgbs.%d" % j, cout)
        cin = cout
    return sd


def trainable(sd):
    return {k: v for k, v in sd.items() if not k.endswith(".kernel")}


# ------------------------------------------------------------------------------------------------ losses / one step
def nt_xent(out1, out2, temperature=0.1):
    """training/criterion.py:24-45 (single process)."""
    n = out1.shape[0]
    z = torch.cat([out1, out2], 0)
    sim = z @ z.t() / temperature
    sim.fill_diagonal_(-5e4)
    lsm = F.log_softmax(sim, dim=1)
    return -(lsm[:n, n:].diag() + lsm[n:, :n].diag()).sum() / (2 * n)


def supcon_fake(out1, out2, others, temperature=0.1):
    """training/gan/contrad.py:8-32."""
    n = out1.shape[0]
    z = torch.cat([out1, out2, others], 0)
    sim = z @ z.t() / temperature
    sim.fill_diagonal_(-5e4)
    mask = torch.zeros_like(sim)
    mask[2 * n:, 2 * n:] = 1
    mask.fill_diagonal_(0)
    sim = sim[2 * n:]
    mask = mask[2 * n:]
    mask = mask / mask.sum(1, keepdim=True)
    lsm = F.log_softmax(sim, dim=1)
    return -(mask * lsm).sum(1).mean()


def gd_losses(sd_d, size, real_aug2, fake_aug, temp=0.1, lbd_a=1.0):
    """D-step of train_stylegan2_contraD.py:95-164 on already-augmented inputs: D on the fakes [n] and on
    cat(real, real) [2n] separately (two minibatch-stddev groupings), the four normalisations, `_loss_D_fn`.
    Returns (simclr + lbd_a * supcon, penalty (nonsat L_dis), d_real mean, d_gen mean)."""
    n = fake_aug.shape[0]
    d_gen, others, fakes = d_forward(sd_d, fake_aug, size, sg_linear=True)
    d_rs, views_r, reals = d_forward(sd_d, real_aug2, size, sg_linear=True)
    views_r, reals, others, fakes = (F.normalize(t) for t in (views_r, reals, others, fakes))
    simclr = nt_xent(views_r[:n], views_r[n:], temp)
    sup = supcon_fake(reals[:n], reals[n:], fakes, temp)
    d_real = d_rs[:n]
    penalty = F.softplus(d_gen).mean() + F.softplus(-d_real).mean()
    return simclr + lbd_a * sup, penalty, d_real.mean(), d_gen.mean()


def g_loss(sd_d, size, fake_aug):
    """G-step: `_loss_G_fn(D(aug(G(z))))` (train_stylegan2_contraD.py:117-119,141-146)."""
    d_gen, _, _ = d_forward(sd_d, fake_aug, size, sg_linear=False)
    return F.softplus(-d_gen).mean()
