"""Mirror of models/gan/__init__.py:2-31 (``get_architecture``)."""


def get_architecture(architecture, image_size, P=None):
    if architecture == "sndcgan":
        from .sndcgan import D_SNDCGAN, G_SNDCGAN
        generator = G_SNDCGAN(image_size=image_size)
        discriminator = D_SNDCGAN(image_size=image_size, mlp_linear=True, d_hidden=512)
        return generator, discriminator
    if architecture == "stylegan2":
        from .stylegan2.discriminator import ResidualDiscriminatorP
        from .stylegan2.generator import Generator
        resolution = image_size[0]
        generator = Generator(size=resolution, n_mlp=8, small32=True)
        discriminator = ResidualDiscriminatorP(size=resolution, small32=True, mlp_linear=True, d_hidden=512)
        return generator, discriminator
    if architecture == "stylegan2_512":
        from .stylegan2.discriminator import ResidualDiscriminatorP
        from .stylegan2.generator import Generator
        resolution = image_size[0]
        generator = Generator(size=resolution, n_mlp=8, channel_multiplier=1.0)
        discriminator = ResidualDiscriminatorP(size=resolution, channel_multiplier=1.0, mlp_linear=True, d_hidden=512)
        return generator, discriminator
    if architecture == "snresnet18":
        from .sndcgan import G_SNDCGAN
        from .snresnet import D_SNResNet18
        generator = G_SNDCGAN(image_size=image_size)
        discriminator = D_SNResNet18(mlp_linear=True, d_hidden=1024)
        return generator, discriminator
    raise NotImplementedError()
