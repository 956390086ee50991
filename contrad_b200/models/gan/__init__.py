"""Mirror of models/gan/__init__.py:2-31 (``get_architecture``)."""


def get_architecture(architecture, image_size, P=None):
    if architecture == "sndcgan":
        from .sndcgan import D_SNDCGAN, G_SNDCGAN
        generator = G_SNDCGAN(image_size=image_size)
        discriminator = D_SNDCGAN(image_size=image_size, mlp_linear=True, d_hidden=512)
        return generator, discriminator
    if architecture in ("snresnet18", "stylegan2", "stylegan2_512"):
        raise NotImplementedError(
            "architecture %r is a later row of the hot-path scope table (SURVEY 8a a18-a22 / 8f f4); "
            "round 1 of contrad_b200 builds 'sndcgan'" % architecture)
    raise NotImplementedError()
