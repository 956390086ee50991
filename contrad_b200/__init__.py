"""contrad_b200: B200-native (sm_100a) implementation of the ContraD per-step training hot path.

Host side mirrors the reference's Python surfaces (augment / training.criterion /
training.gan.contrad / third_party.gather_layer / models.gan / penalty); every operator is a
``torch.autograd.Function`` over the C ABI in ``include/contrad_b200.h`` (hand-written CUDA,
tcgen05 + TMA for the dense contractions).  ``contrad_b200.dropin.install()`` registers the
mirrors under the reference's module names so ``train_gan.py`` runs unchanged.
"""
__version__ = "0.1.0"
