"""Import stub for ``imageio`` (absent; reference train_gan.py:6-9 imports it and pokes
``imageio.core.util._precision_warn``).  Saving falls back to PIL when available."""
from . import core  # noqa: F401
import numpy as _np


def _to_uint8(arr):
    arr = _np.asarray(arr)
    if arr.dtype != _np.uint8:
        arr = (_np.clip(arr, 0.0, 1.0) * 255.0 + 0.5).astype(_np.uint8)
    return arr


def imsave(path, image, **kwargs):
    from PIL import Image
    Image.fromarray(_to_uint8(image)).save(path)


imwrite = imsave


def mimsave(path, images, **kwargs):
    from PIL import Image
    frames = [Image.fromarray(_to_uint8(im)) for im in images]
    if frames:
        frames[0].save(path, save_all=True, append_images=frames[1:], loop=0)
