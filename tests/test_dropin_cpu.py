"""The drop-in path on CPU: the UNMODIFIED reference script `train_gan.py` is executed (module body, not `__main__`) after
`contrad_b200.dropin.install()`, and its own `parse_args -> setup -> gin files -> get_options_dict -> get_architecture ->
get_augment` sequence (train_gan.py:230-318) must resolve to the contrad_b200 mirrors with the reference's parameter
counts and state_dict layout.  No kernel runs (no GPU here); the GPU counterpart is tests/test_gpu_dropin.py.

Needs the reference sources: /root/reference in the build container or the copy under oracle/_ref (oracle/make_ref.py);
skipped when neither exists."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in (os.path.join(REPO, "oracle", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(cand, "train_gan.py")):
            return cand
    return None


_CODE = r'''
import json, os, runpy, sys
repo, root = sys.argv[1], sys.argv[2]
sys.path.insert(0, repo)
from contrad_b200 import dropin
installed = dropin.install()
sys.path.insert(0, root)
os.chdir(root)
sys.argv = ["train_gan.py", "configs/gan/cifar10/c10_b512.gin", "sndcgan", "--mode=contrad", "--aug=simclr", "--use_warmup"]
ns = runpy.run_path("train_gan.py", run_name="train_gan_under_test")       # imports + definitions, not the __main__ block
import gin
from pathlib import Path
P = ns["parse_args"]()
P.gin_stem = Path(P.gin_config).stem
P = ns["setup"](P)
gin.parse_config_files_and_bindings(["configs/defaults/gan.gin", "configs/defaults/augment.gin", P.gin_config], [])
options = ns["get_options_dict"]()
G, D = ns["get_architecture"](P.architecture, (32, 32, 3), P=P)
aug = ns["get_augment"](mode=P.aug)
out = {
    "installed": len(installed),
    "filename": P.filename,
    "train_fn_module": P.train_fn["D"].__module__,
    "options": {k: options[k] for k in ("batch_size", "loss", "warmup", "lr", "n_critic")},
    "G_module": type(G).__module__, "D_module": type(D).__module__, "aug_module": type(aug).__module__,
    "aug_layers": [type(m).__name__ for m in aug],
    "n_params_G": ns["count_parameters"](G), "n_params_D": ns["count_parameters"](D),
    "D_keys": sorted(D.state_dict().keys()), "G_keys": sorted(G.state_dict().keys()),
    "rrc_scale": list(aug[0].scale), "cj_hue": list(aug[2].fn.hue),
    "utils_is_reference": os.path.abspath(sys.modules["utils"].__file__).startswith(os.path.abspath(root)),
}
print("RESULT " + json.dumps(out))
'''


@pytest.mark.timeout(300)
def test_unmodified_train_gan_resolves_to_the_mirrors():
    root = _reference_root()
    if root is None:
        pytest.skip("reference sources not available (neither oracle/_ref nor /root/reference)")
    r = subprocess.run([sys.executable, "-c", _CODE, REPO, root], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=280)
    text = r.stdout.decode()
    assert r.returncode == 0, text[-3000:]
    res = json.loads([ln for ln in text.splitlines() if ln.startswith("RESULT ")][-1][7:])
    assert res["filename"] == "contrad_simclr_L1.0_T0.1"                   # training/gan/__init__.py:19-20
    assert res["train_fn_module"] == "contrad_b200.training.gan.contrad"
    assert res["options"] == {"batch_size": 512, "loss": "nonsat", "warmup": 3000, "lr": 2e-4, "n_critic": 1}
    assert res["G_module"].startswith("contrad_b200.") and res["D_module"].startswith("contrad_b200.")
    assert res["aug_module"].startswith("contrad_b200.")
    assert res["aug_layers"] == ["RandomResizeCropLayer", "HorizontalFlipLayer", "RandomApply", "RandomApply"]
    assert res["n_params_D"] == 18568961 and res["n_params_G"] == 3828739    # the reference's own counts
    assert res["rrc_scale"] == [0.2, 1.0] and res["cj_hue"] == [-0.1, 0.1]  # gin bindings reached the mirrors
    assert res["utils_is_reference"]                                         # host glue stays the reference's own
    # state_dict layout of the reference: spectral-norm triplets and BatchNorm buffers
    from oracle import contrad_oracle as O
    assert res["D_keys"] == sorted(O.make_d_state().keys())
    assert res["G_keys"] == sorted(O.make_g_state().keys())
