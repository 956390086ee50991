"""Time the UNMODIFIED reference (oracle/_ref, made by oracle/make_ref.py) on the benchmark workload.

TEST / MEASUREMENT INFRASTRUCTURE: imported only by bench.py's reference legs (`--impl reference`, `cpu_baseline`,
`eager_gpu_baseline`) and by tests.  Nothing here is on the product path and nothing here calls contrad_b200 kernels:
the modules that run are the reference's own files (`models/gan/sndcgan.py`, `models/gan/base.py`, `augment/`,
`training/gan/contrad.py`, `training/criterion.py`, `third_party/gather_layer.py`), the optimisers are
`torch.optim.Adam`, and the only repo code involved are the import shims for gin / tensorboardX / imageio / kornia
(`contrad_b200/compat`), which carry no arithmetic of the path.

Two drivers:

* ``run_gpu`` - the reference's own training loop: ``train_gan.train`` (train_gan.py:123-227) is CALLED, with the
  reference's own set-up sequence of ``worker`` (train_gan.py:230-318: gin files, `get_architecture`, SyncBatchNorm
  conversion, `.cuda()`, Adam, `get_augment(...).cuda()`, DistributedDataParallel with broadcast_buffers=False) around
  it.  The dataset is replaced by a synthetic loader (no network for CIFAR-10) yielding pinned fp32 `[B,3,32,32]`
  batches, which also takes the step time stamps: every iteration of the loop ends with `.item()` reads and
  `dist.barrier()`, so the host clock at `next(loader)` is a device-synchronised step boundary.
* ``run_cpu`` - the same modules on the host cores.  `train()` hard-codes `.cuda()` / NCCL, so the loop body
  (train_gan.py:141-179) is restated here line by line, without DDP.
"""
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

from . import ref_import

GIN_FILES = ("configs/defaults/gan.gin", "configs/defaults/augment.gin")


def _activate(gin_config):
    if os.path.isdir(os.path.join(ref_import._VENDORED, "augment")):      # prefer the copy that also exists on the GPU box
        ref_import.REFERENCE_ROOT = ref_import._VENDORED
    gin = ref_import.activate(gin_files=GIN_FILES + (gin_config,))
    root = ref_import.REFERENCE_ROOT
    return gin, root


def _import_train_gan(root):
    """`import train_gan` with the reference root as the working directory (utils.Logger et al. use relative paths)."""
    import importlib
    if "train_gan" in sys.modules and not (getattr(sys.modules["train_gan"], "__file__", "") or "").startswith(root):
        del sys.modules["train_gan"]
    return importlib.import_module("train_gan")


class _SyntheticLoader(object):
    """Stands in for `cycle(DataLoader(CIFAR10, pin_memory=True, ...))`: an endless iterator of (images, labels) with
    images fp32 U[0,1) `[B,3,32,32]` in pinned host memory.  Records the host time of every `next()`."""

    def __init__(self, batch, size=32, pool=4, pin=True, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.pool = [torch.rand(batch, 3, size, size, generator=g) for _ in range(pool)]
        if pin:
            self.pool = [t.pin_memory() for t in self.pool]
        self.labels = torch.zeros(batch, dtype=torch.long)
        self.stamps = []
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        self.stamps.append(time.perf_counter())
        self.i += 1
        return self.pool[self.i % len(self.pool)], self.labels


def run_gpu(steps, warmup, global_batch=512, architecture="sndcgan", gin_config="configs/gan/cifar10/c10_b512.gin",
            local_rank=0, port=29731):
    """Returns {"ms_per_step", "images_per_s", ...} measured on this rank (max over ranks is taken by the caller)."""
    import torch.distributed as dist
    import torch.nn as nn
    import torch.optim as optim
    from torch.nn.parallel import DistributedDataParallel

    gin, root = _activate(gin_config)
    cwd = os.getcwd()
    own_pg = False
    try:
        os.chdir(root)
        tg = _import_train_gan(root)
        from augment import get_augment
        from models.gan import get_architecture
        from training.gan import setup

        torch.cuda.set_device(local_rank)
        if not dist.is_initialized():                      # train_gan.py:239-242 (world 1 = the single-GPU run)
            dist.init_process_group(backend="nccl", init_method="tcp://127.0.0.1:%d" % port, world_size=1, rank=0)
            own_pg = True
        world, rank = dist.get_world_size(), dist.get_rank()
        P = SimpleNamespace(mode="contrad", aug="simclr", penalty="none", temp=0.1, lbd_a=1.0, use_warmup=True,
                            architecture=architecture, distributed=True, rank=rank, n_gpus_per_node=world,
                            no_fid=True, no_gif=True, n_eval_avg=1, print_every=10 ** 9, evaluate_every=10 ** 9,
                            save_every=10 ** 9, starting_step=1, eval_seed=0)
        P = setup(P)
        options = tg.get_options_dict()
        options["batch_size"] = options["batch_size"] // world                     # train_gan.py:247
        assert options["batch_size"] * world == global_batch, (options["batch_size"], world, global_batch)
        image_size = (32, 32, 3)
        torch.manual_seed(1234 + rank); np.random.seed(1234 + rank)
        generator, discriminator = get_architecture(architecture, image_size, P=P)
        generator = nn.SyncBatchNorm.convert_sync_batchnorm(generator)              # train_gan.py:268-271
        discriminator = nn.SyncBatchNorm.convert_sync_batchnorm(discriminator)
        generator, discriminator = generator.cuda(), discriminator.cuda()
        G_opt = optim.Adam(generator.parameters(), lr=options["lr"], betas=options["beta"])
        D_opt = optim.Adam(discriminator.parameters(), lr=options["lr_d"], betas=options["beta"])
        P.augment_fn = get_augment(mode=P.aug).cuda()                               # train_gan.py:310-313
        generator = DistributedDataParallel(generator, device_ids=[local_rank], broadcast_buffers=False)
        generator.sample_latent = generator.module.sample_latent
        discriminator = DistributedDataParallel(discriminator, device_ids=[local_rank], broadcast_buffers=False)

        class _Quiet(object):
            def log(self, s): pass
            def log_dirname(self, s): pass
            def scalar_summary(self, *a): pass

        loader = _SyntheticLoader(options["batch_size"], seed=rank)
        options["max_steps"] = warmup + steps                                       # the loop runs steps 1 .. max_steps
        tg.train(P, options, P.train_fn, models=(generator, discriminator), optimizers=(G_opt, D_opt),
                 train_loader=loader, logger=_Quiet())
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        dt = t_end - loader.stamps[warmup]
        assert len(loader.stamps) == warmup + steps
        del generator, discriminator, G_opt, D_opt
        torch.cuda.empty_cache()
        return {"ms_per_step": 1e3 * dt / steps, "images_per_s": global_batch * steps / dt, "steps": steps,
                "warmup": warmup, "world": world, "per_gpu_batch": options["batch_size"],
                "flags": {"cudnn.allow_tf32": bool(torch.backends.cudnn.allow_tf32),
                          "matmul.allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32),
                          "cudnn.benchmark": bool(torch.backends.cudnn.benchmark)},
                "what": "oracle/_ref train_gan.train() (train_gan.py:123-227) on unmodified reference modules, "
                        "DDP + SyncBatchNorm + torch.optim.Adam, PyTorch default flags, pinned-host batches, "
                        "5 .item() reads + dist.barrier() per step"}
    finally:
        os.chdir(cwd)
        if own_pg:
            dist.destroy_process_group()
        ref_import.deactivate()


def run_gpu_stylegan2(steps, warmup, architecture="stylegan2", size=32, batch=64,
                      gin_config="configs/gan/stylegan2/c10_style64.gin", lbd_r1=0.1, no_lazy=True, halflife_k=1000,
                      device_ids=None):
    """BASELINE configs 4 / 5 through the reference's own `train_stylegan2_contraD.train()` (:166-296) on unmodified
    reference modules with its own CUDA ops (pre-built by oracle/make_ref.py --ext), the set-up of `worker()`
    (:298-378: G_D, g_ema, torch.optim.Adam, `nn.DataParallel(GD)`), synthetic pinned-host batches.  One process;
    `device_ids` = the GPUs DataParallel replicates over (None = all visible, as the script does)."""
    import importlib
    import torch.nn as nn
    import torch.optim as optim
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    ext = os.path.join(ref_import._VENDORED, "_torch_ext")
    if os.path.isdir(ext):
        os.environ["TORCH_EXTENSIONS_DIR"] = ext
    gin, root = _activate(gin_config)
    cwd = os.getcwd()
    try:
        os.chdir(root)
        if "train_stylegan2_contraD" in sys.modules:
            del sys.modules["train_stylegan2_contraD"]
        ts = importlib.import_module("train_stylegan2_contraD")
        from augment import get_augment
        from models.gan import get_architecture
        from training.gan import setup
        torch.cuda.set_device(0)
        P = SimpleNamespace(mode="contrad", aug="simclr", penalty="none", temp=0.1, lbd_a=1.0, use_warmup=True,
                            architecture=architecture, distributed=False, no_lazy=no_lazy, d_reg_every=1 if no_lazy else 16,
                            lbd_r1=lbd_r1, style_mix=0.9, halflife_k=halflife_k, ema_start_k=halflife_k, halflife_lr=0,
                            no_fid=True, no_gif=True, n_eval_avg=1, print_every=10 ** 9, evaluate_every=10 ** 9,
                            save_every=10 ** 9, starting_step=1, eval_seed=0, rank=0)
        P = setup(P)
        options = ts.get_options_dict()
        assert options["batch_size"] == batch, (options["batch_size"], batch)
        P.accum = 0.5 ** (options["batch_size"] / (P.halflife_k * 1000))
        torch.manual_seed(1234); np.random.seed(1234)
        image_size = (size, size, 3)
        generator, discriminator = get_architecture(architecture, image_size, P=P)
        g_ema, _ = get_architecture(architecture, image_size, P=P)
        generator, discriminator = generator.cuda(), discriminator.cuda()
        P.augment_fn = get_augment(mode=P.aug).cuda()
        GD = ts.G_D(generator, discriminator, P.augment_fn).cuda()
        g_ema = g_ema.cuda(); g_ema.eval()
        G_opt = optim.Adam(generator.parameters(), lr=options["lr"], betas=options["beta"])
        D_opt = optim.Adam(discriminator.parameters(), lr=options["lr_d"], betas=options["beta"])
        GD = nn.DataParallel(GD, device_ids=device_ids)

        class _Quiet(object):
            logdir = None
            def log(self, s): pass
            def log_dirname(self, s): pass
            def scalar_summary(self, *a): pass

        loader = _SyntheticLoader(batch, size=size, pool=2)
        options["max_steps"] = warmup + steps
        ts.train(P, options, models=(generator, discriminator, GD, g_ema), optimizers=(G_opt, D_opt),
                 train_loader=loader, logger=_Quiet())
        for d in range(torch.cuda.device_count() if device_ids is None else len(device_ids)):
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - loader.stamps[warmup]
        return {"ms_per_step": 1e3 * dt / steps, "images_per_s": batch * steps / dt, "steps": steps, "warmup": warmup,
                "n_gpus": torch.cuda.device_count() if device_ids is None else len(device_ids),
                "what": "oracle/_ref train_stylegan2_contraD.train() (:166-296) on unmodified reference modules and CUDA ops, "
                        "nn.DataParallel(G_D), torch.optim.Adam, PyTorch default flags, R1 every %d step(s)" % P.d_reg_every}
    finally:
        os.chdir(cwd)
        ref_import.deactivate()


def run_cpu(steps, warmup, batch=512, threads=None, architecture="sndcgan",
            gin_config="configs/gan/cifar10/c10_b512.gin", seconds_budget=None):
    """The reference modules on the host CPU (all `threads` torch threads): loop body of train_gan.py:141-179.
    With `seconds_budget`, stops after at least 2 timed steps once the budget is used up."""
    import torch.optim as optim
    gin, root = _activate(gin_config)
    cwd = os.getcwd()
    prev_threads = torch.get_num_threads()
    st_np, st_t = np.random.get_state(), torch.get_rng_state()
    try:
        os.chdir(root)
        tg = _import_train_gan(root)
        from augment import get_augment
        from models.gan import get_architecture
        from training.gan import setup
        from utils import set_grad
        if threads:
            torch.set_num_threads(threads)
        P = SimpleNamespace(mode="contrad", aug="simclr", penalty="none", temp=0.1, lbd_a=1.0, use_warmup=True,
                            architecture=architecture, distributed=False, rank=0)
        P = setup(P)
        opt = tg.get_options_dict()
        opt["batch_size"] = batch
        torch.manual_seed(1234); np.random.seed(1234)
        generator, discriminator = get_architecture(architecture, (32, 32, 3), P=P)
        opt_G = optim.Adam(generator.parameters(), lr=opt["lr"], betas=opt["beta"])
        opt_D = optim.Adam(discriminator.parameters(), lr=opt["lr_d"], betas=opt["beta"])
        P.augment_fn = get_augment(mode=P.aug)
        train_fn = P.train_fn
        loader = _SyntheticLoader(batch, pin=False)
        reads = []

        def one(step):                                   # train_gan.py:141-179, n_critic = 1
            generator.train(); discriminator.train()
            tg._update_warmup(opt_G, step, opt["warmup"], opt["lr"])
            tg._update_warmup(opt_D, step, opt["warmup"], opt["lr_d"])
            set_grad(generator, False); set_grad(discriminator, True)
            images, _ = next(loader)
            gen_images = tg._sample_generator(generator, images.size(0), enable_grad=False)
            d_loss, aux = train_fn["D"](P, discriminator, opt, images, gen_images)
            loss = d_loss + aux["penalty"]
            opt_D.zero_grad(); loss.backward(); opt_D.step()
            r = [d_loss.item(), aux["penalty"].item(), aux["d_real"].item(), aux["d_gen"].item()]
            set_grad(generator, True); set_grad(discriminator, False)
            gen_images = tg._sample_generator(generator, images.size(0))
            g_loss = train_fn["G"](P, discriminator, opt, images, gen_images)
            opt_G.zero_grad(); g_loss.backward(); opt_G.step()
            r.append(g_loss.item())
            generator.eval(); discriminator.eval()
            reads.append(r)

        for w in range(warmup):
            one(w + 1)
        t0 = time.perf_counter()
        done = 0
        for s in range(steps):
            one(warmup + s + 1)
            done += 1
            if seconds_budget is not None and done >= 2 and time.perf_counter() - t0 > seconds_budget:
                break
        dt = time.perf_counter() - t0
        assert all(np.isfinite(v) for r in reads for v in r), reads[-1]
        return {"ms_per_step": 1e3 * dt / done, "images_per_s": batch * done / dt, "steps": done, "warmup": warmup,
                "threads": torch.get_num_threads(), "batch": batch, "last_losses": reads[-1],
                "what": "oracle/_ref modules (get_architecture('%s'), training.gan.contrad, get_augment('simclr'), "
                        "torch.optim.Adam) on the host CPU, loop body of train_gan.py:141-179, fp32" % architecture}
    finally:
        os.chdir(cwd)
        torch.set_num_threads(prev_threads)
        np.random.set_state(st_np); torch.set_rng_state(st_t)
        ref_import.deactivate()
