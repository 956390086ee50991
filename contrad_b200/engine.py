"""Host-side mirror of one iteration of the reference training loop (train_gan.py:141-179):
warm-up LR, `set_grad` toggling, D step (G forward without grad, loss_D_fn, backward, Adam), G step
(G forward, loss_G_fn through the frozen D, backward, Adam).  Used by bench.py, smoke() and the parity
tests; the reference's own `train_gan.py` drives the same plug-ins through contrad_b200.dropin.

``GraphedTrainStep`` replays the same step as ONE CUDA graph (about 240 kernel launches and their Python dispatch
collapse into a single cudaGraphLaunch), which is what keeps the GPU busy when the per-GPU batch gets small
(DDP b512 over 8 GPUs = 64 images per rank)."""
import torch

from . import staging


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


def _is_ddp(model):
    return isinstance(model, torch.nn.parallel.DistributedDataParallel)


def broadcast_parameters(model, src=0):
    """What DistributedDataParallel does at construction (train_gan.py:311-313): every rank starts from rank
    `src`'s parameters and buffers.  For models that are NOT wrapped in DDP (see `allreduce_gradients`)."""
    import torch.distributed as dist
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t, src)


def allreduce_gradients(model, params=None, group=None, flat=None):
    """DDP's gradient averaging for a bare module (capturable in a CUDA graph, no bucket hooks): used when
    `P.distributed` and the model is not DDP-wrapped.

    The gradients are packed into ONE contiguous buffer by a multi-tensor copy, all-reduced with ReduceOp.AVG in a
    single NCCL call (a contiguous message lets NCCL use its bandwidth protocols; a grouped launch over the 26 separate
    tensors ran in the latency protocol: 15 MB took 350 us on 8 GPUs, profiles/launches_r2_n8.md), and the parameters'
    `.grad` are re-pointed at views of that buffer - no copy back, no divide.  `flat` (optional) is a persistent buffer of
    the right size (engine.GradSync keeps one per bucket).  Backends without AVG (gloo in the CPU tests): SUM + divide."""
    import torch.distributed as dist
    params = [p for p in (list(model.parameters()) if params is None else params) if p.grad is not None]
    if not params:
        return
    grads = [p.grad for p in params]
    pad = lambda n: (n + 31) // 32 * 32                  # every view starts on a 128-byte boundary (the fused Adam kernel
    total = sum(pad(g.numel()) for g in grads)           # reads gradients as float4); the padding stays zero for ever
    if flat is None or flat.numel() != total or flat.device != grads[0].device:
        flat = torch.zeros(total, device=grads[0].device, dtype=grads[0].dtype)
    views, off = [], 0
    for g in grads:
        views.append(flat[off:off + g.numel()].view_as(g))
        off += pad(g.numel())
    torch._foreach_copy_(views, grads)
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, group=group)
        flat.div_(dist.get_world_size(group))
    for p, v in zip(params, views):
        p.grad = v
    return flat


class GradSync(object):
    """Overlapped gradient averaging for a bare (non-DDP) module under `P.distributed` (train_gan.py:311-313 semantics).

    * `early` parameters (for D_SNDCGAN: the six head layers, 50 of the 74 MB, whose gradients are complete right after
      the heads' backward) are all-reduced on a side stream as soon as the last of them has accumulated its gradient -
      the transfer then runs under the backbone's backward;
    * `finish()` reduces the remaining gradients on the same side stream and makes the current stream wait for it;
    * the collectives use their OWN process group (own NCCL communicator), so they do not queue behind - or delay - the
      latency-bound SyncBatchNorm / embedding collectives of the default group;
    * everything is stream-ordered (fork = side.wait_stream(current), join = current.wait_stream(side)), so the whole
      thing is capturable inside the CUDA graph of the step, where it becomes a parallel branch.
    Without NCCL (gloo CPU tests) it degrades to one flat all-reduce in `finish()`."""

    def __init__(self, model, early=()):
        import torch.distributed as dist
        self.model = model
        self.nccl = dist.get_backend() == "nccl"
        self.early = [p for p in early]
        early_ids = {id(p) for p in self.early}
        self.rest = [p for p in model.parameters() if id(p) not in early_ids]
        self.armed, self.seen, self.fired = False, 0, False
        self.group, self.stream = None, None
        self.flat_early, self.flat_rest = None, None          # persistent contiguous all-reduce buffers
        if self.nccl:
            self.group = dist.new_group(backend="nccl")
            self.stream = torch.cuda.Stream()
            for p in self.early:
                p.register_post_accumulate_grad_hook(self._on_grad)

    def arm(self):
        """Call right before the backward pass whose gradients are to be averaged."""
        self.armed, self.seen, self.fired = True, 0, False

    def _on_grad(self, p):
        if not self.armed:
            return
        self.seen += 1
        if self.seen == len(self.early):
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                self.flat_early = allreduce_gradients(self.model, self.early, group=self.group, flat=self.flat_early)
            self.fired = True

    def finish(self):
        """After backward(): reduce what the hooks did not, then join the side stream."""
        self.armed = False
        if not self.nccl:
            allreduce_gradients(self.model)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            if not self.fired and self.early:
                self.flat_early = allreduce_gradients(self.model, self.early, group=self.group, flat=self.flat_early)
            self.flat_rest = allreduce_gradients(self.model, self.rest, group=self.group, flat=self.flat_rest)
        torch.cuda.current_stream().wait_stream(self.stream)


def _grad_sync(model):
    """The GradSync of a bare module (created on first use; every rank creates them in the same order)."""
    gs = getattr(model, "_cb200_grad_sync", None)
    if gs is None:
        early = list(model.early_gradient_parameters()) if hasattr(model, "early_gradient_parameters") else []
        gs = GradSync(model, early)
        object.__setattr__(model, "_cb200_grad_sync", gs)
    return gs


def train_step(P, opt, train_fn, models, optimizers, images, step, use_warmup=True, record_grad_norms=False):
    """One step (n_critic = 1).  Returns a dict of 0-dim tensors (no host sync inside)."""
    generator, discriminator = models
    sync_d = getattr(P, "distributed", False) and not _is_ddp(discriminator)
    sync_g = getattr(P, "distributed", False) and not _is_ddp(generator)
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
    if sync_d:
        _grad_sync(discriminator).arm()
    loss.backward()
    if sync_d:
        _grad_sync(discriminator).finish()
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
    if sync_g:
        _grad_sync(generator).arm()
    g_loss.backward()
    if sync_g:
        _grad_sync(generator).finish()
    if record_grad_norms:
        out["g_grad_norm"] = grad_norm(generator)
    opt_G.step()
    out["g_loss"] = g_loss.detach()
    return out


def stylegan2_schedule(P, opt, optimizers, step):
    """Host-side LR schedule of train_stylegan2_contraD.py:198-205."""
    from .training.gan import stylegan2 as T
    opt_G, opt_D = optimizers
    if P.use_warmup:
        update_warmup(opt_G, step, opt["warmup"], opt["lr"])
        update_warmup(opt_D, step, opt["warmup"], opt["lr_d"])
    if (not P.use_warmup) or step > opt["warmup"]:
        T.update_lr(opt_G, step, opt["batch_size"], P.halflife_lr, opt["lr"])
        T.update_lr(opt_D, step, opt["batch_size"], P.halflife_lr, opt["lr_d"])


def stylegan2_ema_decay(P, opt, step):
    """train_stylegan2_contraD.py:207-208."""
    return P.accum if (step * opt["batch_size"]) > (P.ema_start_k * 1000) else 0


def train_step_stylegan2(P, opt, GD, g_ema, optimizers, images, step, record_grad_norms=False, schedule=True):
    """One iteration of train_stylegan2_contraD.py:195-236 (n_critic = 1): LR schedule, EMA `accumulate`, G step
    through the frozen D, D step (contrastive losses + nonsat L_dis + lazy R1 every `P.d_reg_every` steps).
    GD is `training.gan.stylegan2.G_D(G, D, augment_fn)`; P carries use_warmup, halflife_lr, ema_start_k, accum,
    d_reg_every, lbd_r1, style_mix, temp, lbd_a, distributed.  Returns a dict of 0-dim tensors (no host sync)."""
    from .training.gan import stylegan2 as T
    generator, discriminator = GD.G, GD.D
    opt_G, opt_D = optimizers
    d_regularize = (step % P.d_reg_every == 0) and (P.lbd_r1 > 0)
    if schedule:
        stylegan2_schedule(P, opt, optimizers, step)
    if g_ema is not None:
        T.accumulate(g_ema, generator, stylegan2_ema_decay(P, opt, step))
    generator.train()
    discriminator.train()
    out = {}

    set_grad(generator, True)
    set_grad(discriminator, False)
    d_gen = GD(P, images, style_mix=P.style_mix, train_G=True)
    g_loss = T.loss_G_fn(d_gen)
    opt_G.zero_grad()
    g_loss.backward()
    if record_grad_norms:
        out["g_grad_norm"] = grad_norm(generator)
    opt_G.step()
    out["g_loss"] = g_loss.detach()

    set_grad(generator, False)
    set_grad(discriminator, True)
    d_all, view_r, view_f = GD(P, images, style_mix=P.style_mix)
    d_loss, aux = T.loss_D_fn(P, d_all, view_r, view_f)
    loss = d_loss + aux["penalty"]
    if d_regularize:
        r1 = GD(P, images, return_r1_loss=True).mean()
        loss = loss + (0.5 * P.lbd_r1) * r1 * P.d_reg_every
        out["d_r1"] = r1.detach()
    opt_D.zero_grad()
    loss.backward()
    if record_grad_norms:
        out["d_grad_norm"] = grad_norm(discriminator)
    opt_D.step()
    out.update(d_loss=d_loss.detach(), d_penalty=aux["penalty"].detach(), d_real=aux["d_real"].detach(),
               d_gen=aux["d_gen"].detach())
    return out


class GraphedTrainStep(object):
    """``train_step`` captured once into a CUDA graph and replayed.

    usage:  step_fn = GraphedTrainStep(P, opt, train_fn, (G, D), (opt_G, opt_D));  out = step_fn(images, step)

    * the first ``eager_steps`` calls run the ordinary eager step on a side stream (allocator warm-up, optimiser
      state, lazily created buffers), the next call captures, every later call replays;
    * per-step HOST inputs (crop boxes, jitter order, latents, Adam scalars) are not baked in: the capture runs under
      ``staging.Recorder``; each call re-draws them in the eager order and refreshes the static device buffers
      before the replay, so the host RNG streams advance exactly as in the eager loop;
    * device RNG draws (flip / jitter factors / apply masks) are torch CUDA-generator ops, which torch's graph
      support re-seeds per replay;
    * optimisers must be ``contrad_b200.optim.FusedAdam`` (device-resident step scalars); the models must not be
      DDP-wrapped - with ``P.distributed`` the gradients are averaged by one captured NCCL all-reduce;
    * the returned dict holds STATIC tensors that the next call overwrites;
    * the host RNG draws of step s+1 are made right after the replay of step s is launched, so that the host
      sampling (~2 ms) overlaps the GPU (same draw order as the eager loop; one set of draws is left unused when
      the loop ends).
    """

    def __init__(self, P, opt, train_fn, models, optimizers, use_warmup=True, eager_steps=3):
        from .optim import FusedAdam
        for o in optimizers:
            if not isinstance(o, FusedAdam):
                raise TypeError("GraphedTrainStep needs contrad_b200.optim.FusedAdam optimisers")
        for m in models:
            if _is_ddp(m):
                raise TypeError("GraphedTrainStep: pass the bare modules (gradient averaging is captured in the graph)")
        self.P, self.opt, self.train_fn, self.models, self.optimizers = P, opt, train_fn, models, optimizers
        self.use_warmup, self.eager_steps = use_warmup, max(1, int(eager_steps))
        self.calls, self.graph, self.recorder = 0, None, None
        self.static_images, self.static_out = None, None
        self.launches_per_replay = 0
        self.side = None
        self.prefetched = False           # the next step's host draws already sit in pinned memory

    def _step(self, images, step, schedule):
        """The step body that is warmed up, captured and replayed; `schedule` = also run the host-side LR schedule."""
        return train_step(self.P, self.opt, self.train_fn, self.models, self.optimizers, images, step,
                          use_warmup=self.use_warmup and schedule)

    def _eager(self, images, step, plan=False):
        if self.side is None:
            self.side = torch.cuda.Stream()
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            if plan:        # the last eager step also lays out the static input buffers (staging.Recorder.plan)
                self.recorder = staging.Recorder()
                with self.recorder.plan():
                    out = self._step(images, step, True)
            else:
                out = self._step(images, step, True)
        torch.cuda.current_stream().wait_stream(self.side)
        images.record_stream(self.side)
        return out

    def _host_schedule(self, step):
        if self.use_warmup:
            opt_G, opt_D = self.optimizers
            update_warmup(opt_G, step, self.opt["warmup"], self.opt["lr"])
            update_warmup(opt_D, step, self.opt["warmup"], self.opt.get("lr_d", self.opt["lr"]))

    def _capture(self, images, step):
        from . import _capi
        torch.cuda.synchronize()
        self.static_images = images.clone()
        self.graph = torch.cuda.CUDAGraph()
        before = _capi.launch_count()
        with self.recorder.capture():
            # capture on the stream the eager warm-up steps ran on: the parameters' AccumulateGrad nodes stay bound
            # to the stream of their first backward, and a mismatch would fork the capture across streams
            with torch.cuda.graph(self.graph, stream=self.side):
                self.static_out = self._step(self.static_images, step, False)
        self.launches_per_replay = _capi.launch_count() - before

    def release(self):
        """Drop the captured graph and its static tensors (required before torch.distributed.destroy_process_group
        when the graph holds NCCL kernels).  The object falls back to capturing again on the next call."""
        torch.cuda.synchronize()
        self.graph, self.static_out, self.static_images = None, None, None
        self.prefetched = False

    def __call__(self, images, step):
        from . import _capi
        self.calls += 1
        if self.calls <= self.eager_steps:
            return self._eager(images, step, plan=(self.calls == self.eager_steps))
        if self.graph is None:
            self._capture(images, step)          # records the launches, executes nothing
            _capi.add_launch_count(-self.launches_per_replay)
        self._host_schedule(step)                # Adam's scalars are staged `late`: they see this step's lr
        if not self.prefetched:
            self.recorder.produce()
        self.static_images.copy_(images, non_blocking=True)
        self.recorder.upload()
        self.graph.replay()
        _capi.add_launch_count(self.launches_per_replay)
        self.recorder.produce()                  # the NEXT step's host draws, while the GPU is busy with this one
        self.prefetched = True
        return self.static_out


class _GraphedSG2Variant(GraphedTrainStep):
    """One CUDA graph of train_step_stylegan2 for a fixed structure (with or without the R1 term) and EMA decay."""

    def __init__(self, owner, eager_steps):
        self.owner = owner
        self.P, self.opt, self.optimizers = owner.P, owner.opt, owner.optimizers
        self.models = (owner.GD.G, owner.GD.D)
        self.use_warmup, self.eager_steps = True, max(1, int(eager_steps))
        self.calls, self.graph, self.recorder = 0, None, None
        self.static_images, self.static_out = None, None
        self.launches_per_replay = 0
        self.side = owner.side
        self.prefetched = False

    def _step(self, images, step, schedule):
        o = self.owner
        return train_step_stylegan2(o.P, o.opt, o.GD, o.g_ema, o.optimizers, images, step, schedule=schedule)

    def _host_schedule(self, step):
        stylegan2_schedule(self.P, self.opt, self.optimizers, step)


class GraphedStyleGAN2Step(object):
    """``train_step_stylegan2`` replayed as CUDA graphs (same mechanics as GraphedTrainStep: eager warm-up steps on a
    side stream, a planned step that lays out the staged host inputs - crop boxes, jitter order, the style-mixing
    layer indices, Adam's scalars - then capture and replay).

    The step's STRUCTURE depends on two host decisions, so one graph is kept per value: whether the lazy R1 term is
    added this step (`step % P.d_reg_every == 0`, train_stylegan2_contraD.py:196) and the EMA decay (0 before
    `ema_start_k`, `P.accum` after; it is a by-value kernel argument).  With `--no_lazy` (config 4) and after the EMA
    start there is exactly one graph."""

    def __init__(self, P, opt, GD, g_ema, optimizers, eager_steps=3):
        from .optim import FusedAdam
        for o in optimizers:
            if not isinstance(o, FusedAdam):
                raise TypeError("GraphedStyleGAN2Step needs contrad_b200.optim.FusedAdam optimisers")
        if getattr(P, "distributed", False):
            raise NotImplementedError("GraphedStyleGAN2Step: single-process replicas only (SURVEY 8e, config 5 = DataParallel)")
        self.P, self.opt, self.GD, self.g_ema, self.optimizers = P, opt, GD, g_ema, optimizers
        self.eager_steps = eager_steps
        self.side = torch.cuda.Stream()
        self.variants = {}

    def __call__(self, images, step):
        key = (bool((step % self.P.d_reg_every == 0) and (self.P.lbd_r1 > 0)),
               float(stylegan2_ema_decay(self.P, self.opt, step)) if self.g_ema is not None else None)
        v = self.variants.get(key)
        if v is None:
            v = self.variants[key] = _GraphedSG2Variant(self, self.eager_steps)
        return v(images, step)

    @property
    def launches_per_replay(self):
        return max([v.launches_per_replay for v in self.variants.values()] or [0])

    def release(self):
        for v in self.variants.values():
            v.release()
