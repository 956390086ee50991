#!/usr/bin/env python
"""Benchmark of the ContraD per-step training hot path (BASELINE.json metric: train-step images/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

Workload (BASELINE.json configs[1]/[2]): SNDCGAN + ContraD, `c10_b512.gin` hyper-parameters
(global batch 512, nonsat loss, Adam(2e-4, (0.5,0.999)), warm-up 3000), `--aug=simclr`, synthetic
32x32 images U[0,1), random-init weights.  One "step" = one full iteration of train_gan.py:141-179
(D step on 3N augmented images + G step, both Adam updates).  For N GPUs the global batch is split
(512 // N per rank, DDP + SyncBN(G) + the packed embedding all-gather) exactly like train_gan.py:247 ->
"scaling": "strong".

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same through
the public API with the per-step pinned-host -> device image copy and the reference's per-step loss
read-backs inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.append(os.path.join(REPO, "contrad_b200", "compat"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

GLOBAL_BATCH = 512
GIN_DEFAULTS = """
ColorJitterLayer.brightness = 0.4
ColorJitterLayer.contrast = 0.4
ColorJitterLayer.saturation = 0.4
ColorJitterLayer.hue = 0.1
RandomResizeCropLayer.scale = (0.2, 1.0)
"""
OPTIONS = {"loss": "nonsat", "warmup": 3000, "lr": 2e-4, "lr_d": 2e-4, "beta": (0.5, 0.999), "batch_size": GLOBAL_BATCH}
# algorithmic FLOPs per real image per full step (D fwd+dgrad+wgrad on 3N, G fwd/bwd, D fwd+dgrad on N)
FLOP_PER_IMAGE_STEP = 5.85e9
WORKLOAD = "SNDCGAN+ContraD CIFAR-10 32x32 b512 --aug=simclr (c10_b512.gin), synthetic images, full D+G step"


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def usable_cores():
    """Host threads the CPU baseline may use: scheduler affinity, capped by the cgroup CPU quota if there is one,
    and by 32 (torch's CPU convolutions at batch 64 stop scaling - and badly oversubscribe - beyond that)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, 32))


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) < 9:
                        continue
                    try:
                        sm.append(float(parts[1])); smax = float(parts[2])
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = sorted(sm)[len(sm) // 2:]           # upper half = samples under load
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- reference arm
def reference_available():
    from oracle import ref_import
    return os.path.isdir(os.path.join(ref_import._VENDORED, "augment")) or ref_import.reference_available()


def run_reference(args):
    """The UNMODIFIED reference (oracle/_ref, copied from /root/reference by oracle/make_ref.py) on all host cores:
    its own `get_architecture('sndcgan')`, `training.gan.contrad`, `get_augment('simclr')` modules and torch.optim.Adam
    driven through the loop body of train_gan.py:141-179 at the benchmark's batch 512 (kind "reference"; each timed
    step is one FULL b512 D+G step = the workload itself).  Falls back to the oracle port (kind "port") only when
    oracle/_ref was never built."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = usable_cores()
    if reference_available():
        from oracle import ref_runner
        r = ref_runner.run_cpu(args.steps, args.warmup, batch=GLOBAL_BATCH, threads=cores)
        value, ms, kind = r["images_per_s"], r["ms_per_step"], "reference"
        sample = "one full D+G step at batch 512 (the whole b512 workload) per timed step; " + r["what"]
        threads = r["threads"]
    else:
        value, ms, threads = _port_cpu(args.steps, args.warmup, cores)
        kind, sample = "port", "full D+G step at batch 64 (1/8 of the b512 workload) per timed step, oracle port"
    line = {"impl": "reference", "metric": "train_step_images_per_sec", "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": GLOBAL_BATCH, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def _port_cpu(steps, warmup, cores):
    from oracle import contrad_oracle as O
    torch.set_num_threads(cores)
    n = 64
    gen_w = torch.Generator().manual_seed(0)
    sd_d, sd_g = O.make_d_state(generator=gen_w), O.make_g_state(generator=gen_w)
    opt_g = O.Adam(O.trainable(sd_g).values(), 2e-4)
    opt_d = O.Adam(O.trainable(sd_d).values(), 2e-4)
    np.random.seed(0); torch.manual_seed(0)

    def one(step):
        images = torch.rand(n, 3, 32, 32)
        z_d = O.sample_latent(n); aug_d = O.sample_simclr_params(3 * n, 32, 32)
        z_g = O.sample_latent(n); aug_g = O.sample_simclr_params(n, 32, 32)
        return O.train_step(sd_g, sd_d, opt_g, opt_d, images, z_d, z_g, aug_d, aug_g, step=step)

    for w in range(warmup):
        one(w + 1)
    t0 = time.perf_counter()
    for s in range(steps):
        one(warmup + s + 1)
    dt = time.perf_counter() - t0
    return n * steps / dt, 1e3 * dt / steps, torch.get_num_threads()


def cpu_baseline_leg(seconds_budget=20.0):
    """Bounded sample for the native line: 1 warm-up + at least 2 full b512 steps of the unmodified reference on the host
    cores (stops once `seconds_budget` is used)."""
    cores = usable_cores()
    if reference_available():
        from oracle import ref_runner
        r = ref_runner.run_cpu(12, 1, batch=GLOBAL_BATCH, threads=cores, seconds_budget=seconds_budget)
        return {"value": r["images_per_s"], "unit": "images/s", "cores": r["threads"], "kind": "reference",
                "sample": "%d full D+G steps at batch 512 of the unmodified reference (oracle/_ref) on the host CPU, "
                          "torch fp32, after 1 warm-up step" % r["steps"]}
    st_np, st_t = np.random.get_state(), torch.get_rng_state()
    prev = torch.get_num_threads()
    value, ms, threads = _port_cpu(6, 1, cores)
    np.random.set_state(st_np); torch.set_rng_state(st_t); torch.set_num_threads(prev)
    return {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": "6 full D+G steps at batch 64 (1/8 of the b512 workload), oracle port, torch fp32"}


# ----------------------------------------------------------------------------------------------- this repo's arm
def build_world(args):
    import gin
    from contrad_b200 import _capi
    from contrad_b200.augment import get_augment
    from contrad_b200.models.gan import get_architecture
    from contrad_b200.training.gan import setup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    _capi.lib()                                   # fail loudly if the CUDA library is missing
    gin.clear_config()
    gin.parse_config(GIN_DEFAULTS)
    P = SimpleNamespace(mode="contrad", aug="simclr", penalty="none", temp=0.1, lbd_a=1.0, distributed=world > 1,
                        rank=rank)
    P = setup(P)
    torch.manual_seed(1234 + rank); np.random.seed(1234 + rank)
    G, D = get_architecture("sndcgan", (32, 32, 3))
    if world > 1:
        G = torch.nn.SyncBatchNorm.convert_sync_batchnorm(G)
    G.cuda(); D.cuda()
    from contrad_b200.optim import FusedAdam      # same update rule / state as torch.optim.Adam (train_gan.py:273-274)
    opt_G = FusedAdam(G.parameters(), lr=OPTIONS["lr"], betas=OPTIONS["beta"])
    opt_D = FusedAdam(D.parameters(), lr=OPTIONS["lr_d"], betas=OPTIONS["beta"])
    P.augment_fn = get_augment(mode=P.aug).cuda()
    if world > 1 and getattr(args, "no_graph", False):
        from torch.nn.parallel import DistributedDataParallel as DDP          # the reference's wrapping (train_gan.py:311-313)
        G_w = DDP(G, device_ids=[local_rank], broadcast_buffers=False)
        G_w.sample_latent = G.sample_latent
        D_w = DDP(D, device_ids=[local_rank], broadcast_buffers=False)
    else:
        if world > 1:       # graphed step: bare modules, DDP's start-up broadcast + gradient averaging done by the engine
            from contrad_b200 import engine
            engine.broadcast_parameters(G); engine.broadcast_parameters(D)
        G_w, D_w = G, D
    return SimpleNamespace(P=P, G=G_w, D=D_w, opt_G=opt_G, opt_D=opt_D, world=world, rank=rank, local_rank=local_rank)


def _log(msg):
    """Progress marker on stderr (rank-tagged): locates a hang without touching the JSON line on stdout."""
    sys.stderr.write("[bench rank %s] %s\n" % (os.environ.get("RANK", "0"), msg))
    sys.stderr.flush()


def run_native(args):
    import faulthandler
    from contrad_b200 import _capi, engine, kernels as K
    # a hung collective / capture must not burn the whole time limit: dump every thread's stack and exit
    faulthandler.dump_traceback_later(int(os.environ.get("CB200_BENCH_WATCHDOG", "900")), exit=True)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a path has no CPU fallback (use --impl reference)")
    W = build_world(args)
    world, rank = W.world, W.rank
    n_local = GLOBAL_BATCH // world
    dev = torch.device("cuda", W.local_rank if world > 1 else 0)
    train_fn = W.P.train_fn
    step_no = [0]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def eager_step(images):
        step_no[0] += 1
        return engine.train_step(W.P, OPTIONS, train_fn, (W.G, W.D), (W.opt_G, W.opt_D), images, step_no[0])

    if args.no_graph:
        one_step = eager_step
    else:
        # the public fast loop: the same train_step captured once into a CUDA graph and replayed (engine.py)
        graphed = engine.GraphedTrainStep(W.P, OPTIONS, train_fn, (W.G, W.D), (W.opt_G, W.opt_D))

        def one_step(images):
            step_no[0] += 1
            return graphed(images, step_no[0])

    gen = torch.Generator(device=dev).manual_seed(rank)
    pool = [torch.rand(n_local, 3, 32, 32, device=dev, generator=gen) for _ in range(4)]

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(steps):
            fn(s)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- W > 1: the distributed D step against the single-process full-batch step on the same seeded batch
    parity = None
    if world > 1 and not args.no_parity_check:
        sys.path.insert(0, os.path.join(REPO, "tools"))
        import parity_multi
        parity = parity_multi.check(W.P, W.G, W.D, OPTIONS, GLOBAL_BATCH, dev)
        _log("parity check done: %s" % (parity,))

    # ---- device-resident arm
    _log("world built; set-up steps")
    if not args.no_graph:                         # untimed set-up: eager allocator warm-up steps + the capture
        for w in range(graphed.eager_steps + 1):
            one_step(pool[w % len(pool)])
    torch.cuda.synchronize()
    _log("set-up done (graph captured: %s); warm-up" % (not args.no_graph))
    for w in range(args.warmup):
        one_step(pool[w % len(pool)])
    torch.cuda.synchronize()
    _log("timed region")
    clocks = ClockSampler(W.local_rank if world > 1 else 0)
    if rank == 0:
        clocks.start()
    _capi.reset_launch_count()
    ms = timed(lambda s: one_step(pool[s % len(pool)]), args.steps)
    launches = _capi.launch_count()
    clk = clocks.stop() if rank == 0 else {}
    value = GLOBAL_BATCH * args.steps / (ms / 1e3)

    _log("device-resident arm done: %.3f ms/step" % (ms / args.steps))
    # ---- end-to-end arm: pinned host images -> device every step + the reference's per-step loss read-backs
    host_pool = [torch.rand(n_local, 3, 32, 32).pin_memory() for _ in range(4)]
    d2h = [0]

    def e2e_step(s):
        images = host_pool[s % len(host_pool)].to(dev, non_blocking=True)
        out = one_step(images)
        vals = [out[k].item() for k in ("g_loss", "d_loss", "d_penalty", "d_real", "d_gen")]   # train_gan.py:164-167,179
        d2h[0] = 4 * len(vals)
        return vals

    for w in range(2):
        e2e_step(w)
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = GLOBAL_BATCH * args.steps / (ms_e2e / 1e3)

    _log("e2e arm done: %.3f ms/step" % (ms_e2e / args.steps))
    # ---- optional: the same end-to-end loop fed with the dataset's raw uint8 images (ToTensor, the
    # two-view duplication and the concatenation run inside the augmentation kernel; 1 B/element crosses PCIe)
    e2e_u8 = None
    if args.u8_input and world == 1:
        try:
            graphed_u8 = None if args.no_graph else engine.GraphedTrainStep(W.P, OPTIONS, train_fn, (W.G, W.D),
                                                                            (W.opt_G, W.opt_D))
            host_u8 = [torch.randint(0, 256, (n_local, 3, 32, 32), dtype=torch.uint8).pin_memory() for _ in range(4)]

            def u8_step(s):
                images = host_u8[s % len(host_u8)].to(dev, non_blocking=True)
                step_no[0] += 1
                out = graphed_u8(images, step_no[0]) if graphed_u8 is not None else engine.train_step(
                    W.P, OPTIONS, train_fn, (W.G, W.D), (W.opt_G, W.opt_D), images, step_no[0])
                return [out[k].item() for k in ("g_loss", "d_loss", "d_penalty", "d_real", "d_gen")]

            for w in range((graphed_u8.eager_steps + 1 if graphed_u8 is not None else 0) + 2):
                u8_step(w)
            ms_u8 = timed(u8_step, args.steps)
            e2e_u8 = {"value": GLOBAL_BATCH * args.steps / (ms_u8 / 1e3), "unit": "images/s", "ms_per_step": ms_u8 / args.steps,
                      "h2d_bytes_per_step": n_local * 3 * 32 * 32, "d2h_bytes_per_step": d2h[0]}
            if graphed_u8 is not None:
                graphed_u8.release()
        except Exception as e:                                # noqa: BLE001 - optional leg: report, do not lose the line
            e2e_u8 = {"error": "%s: %s" % (type(e).__name__, e)}
    # ---- the strict precision modes (contrad_b200/precision.py) on the same workload: what 1e-3 on the generator's
    # gradient norm costs (N = 1 only; the headline `value` above is the default single-pass TF32 mode)
    modes = None
    if world == 1 and not args.no_precision_modes and not args.no_graph:
        from contrad_b200 import precision
        modes = {"default": {"value": value, "ms_per_step": ms / args.steps,
                             "what": "single-pass TF32 (the reference's own GPU arithmetic class)"}}
        for mode, what in (("strict", "3xTF32 operands in the generator step"), ("full", "3xTF32 operands in both steps")):
            try:
                precision.set_strict(mode)
                g2 = engine.GraphedTrainStep(W.P, OPTIONS, train_fn, (W.G, W.D), (W.opt_G, W.opt_D))

                def step2(s, g2=g2):
                    step_no[0] += 1
                    return g2(pool[s % len(pool)], step_no[0])

                for w in range(g2.eager_steps + 1 + 3):
                    step2(w)
                ms2 = timed(step2, 10)
                modes[mode] = {"value": GLOBAL_BATCH * 10 / (ms2 / 1e3), "ms_per_step": ms2 / 10, "what": what}
                g2.release()
            except Exception as e:                            # noqa: BLE001
                modes[mode] = {"error": "%s: %s" % (type(e).__name__, e)}
            finally:
                precision.set_strict(False)
        _log("precision modes done: %s" % ({k: v.get("ms_per_step") for k, v in modes.items()},))
    # ---- the north_star denominator: the UNMODIFIED reference (oracle/_ref) on the same GPU(s), PyTorch eager
    eager = None
    if not args.no_eager_baseline and reference_available():
        if not args.no_graph:
            graphed.release()                      # free the graph's pool (and its NCCL kernels) before DDP starts
        try:
            eager = (eager_gpu_baseline_leg(world, W.local_rank, dev, steps=30, warmup=8) if world > 1
                     else eager_gpu_baseline_leg(world, 0, dev))
            _log("eager reference leg done: %s" % (eager.get("ms_per_step", eager.get("error")),))
        except Exception as e:                                # noqa: BLE001 - report, do not lose the line
            eager = {"error": "%s: %s" % (type(e).__name__, e)}
            _log("eager reference leg failed: %s" % eager["error"])
    faulthandler.cancel_dump_traceback_later()
    line = None
    if rank == 0:
        peaks = load_peaks()
        roof = roofline_legs(K, engine, W, eager_step, pool, peaks) if world == 1 else None
        cpu = cpu_baseline_leg() if world == 1 and not args.no_cpu_baseline else None
        side = side_workloads() if world == 1 and not args.no_side_workloads else None
        line = {
            "metric": "train_step_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": GLOBAL_BATCH, "per_gpu_batch": n_local,
                       "parallelism": "dp%d" % world, "image": "3x32x32",
                       "launch": "eager" if args.no_graph else "cuda-graph replay of the whole D+G step",
                       "l2": "per-step working set ~1.3 GB of activations >> 126 MB L2 (no explicit flush needed)"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": n_local * 3 * 32 * 32 * 4, "d2h_bytes_per_step": d2h[0]},
            "gpu_launches": int(launches),
            "step_tflops": FLOP_PER_IMAGE_STEP * value / 1e12,
        }
        if e2e_u8 is not None:
            line["e2e_uint8_input"] = e2e_u8
        if modes is not None:
            line["precision_modes"] = modes
        if parity is not None:
            line["parity_check"] = parity
        if eager is not None:
            if "value" in eager:
                eager["ratio"] = value / eager["value"]            # device-resident value / reference eager
                eager["ratio_e2e"] = e2e_value / eager["value"]    # like for like: both include H2D + 5 read-backs
                eager["target"] = ">= 10x at 1 GPU, >= 6x at 8 GPUs (north_star)"
            line["eager_gpu_baseline"] = eager
        if roof:
            line.update(roof)
        if cpu:
            line["cpu_baseline"] = cpu
        if side:
            line["other_workloads"] = side
    if line is not None:
        print(json.dumps(line), flush=True)
    _log("teardown")
    if not args.no_graph:
        graphed.release()                          # graphs holding NCCL kernels must go before the communicator
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    _log("done")
    return 0


def eager_gpu_baseline_leg(world, local_rank, dev, steps=50, warmup=10, budget_s=240):
    """`oracle/_ref/train_gan.py`'s own `train()` loop on unmodified reference modules, `.cuda()`, PyTorch default
    flags, DDP + SyncBatchNorm wrapping exactly as `worker()` does (train_gan.py:268-271,311-313), `steps` steps after
    `warmup` warm-up steps, per-rank batch 512 // N.  Every rank runs it in a CHILD process with its own NCCL
    rendezvous (MASTER_PORT + 23): the reference leg can neither inherit communicator / graph state from the native arm
    nor take the native line down - a child that exceeds `budget_s` is killed and the leg reports the failure.  The step
    time is the max over ranks."""
    env = dict(os.environ)
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    env["MASTER_PORT"] = str(int(env.get("MASTER_PORT", "29500")) + 23)
    env.setdefault("RANK", "0"); env.setdefault("WORLD_SIZE", "1"); env.setdefault("LOCAL_RANK", str(local_rank))
    for k in [k for k in env if k.startswith("TORCHELASTIC_")]:      # rank 0 of the children hosts its own TCP store
        env.pop(k)
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "eager-child", "--steps", str(steps), "--warmup", str(warmup)]
    ms, info, err = float("inf"), {}, None
    try:
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=budget_s)
        lines = [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
        if r.returncode == 0 and lines:
            info = json.loads(lines[-1])
            ms = float(info["ms_per_step"])
        else:
            err = "child exit %d: %s" % (r.returncode, r.stderr.decode()[-400:].replace("\n", " | "))
    except subprocess.TimeoutExpired:
        err = "reference leg exceeded %d s and was killed" % budget_s
    t = torch.tensor([ms if ms != float("inf") else 1e30], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    if ms >= 1e29:
        return {"error": err or "the reference leg failed on another rank", "n_gpus": world, "kind": "reference"}
    return {"value": GLOBAL_BATCH / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms, "n_gpus": world,
            "steps": steps, "warmup": warmup, "per_gpu_batch": info.get("per_gpu_batch"), "flags": info.get("flags"),
            "kind": "reference", "what": info.get("what")}


def run_eager_child(args):
    """Child of eager_gpu_baseline_leg: one rank of the unmodified reference's training loop on its GPU."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from oracle import ref_runner
    r = ref_runner.run_gpu(args.steps, args.warmup, global_batch=GLOBAL_BATCH, local_rank=local_rank)
    print(json.dumps(r), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


def side_workloads():
    """Not part of the headline metric: BASELINE config 4 (StyleGAN2 small32 + ContraD, c10_style64.gin: b64, R1 every
    step) through engine.GraphedStyleGAN2Step on the same GPU, so that the StyleGAN2 rows of the hot path (SURVEY 8a
    a18-a22) carry a measured images/s in the same artefact.  Failures are reported, never raised."""
    try:
        sys.path.insert(0, os.path.join(REPO, "tools"))
        import bench_sg2
        res = bench_sg2.measure(batch=64, steps=12, warmup=3, d_reg_every=1, graph=True)
        res.pop("losses", None)
        return {"stylegan2_config4": res}
    except Exception as e:                                    # noqa: BLE001
        return {"stylegan2_config4": {"error": "%s: %s" % (type(e).__name__, e)}}


def roofline_legs(K, engine, W, one_step, pool, peaks):
    """Per-kernel-family device time measured with CUDA events around every C-ABI call during 3 extra,
    instrumented steps (same stream); the dominant family gives `roofline`.  Plus the fused-augment HBM
    point at a saturating size (B = 65536 images of 32x32, 805 MB in)."""
    records = []
    K.set_profile_hook(records)
    for s in range(3):
        one_step(pool[s % len(pool)])
    torch.cuda.synchronize()
    K.set_profile_hook(None)
    fam = {}
    for name, e0, e1, flops, nbytes in records:
        d = fam.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += e0.elapsed_time(e1); d["flops"] += flops; d["bytes"] += nbytes; d["launches"] += 1
    total_ms = sum(d["ms"] for d in fam.values()) or 1.0
    tc = {k: v for k, v in fam.items() if v["flops"] > 0}
    out = {"kernel_time_share": {k: round(v["ms"] / total_ms, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}}
    if tc:
        name, d = max(tc.items(), key=lambda kv: kv[1]["ms"])
        achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
        # TF32 peak = half the measured bf16 BURST figure: the instrumented kernels run at the full 1965 MHz in short
        # bursts (the sustained figure was taken at a power-capped ~1.3 GHz) - VERDICT r1 item 4
        peak_tf32 = peaks["bf16_tflops"] / 2.0
        out["roofline"] = {"kernel": name, "bound": "tensor", "achieved": achieved, "peak": peak_tf32, "unit": "TFLOP/s",
                           "frac": achieved / peak_tf32, "traffic": load_traffic(name),
                           "launches_per_step": d["launches"] / 3.0, "avg_launch_ms": d["ms"] / d["launches"],
                           "peak_note": "TF32 peak taken as %s bf16 BURST (%.1f TF/s) / 2; against the sustained figure (%.1f / 2) "
                                        "frac = %.3f" % (peaks["source"], peaks["bf16_tflops"], peaks["bf16_tflops_sustained"],
                                                         achieved / (peaks["bf16_tflops_sustained"] / 2.0))}
        # the profiled instance of that family (tools/profile_target.py "dgrad": 3x3, 128 -> 128 channels, 16x16,
        # B = 1536), timed live: this is the launch the committed `ncu --set full` capture and `traffic` refer to
        inst = roofline_instance(K)
        if inst:
            out["roofline"].update({"instance": inst["what"], "instance_achieved": inst["tflops"],
                                    "instance_frac": inst["tflops"] / peak_tf32, "instance_ms": inst["ms"],
                                    "instance_algorithmic_bytes": inst["bytes"], "traffic": load_traffic("dgrad")})
        all_flops = sum(v["flops"] for v in tc.values()); all_ms = sum(v["ms"] for v in tc.values())
        out["tensor_kernels"] = {"achieved_tflops": all_flops / (all_ms * 1e-3) / 1e12,
                                 "frac_of_tf32_peak": all_flops / (all_ms * 1e-3) / 1e12 / peak_tf32}
    # fused augment at the HBM-saturating size
    from contrad_b200.augment.layers import FusedSimCLR  # noqa: F401
    B = 65536
    x = torch.rand(B, 3, 32, 32, device="cuda")
    params, order = W.P.augment_fn.sample_params(x)
    for _ in range(3):
        K.augment_simclr_fwd(x, params, order)
    evs = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.augment_simclr_fwd(x, params, order); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    gbs = 8.0 * x.numel() / (ms * 1e-3) / 1e9
    out["roofline_augment"] = {"kernel": "augment_simclr_fwd", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                               "traffic": load_traffic("augment_v2") or load_traffic("augment"),
                               "size": "B=65536 x 3x32x32 fp32 (805 MB in, 805 MB out), 8 algorithmic B/element",
                               "peak_note": "%s copy bandwidth" % peaks["source"]}
    # the any-size path of the same chain at config 5's image size (two launches: per-image means, then apply), reported
    # next to the small-image kernel; optional leg - a failure is recorded, it cannot take the line down
    try:
        Bl = 48
        imgs = [torch.rand(Bl, 3, 512, 512, device="cuda") for _ in range(3)]      # 151 MB each: rotated, > L2
        prm = torch.zeros(11, Bl, device="cuda")
        prm[0] = 0.7; prm[1] = 0.8; prm[2] = 0.1; prm[3] = -0.1; prm[4] = 1.0; prm[5] = 1.0; prm[6] = 1.2; prm[7] = 0.05
        prm[8] = 1.1; prm[9] = 0.9
        prm[4, ::2] = -1.0; prm[10, ::5] = 1.0
        for i in range(3):
            K.augment_simclr_large_fwd(imgs[i], prm, 0)
        evs = []
        for i in range(9):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); K.augment_simclr_large_fwd(imgs[i % 3], prm, 0); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ms_l = float(np.median([a.elapsed_time(b) for a, b in evs]))
        gbs_l = 8.0 * imgs[0].numel() / (ms_l * 1e-3) / 1e9
        out["roofline_augment_large"] = {"kernel": "augment_large_mean + augment_large_apply", "bound": "hbm", "achieved": gbs_l,
                                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs_l / peaks["hbm_gbs"],
                                         "traffic": load_traffic("augment_large"), "ms": ms_l,
                                         "size": "B=48 x 3x512x512 fp32, 8 algorithmic B/element, colour jitter on every image"}
        del imgs
    except Exception as e:                                    # noqa: BLE001
        out["roofline_augment_large"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out


def roofline_instance(K):
    """The dominant kernel on the shape the committed ncu capture uses (profiles/prof_r1_dgrad.md)."""
    try:
        B, H, Cin, Cout = 1536, 16, 128, 128
        x = K.round_tf32(torch.randn(B, H, H, Cin, device="cuda"))
        w = K.round_tf32(torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.02)
        wt = K.pack_dgrad_weight(w, 1)
        dy = K.round_tf32(torch.randn(B, H, H, Cout, device="cuda"))
        run = lambda: K.conv2d_nhwc_dgrad(dy, wt, (B, H, H, Cin), 3, 1, act_in=x, slope=0.1, round_out=True)
        for _ in range(3):
            run()
        times = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        flops = 2.0 * B * H * H * Cin * 9 * Cout
        return {"what": "tap_gemm_persist_kernel<128,2,4>: data gradient of a 3x3 conv, 128->128 ch, 16x16, B=1536 "
                        "(dY 201 MB + activation mask 201 MB in, 201 MB out: operands > L2)",
                "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12, "bytes": 3 * 4 * B * H * H * Cin}
    except Exception as exc:      # the roofline leg must never take the bench line down
        sys.stderr.write("roofline_instance failed: %r\n" % (exc,))
        return None


def load_traffic(kernel):
    """dram bytes per launch from the committed ncu summary (profiles/traffic.json), or None."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f).get(kernel)
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "eager-child"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-workloads", action="store_true", help="skip the StyleGAN2 config-4 side measurement")
    ap.add_argument("--no-u8-input", dest="u8_input", action="store_false",
                    help="skip the extra end-to-end leg fed with uint8 host images (row f3; N=1 only, e2e_uint8_input)")
    ap.add_argument("--u8-input", dest="u8_input", action="store_true", help=argparse.SUPPRESS)      # default on
    ap.add_argument("--no-eager-baseline", action="store_true",
                    help="skip the reference-PyTorch-eager-on-this-GPU leg (eager_gpu_baseline)")
    ap.add_argument("--no-precision-modes", action="store_true", help="skip the strict / full precision-mode timings (N = 1)")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the W > 1 parity check before the timed region")
    ap.add_argument("--no-graph", action="store_true", help="eager launches (and DDP wrappers for N > 1) instead of the CUDA-graph step")
    ap.set_defaults(u8_input=True)
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 behind Python's back (NCCL prints
    # "NCCL version ..." there when NCCL_DEBUG is set in the environment) are routed to stderr for the whole run.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "eager-child":
        return run_eager_child(args)
    args.warmup = max(args.warmup, 3)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
