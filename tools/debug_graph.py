"""Debug helper: graphed vs eager G-step intermediates."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from types import SimpleNamespace
from contrad_b200 import engine, _capi
from contrad_b200.training.gan import contrad
args = SimpleNamespace(gpus=1, steps=10, warmup=3, no_cpu_baseline=True, no_graph=False)
W = bench.build_world(args)
keep = {}
def loss_G(P, D, options, images, gen_images):
    keep["gen_images"] = gen_images
    aug = P.augment_fn(gen_images)
    keep["aug"] = aug
    d_gen = D(aug)
    keep["d_gen"] = d_gen
    from contrad_b200.functional import GanGLossFn
    out = GanGLossFn.apply(d_gen, options["loss"])
    keep["g_loss"] = out
    return out
train_fn = {"D": contrad.loss_D_fn, "G": loss_G}
gs = engine.GraphedTrainStep(W.P, bench.OPTIONS, train_fn, (W.G, W.D), (W.opt_G, W.opt_D))
pool = [torch.rand(512, 3, 32, 32, device="cuda") for _ in range(2)]
for s in range(6):
    out = gs(pool[s % 2], s + 1)
    torch.cuda.synchronize()
    print(s, {k: float(v) for k, v in out.items()})
    if gs.recorder is not None:
        for i, e in enumerate(gs.recorder.entries):
            b = e.device_buf
            print("    entry %d shape %s mean %.5f absmax %.5f first %s" % (i, tuple(b.shape), float(b.float().mean()), float(b.abs().max()), b.flatten()[:4].tolist()))
    gd = torch.stack([p.grad.norm() for p in W.D.parameters() if p.grad is not None]) if any(p.grad is not None for p in W.D.parameters()) else torch.zeros(1)
    gg = torch.stack([p.grad.norm() for p in W.G.parameters() if p.grad is not None]) if any(p.grad is not None for p in W.G.parameters()) else torch.zeros(1)
    print("    D grad norms max %.4g  G grad norms max %.4g  D param max %.4g G param max %.4g" % (float(gd.max()), float(gg.max()),
          max(float(p.abs().max()) for p in W.D.parameters()), max(float(p.abs().max()) for p in W.G.parameters())))
    for k, v in keep.items():
        v = v.detach().float()
        print("    %-10s shape %s mean %.5f absmax %.5f nan %d" % (k, tuple(v.shape), float(v.mean()), float(v.abs().max()), int(torch.isnan(v).sum())))
# timing: host enqueue per step in both modes
for name, fn in (("graph", lambda im, s: gs(im, s)),
                 ("eager", lambda im, s: engine.train_step(W.P, bench.OPTIONS, W.P.train_fn, (W.G, W.D), (W.opt_G, W.opt_D), im, s))):
    for i in range(3): fn(pool[i % 2], 10 + i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(10): fn(pool[i % 2], 20 + i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("%s: host enqueue %.2f ms/step, total %.2f ms/step" % (name, (t1 - t0) * 100, (t2 - t0) * 100))
t0 = time.perf_counter()
for i in range(10): gs.recorder.refresh()
torch.cuda.synchronize()
print("refresh only: %.2f ms" % ((time.perf_counter() - t0) * 100))
import os
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "loadavg", os.getloadavg())
