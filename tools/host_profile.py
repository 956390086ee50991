"""Host-side (Python) profile of the train step: where do the CPU microseconds go?"""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from types import SimpleNamespace
args = SimpleNamespace(gpus=1, steps=10, warmup=3, no_cpu_baseline=True, no_graph=True)
from contrad_b200 import engine
W = bench.build_world(args)
pool = [torch.rand(512, 3, 32, 32, device="cuda") for _ in range(2)]
step = [0]
def one(images):
    step[0] += 1
    return engine.train_step(W.P, bench.OPTIONS, W.P.train_fn, (W.G, W.D), (W.opt_G, W.opt_D), images, step[0])
for i in range(5): one(pool[i % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10): one(pool[i % 2])
t1 = time.perf_counter()          # host time to ENQUEUE 10 steps (no sync)
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue time per step: %.2f ms ; incl. GPU drain: %.2f ms" % ((t1 - t0) * 100, (t2 - t0) * 100))
pr = cProfile.Profile(); pr.enable()
for i in range(10): one(pool[i % 2])
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35); print(s.getvalue()[:6000])
