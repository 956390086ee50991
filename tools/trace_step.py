#!/usr/bin/env python
"""Kernel timeline of the graphed train step on rank 0 (works under torchrun at any N; no nsys in the image).

    python -m torch.distributed.run --nproc-per-node 8 ... tools/trace_step.py --out gpurun_out/trace_n8 [--global-batch 512]

torch.profiler (Kineto / CUPTI activity records) sees the kernels of a CUDA-graph replay, NCCL kernels included, with
start time, duration and stream.  The script replays the step a few times under the profiler on EVERY rank (the
collectives need all of them) and rank 0 writes
  <out>.json  - per-kernel-name totals per step, NCCL totals, busy time per stream, union busy time, idle gaps
  <out>.md    - the same as a table (copied into profiles/ by hand)
Numbers taken under a profiler are not bench values; they give SHARES and gaps.
"""
import argparse
import json
import os
import re
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.append(os.path.join(REPO, "contrad_b200", "compat"))

import torch  # noqa: E402

import bench  # noqa: E402


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("(anonymous namespace)::", "").replace("contrad_b200::", "")
    name = re.sub(r"\(.*$", "", name)
    return name[:100]


def summarise(trace_path, steps):
    with open(trace_path) as f:
        tr = json.load(f)
    evs = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    evs.sort(key=lambda e: e["ts"])
    if not evs:
        return {"error": "no device events in the trace"}
    t0, t1 = evs[0]["ts"], max(e["ts"] + e["dur"] for e in evs)
    per = {}
    streams = {}
    for e in evs:
        k = short(e["name"])
        d = per.setdefault(k, {"us": 0.0, "n": 0})
        d["us"] += e["dur"]; d["n"] += 1
        s = str(e.get("args", {}).get("stream", "?"))
        streams.setdefault(s, 0.0)
        streams[s] += e["dur"]
    # union of busy intervals over all streams
    busy, cur_s, cur_e = 0.0, None, None
    gaps = []
    for e in evs:
        s, en = e["ts"], e["ts"] + e["dur"]
        if cur_e is None:
            cur_s, cur_e = s, en
        elif s <= cur_e:
            cur_e = max(cur_e, en)
        else:
            gaps.append(s - cur_e)
            busy += cur_e - cur_s
            cur_s, cur_e = s, en
    busy += cur_e - cur_s
    nccl = {k: v for k, v in per.items() if "nccl" in k.lower()}
    wall = t1 - t0
    big_gaps = sorted(gaps, reverse=True)[:max(steps - 1, 0)]      # the gaps between replays (host side), excluded below
    inner_gap = sum(gaps) - sum(big_gaps)
    # per-launch list of the LAST replay (name, grid, duration): the launch list of one step
    last_start = evs[-1]["ts"]
    big = sorted(((evs[i + 1]["ts"] - (evs[i]["ts"] + evs[i]["dur"]), i) for i in range(len(evs) - 1)), reverse=True)[:max(steps - 1, 0)]
    cut = max([i for _, i in big], default=-1) + 1
    launches = [{"name": short(e["name"]), "us": e["dur"], "grid": e.get("args", {}).get("grid"),
                 "stream": e.get("args", {}).get("stream"), "t_us": e["ts"] - evs[cut]["ts"]} for e in evs[cut:]]
    out = {
        "launches_last_step": launches,
        "steps": steps, "wall_us_per_step": (wall - sum(big_gaps)) / steps, "busy_us_per_step": busy / steps,
        "idle_inside_step_us": inner_gap / steps, "n_device_events_per_step": len(evs) / steps,
        "nccl_us_per_step": sum(v["us"] for v in nccl.values()) / steps,
        "nccl_calls_per_step": sum(v["n"] for v in nccl.values()) / steps,
        "stream_busy_us_per_step": {s: v / steps for s, v in streams.items()},
        "kernels": {k: {"us_per_step": v["us"] / steps, "n_per_step": v["n"] / steps}
                    for k, v in sorted(per.items(), key=lambda kv: -kv[1]["us"])},
    }
    return out


def to_md(res, title):
    lines = ["# %s" % title, "",
             "Taken with torch.profiler (CUPTI activity records) on rank 0 over %d graph replays; times under a profiler are for"
             " SHARES and gaps, not bench values." % res["steps"], "",
             "| quantity | per step |", "|---|---|",
             "| device wall (first kernel start to last kernel end, replay-to-replay host gaps removed) | %.1f us |" % res["wall_us_per_step"],
             "| union busy time over all streams | %.1f us |" % res["busy_us_per_step"],
             "| idle inside the step (no kernel on any stream) | %.1f us |" % res["idle_inside_step_us"],
             "| device events (kernels + memcpy/memset) | %.0f |" % res["n_device_events_per_step"],
             "| NCCL kernels | %.0f calls, %.1f us |" % (res["nccl_calls_per_step"], res["nccl_us_per_step"]),
             "", "| stream | busy us / step |", "|---|---|"]
    for s, v in sorted(res["stream_busy_us_per_step"].items(), key=lambda kv: -kv[1]):
        lines.append("| %s | %.1f |" % (s, v))
    lines += ["", "| kernel | launches / step | us / step | share of busy |", "|---|---|---|---|"]
    tot = sum(v["us_per_step"] for v in res["kernels"].values()) or 1.0
    for k, v in list(res["kernels"].items())[:60]:
        lines.append("| `%s` | %.1f | %.1f | %.1f %% |" % (k, v["n_per_step"], v["us_per_step"], 100 * v["us_per_step"] / tot))
    return "\n".join(lines) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/trace")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--global-batch", type=int, default=bench.GLOBAL_BATCH)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    bench.GLOBAL_BATCH = args.global_batch
    from contrad_b200 import engine
    W = bench.build_world(argparse.Namespace(no_graph=args.no_graph))
    world, rank = W.world, W.rank
    n_local = args.global_batch // world
    dev = torch.device("cuda", W.local_rank if world > 1 else 0)
    step_no = [0]
    if args.no_graph:
        def one(images):
            step_no[0] += 1
            return engine.train_step(W.P, bench.OPTIONS, W.P.train_fn, (W.G, W.D), (W.opt_G, W.opt_D), images, step_no[0])
        graphed = None
    else:
        graphed = engine.GraphedTrainStep(W.P, bench.OPTIONS, W.P.train_fn, (W.G, W.D), (W.opt_G, W.opt_D))

        def one(images):
            step_no[0] += 1
            return graphed(images, step_no[0])
    pool = [torch.rand(n_local, 3, 32, 32, device=dev) for _ in range(2)]
    for w in range(8):
        one(pool[w % 2])
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for s in range(args.steps):
            one(pool[s % 2])
            torch.cuda.synchronize()
    if rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        fd, path = tempfile.mkstemp(suffix=".json"); os.close(fd)
        prof.export_chrome_trace(path)
        res = summarise(path, args.steps)
        os.unlink(path)
        res["config"] = {"world": world, "global_batch": args.global_batch, "per_gpu_batch": n_local,
                         "graph": not args.no_graph}
        with open(args.out + ".json", "w") as f:
            json.dump(res, f, indent=1)
        with open(args.out + ".md", "w") as f:
            f.write(to_md(res, "Kernel timeline of one train step, N=%d GPUs, %d images per rank (%s)"
                          % (world, n_local, "eager" if args.no_graph else "CUDA-graph replay")))
        print(json.dumps({k: v for k, v in res.items() if k not in ("kernels", "launches_last_step")}))
    if graphed is not None:
        graphed.release()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
