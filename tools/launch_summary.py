"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel device time over one
training step (the window between the 1st and 3rd fused-augment forward launches = D step + G step)."""
import collections
import csv
import re
import sys


def short(k):
    k = k.replace("void ", "")
    m = re.search(r"([A-Za-z_][A-Za-z0-9_]*)\s*(<[^()]*>)?\s*\(", k)
    if not m:
        return k[:50]
    name, targs = m.group(1), (m.group(2) or "")
    if name in ("tap_gemm_kernel", "wgrad_kernel") or name.startswith("augment_simclr"):
        name += targs
    if name in ("vectorized_elementwise_kernel", "elementwise_kernel", "Kernel", "multi_tensor_apply_kernel"):
        f = re.search(r"at::(?:native::)?(?:<unnamed>::)?([A-Za-z_]+(?:Functor|_kernel_cuda|kernel)[A-Za-z_]*)", k)
        name += ":" + (f.group(1) if f else k[k.find("<"):][:40])
    return name


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for row in r:
        try:
            seq.append((short(row[ki]), float(row[vi].replace(",", "")), row[ki]))
        except Exception:
            pass
    idx = [n for n, (k, v, _) in enumerate(seq) if k.startswith("augment_simclr_fwd")]
    if len(idx) >= 3:
        a, b = idx[0], idx[2]
        # the no-grad G forward of the D step precedes the augment launch: start at the previous optimizer tail
        step = seq[a:b]
        label = "one step window (launches %d..%d)" % (a, b)
    else:
        step, label = seq, "all captured launches"
    tot = collections.defaultdict(lambda: [0, 0.0])
    for k, v, _ in step:
        tot[k][0] += 1
        tot[k][1] += v
    T = sum(v for _, v in tot.values())
    mine = ("tap_gemm", "wgrad_kernel", "augment", "sn_", "conv_first", "contrastive", "rownorm", "gan_", "colsum",
            "lrelu_bwd", "sum_kernel", "bn_", "g_final")
    tm = sum(v for k, (n, v) in tot.items() if k.startswith(mine))
    print("%s: %d launches, %.1f us device time; contrad_b200 kernels %.1f us (%.1f%%)" % (label, len(step), T / 1e3, tm / 1e3, 100 * tm / T))
    for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:50]:
        print("%-64s n=%4d %9.1f us %5.1f%%" % (k[:64], n, v / 1e3, 100 * v / T))


if __name__ == "__main__":
    main(sys.argv[1])
