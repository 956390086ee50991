"""GPU, >= 2 devices (skipped on a single-GPU box; bench.py runs the same check before its timed region at every N > 1
and prints it as `parity_check`): BASELINE config 3 - the data-parallel D step over NCCL against the single-process
full-batch step on the same seeded global batch (tools/parity_multi.py; reference semantics:
third_party/gather_layer.py:8-23, training/criterion.py:30-32, training/gan/contrad.py:9-12, train_gan.py:247,311-313)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CODE = r'''
import json, os, sys
repo = sys.argv[1]
sys.path.insert(0, repo); sys.path.insert(0, os.path.join(repo, "tools")); sys.path.append(os.path.join(repo, "contrad_b200", "compat"))
import argparse, torch
import bench, parity_multi
W = bench.build_world(argparse.Namespace(no_graph=False))
dev = torch.device("cuda", W.local_rank)
out = parity_multi.check(W.P, W.G, W.D, bench.OPTIONS, int(sys.argv[2]), dev)
if W.rank == 0:
    print("RESULT " + json.dumps(out))
torch.distributed.barrier()
torch.distributed.destroy_process_group()
'''


@pytest.mark.timeout(600)
@pytest.mark.parametrize("n_global", [512, 128])
def test_distributed_d_step_equals_single_process_full_batch(n_global):
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 1 << (world.bit_length() - 1)                   # 2, 4 or 8 ranks
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29641", "-c", _CODE, REPO, str(n_global)]
    # torchrun has no `-c`: run the snippet through a temporary file
    path = os.path.join(REPO, "gpurun_out", "_parity_multi_entry.py")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(_CODE)
    cmd = cmd[:cmd.index("-c")] + [path, REPO, str(n_global)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=560)
    text = r.stdout.decode()
    assert r.returncode == 0, text[-3000:]
    res = json.loads([ln for ln in text.splitlines() if ln.startswith("RESULT ")][-1][7:])
    print(res)
    assert res["world"] == world and res["ok"], res
    for key in ("L_con_ranks_max_rel", "L_dis_rank_mean_rel", "D_grad_norm_nonlinear_xW_rel", "D_grad_norm_linear_head_rel"):
        assert res[key] < 1e-3, (key, res[key])
