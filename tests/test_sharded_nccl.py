"""Row-sharded path over NCCL on real GPUs (-m gpu, needs >= 2 of them; skipped on a 1-GPU box): torchrun launches
scripts/sharded_nccl_check.py on 2 ranks -- ShardedKnnIndex over B200 shards (delta routing, one all-gather of the
per-shard (key, distance) top-k, merge) against ONE CPU oracle through adds / swap-removes, then the device tier
(pwknn_search_device_packed -> NCCL all-gather -> pwknn_merge_shards_device) against a torch stable-sort merge, including
a cross-position exact tie.  Also the bench's own N = 2 parity guard on a reduced corpus.
Reference order: global (OrderedFloat(distance), slot), brute_force_knn_integration.rs:220-224."""

import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(n, script_args, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port())] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_sharded_index_over_nccl_matches_the_oracle():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _torchrun(2, [os.path.join(ROOT, "scripts", "sharded_nccl_check.py")], 600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("matches the oracle") == 4 and out.stdout.count("matches the torch merge") == 2, out.stdout


def test_bench_parity_guard_at_two_gpus():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--gpus", "2", "--rows", "1000000", "--steps", "5", "--warmup", "3",
                        "--extras", "none", "--no-one-handle"], 900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["parity"]["ok"] and line["parity"]["ids_equal"]
    assert line["parity"]["oracle_subcorpus"]["ids_equal"] and line["e2e"]["value"] > 0
