"""Multi-GPU correctness of the fused exchange, under pytest: each test launches
``tests/mgpu_worker.py`` with ``python -m torch.distributed.run`` on 2 (and, where the box has them,
4 / 8) GPUs and checks the JSON line rank 0 prints.  SURVEY.md section 4 item 4."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]

ROOT = Path(__file__).resolve().parent.parent
_PORT = [29700]


def _run(world: int, *argv: str, env=None, timeout: int = 420) -> dict:
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _PORT[0] += 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]), str(ROOT / "tests" / "mgpu_worker.py"), *argv]
    e = dict(os.environ)
    e.update(env or {})
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e, cwd=str(ROOT))
    out = proc.stdout + "\n" + proc.stderr
    assert proc.returncode == 0, out[-4000:]
    for line in proc.stdout.splitlines():
        if line.startswith("MGPU_RESULT "):
            return json.loads(line[len("MGPU_RESULT "):])
    raise AssertionError("no result line:\n" + out[-4000:])


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("nvls", ["1", "0"])
def test_fused_exchange_matches_nccl_path(world, nvls):
    r = _run(world, "--case", "exchange", "--steps", "4", env={"SRB_NVLS": nvls})
    assert r["buckets"] >= 4 and r["max_abs_err_vs_nccl_path"] < 4e-3
    if nvls == "0":
        assert r["nvls"] is False


@pytest.mark.parametrize("opt", ["radam", "sgd", "adam_avg"])
def test_fused_exchange_other_optimizers(opt):
    r = _run(2, "--case", "exchange", "--steps", "8", "--opt", opt, "--scale", "4")
    assert r["opt"] == opt


def test_fused_exchange_ragged_reference_partition_and_tiny_tensors():
    # the reference's node-count partition (last rank heaviest) and 1/64-size tensors (1 KB messages)
    r = _run(2, "--case", "exchange", "--steps", "3", "--balance", "nodes", "--scale", "64", "--buckets", "12")
    assert r["balance"] == "nodes"


def test_single_bucket_equals_round1_behaviour():
    r = _run(2, "--case", "exchange", "--steps", "3", "--buckets", "1")
    assert r["buckets"] <= 2


@pytest.mark.parametrize("world", [2, 8])
def test_consumer_gate_never_reads_stale_weights(world):
    r = _run(world, "--case", "gate")
    assert r["steps"] == 3


def test_dead_peer_gives_error_code_not_hang():
    r = _run(2, "--case", "timeout", timeout=120)
    assert r["seconds"] < 15


@pytest.mark.parametrize("pipe", ["ner", "tagger"])
def test_two_rank_training_through_the_engine(pipe):
    r = _run(2, "--case", "train", "--steps", "30", "--pipe", pipe)
    assert r["buckets"] >= 2 and r["graphs"] >= 1
