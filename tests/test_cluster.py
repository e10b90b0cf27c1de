"""Multi-node actor runtime (``parallel/cluster.py``) with two "nodes" on this machine: the test process is the
head, a subprocess is the second node's agent."""
import os
import socket
import subprocess
import sys
import time
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class Echo:
    def __init__(self, tag):
        self.tag = tag
        self.peer = None

    def whoami(self):
        return self.tag, os.getpid(), os.getppid()

    def set_peer(self, peer):
        self.peer = peer
        return True

    def ask_peer(self):
        from spacy_ray_b200.parallel import actors

        return actors.get(self.peer.whoami.remote())[0]

    def big(self, n):
        return bytes(n)

    def die(self):
        os._exit(3)


def _agent(port):
    env = dict(os.environ, PYTHONPATH=f"{ROOT}{os.pathsep}{ROOT / 'tests'}{os.pathsep}" + os.environ.get("PYTHONPATH", ""))
    return subprocess.Popen([sys.executable, "-m", "spacy_ray_b200", "ray", "node", "--address", f"127.0.0.1:{port}"],
                            env=env, cwd=str(ROOT), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def test_actors_on_two_nodes_call_each_other_and_report_deaths():
    from spacy_ray_b200.parallel import actors

    port = _free_port()
    agent = _agent(port)
    try:
        actors.init(address=f"127.0.0.1:{port}", nodes=2, accept_timeout=60)
        R = actors.remote(Echo)
        a, b, c = R.remote("a"), R.remote("b"), R.remote("c")          # even fill: a -> node 0, b -> node 1, c -> node 0
        info = {h: actors.get(h.whoami.remote(), timeout=60) for h in (a, b, c)}
        assert [info[h][0] for h in (a, b, c)] == ["a", "b", "c"]
        assert info[a][2] == os.getpid() and info[c][2] == os.getpid()      # children of the driver
        assert info[b][2] == agent.pid                                        # child of the agent: the other "node"
        assert b._index // actors._POOL_SIZE == 1 and a._index // actors._POOL_SIZE == 0
        # handles travel between nodes and calls are routed both ways
        assert actors.get(a.set_peer.remote(b), timeout=30) and actors.get(b.set_peer.remote(c), timeout=30)
        assert actors.get(a.ask_peer.remote(), timeout=30) == "b"
        assert actors.get(b.ask_peer.remote(), timeout=30) == "c"
        assert len(actors.get(b.big.remote(3 << 20), timeout=60)) == 3 << 20  # a multi-megabyte reply over the socket
        # a remote actor that dies is reported, not waited for forever
        b.die.fire()
        with pytest.raises(actors.ActorDiedError):
            actors.get(b.whoami.remote(), timeout=30)
    finally:
        actors.shutdown()
        try:
            agent.wait(timeout=20)
        except subprocess.TimeoutExpired:
            agent.kill()
    assert agent.returncode == 0, agent.stdout.read()[-2000:]


def test_two_workers_on_two_nodes_train_in_sync(tmp_path):
    """`ray train -w 2 --address HOST:PORT --nodes 2`: one worker per node, gradients over torch.distributed (gloo)."""
    port = _free_port()
    agent = _agent(port)
    env = dict(os.environ, PYTHONPATH=f"{ROOT}{os.pathsep}" + os.environ.get("PYTHONPATH", ""))
    try:
        r = subprocess.run(
            [sys.executable, "-m", "spacy_ray_b200", "ray", "train", str(ROOT / "configs" / "tagger_w96.cfg"), "-w", "2",
             "--address", f"127.0.0.1:{port}", "--nodes", "2", "-o", str(tmp_path / "out"),
             "--training.max_steps", "20", "--training.eval_frequency", "10", "--corpora.train.n_docs", "300",
             "--corpora.dev.n_docs", "60"],
            env=env, cwd=str(ROOT), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        rows = [ln for ln in r.stdout.splitlines() if ln.strip() and ln.strip()[0].isdigit()]
        assert len(rows) >= 2, r.stdout
        assert (tmp_path / "out" / "model-best" / "config.cfg").exists()
        agent.wait(timeout=30)
    finally:
        if agent.poll() is None:
            agent.kill()
    assert agent.returncode == 0
