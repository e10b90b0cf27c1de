"""Multi-node extension of the built-in actor runtime (``actors.py``).

The reference reaches other machines through Ray (``spacy ray train ... --address``,
``/root/reference/spacy_ray/train_cli.py:23,66-69``).  Here the DRIVER is the head: ``--address HOST:PORT
--nodes N`` makes it listen on HOST:PORT and wait for N-1 node agents,

    python -m spacy_ray_b200 ray node --address HOST:PORT [--num-gpus K]

one per additional machine.  Actors keep their mailbox semantics: an actor index encodes its node
(``index // POOL_SIZE``); inside a node messages travel through ``multiprocessing`` queues exactly as in the
single-node runtime, between nodes they are pickled over one TCP connection per agent and routed by the head
(a star: this is the CONTROL plane - create workers, ``set_proxy`` / ``train`` / ``is_running``, evaluation
scores, and the reference-faithful ``--mode async`` messages; the gradient exchange of ``--mode sync`` runs
over ``torch.distributed`` (NCCL / gloo) with the head as rendezvous - the peer-memory exchange is
single-node by construction).  Liveness: an agent reports the exit of any of its actors, a lost agent marks
all of its actors dead, so ``get`` raises ``ActorDiedError`` instead of hanging.
"""
from __future__ import annotations

import os
import pickle
import queue
import socket
import struct
import threading
import time
from typing import Any, Dict, List, Optional, Tuple


def _send(sock: socket.socket, lock: threading.Lock, obj: Any) -> None:
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    with lock:
        sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(n - len(buf), 1 << 20))
        if not chunk:
            raise ConnectionError("peer closed the connection")
        buf += chunk
    return bytes(buf)


def _recv(sock: socket.socket) -> Any:
    (n,) = struct.unpack("<Q", _recv_exact(sock, 8))
    return pickle.loads(_recv_exact(sock, n))


def parse_address(address: str) -> Tuple[str, int]:
    host, _, port = address.rpartition(":")
    if not host or not port.isdigit():
        raise ValueError(f"--address must be HOST:PORT, got {address!r}")
    return host, int(port)


class _RemoteSlot:
    """Stands in for the mailbox of an actor that lives on another node."""
    __slots__ = ("uplink", "index")

    def __init__(self, uplink, index: int):
        self.uplink = uplink
        self.index = index

    def put(self, msg) -> None:
        self.uplink.put((self.index, msg))


class Pool:
    """List-like view of every mailbox of the cluster from one node: local slots are this node's
    ``multiprocessing`` queues, the others forward through the node's uplink queue (picklable: it is handed to
    the spawned actor processes)."""

    def __init__(self, node_id: int, n_nodes: int, pool_size: int, local: List[Any], uplink: Any):
        self.node_id, self.n_nodes, self.pool_size = node_id, n_nodes, pool_size
        self.local, self.uplink = local, uplink

    def __len__(self) -> int:
        return self.n_nodes * self.pool_size

    def __getitem__(self, index: int):
        node, slot = divmod(int(index), self.pool_size)
        if node == self.node_id:
            return self.local[slot]
        if not 0 <= node < self.n_nodes:
            raise IndexError(index)
        return _RemoteSlot(self.uplink, int(index))


class RemoteProc:
    """``multiprocessing.Process``-like handle of an actor on another node (what ``_Runtime.wait`` polls)."""

    def __init__(self, head: "Head", index: int):
        self.head, self.index = head, index

    def is_alive(self) -> bool:
        return self.index not in self.head.dead

    @property
    def exitcode(self) -> Optional[int]:
        return self.head.dead.get(self.index)

    def join(self, timeout: Optional[float] = None) -> None:
        t0 = time.time()
        while self.is_alive() and (timeout is None or time.time() - t0 < timeout):
            time.sleep(0.02)

    def terminate(self) -> None:
        self.head.kill_remote(self.index)


class Head:
    """Driver side: listens, registers the agents, routes inter-node messages."""

    def __init__(self, address: str, n_nodes: int, pool_size: int, ctx, accept_timeout: float = 300.0):
        host, port = parse_address(address)
        self.n_nodes, self.pool_size = n_nodes, pool_size
        self.local = [ctx.Queue() for _ in range(pool_size)]
        self.uplink = ctx.Queue()
        self.pool = Pool(0, n_nodes, pool_size, self.local, self.uplink)
        self.dead: Dict[int, Optional[int]] = {}
        self.agents: Dict[int, Tuple[socket.socket, threading.Lock]] = {}
        self.node_info: Dict[int, Dict[str, Any]] = {0: {"gpus": None, "host": host}}
        self.actors_on: Dict[int, List[int]] = {n: [] for n in range(n_nodes)}
        self._gpu_cursor: Dict[int, int] = {n: 0 for n in range(n_nodes)}
        self._next_slot: Dict[int, int] = {n: (1 if n == 0 else 0) for n in range(n_nodes)}   # slot 0 of node 0 = driver
        self._closing = False
        self.host = host
        self._srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._srv.bind((host if host not in ("", "*") else "0.0.0.0", port))
        self._srv.listen(n_nodes)
        self._srv.settimeout(accept_timeout)
        for node_id in range(1, n_nodes):
            try:
                conn, _peer = self._srv.accept()
            except socket.timeout:
                raise TimeoutError(f"only {node_id - 1} of {n_nodes - 1} node agents connected to {address} within "
                                   f"{accept_timeout:.0f}s (start them with `python -m spacy_ray_b200 ray node --address {address}`)")
            conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            hello = _recv(conn)
            lock = threading.Lock()
            self.agents[node_id] = (conn, lock)
            self.node_info[node_id] = dict(hello[1])
            _send(conn, lock, ("welcome", node_id, n_nodes, pool_size))
            threading.Thread(target=self._reader, args=(node_id, conn), daemon=True, name=f"srb-head-rx-{node_id}").start()
        threading.Thread(target=self._uplink_loop, daemon=True, name="srb-head-uplink").start()

    # ---- routing ----------------------------------------------------------
    def route(self, index: int, msg) -> None:
        node, slot = divmod(index, self.pool_size)
        if node == 0:
            self.local[slot].put(msg)
            return
        agent = self.agents.get(node)
        if agent is None or index in self.dead:
            return                                   # receiver is gone: drop (its callers see ActorDiedError)
        try:
            _send(agent[0], agent[1], ("fwd", index, msg))
        except OSError:
            self._lost(node)

    def _uplink_loop(self) -> None:
        while not self._closing:
            try:
                index, msg = self.uplink.get(timeout=0.2)
            except queue.Empty:
                continue
            except (OSError, ValueError, EOFError):      # the queue was torn down under us (interpreter exit)
                return
            self.route(index, msg)

    def _reader(self, node_id: int, conn: socket.socket) -> None:
        try:
            while True:
                kind, *rest = _recv(conn)
                if kind == "fwd":
                    self.route(rest[0], rest[1])
                elif kind == "died":
                    self.dead[rest[0]] = rest[1]
        except (ConnectionError, OSError, EOFError):
            if not self._closing:
                self._lost(node_id)

    def _lost(self, node_id: int) -> None:
        for idx in self.actors_on.get(node_id, []):
            self.dead.setdefault(idx, -1)

    # ---- actor placement ----------------------------------------------------
    def place(self, num_gpus: int) -> int:
        """Node with the fewest actors so far (ties: lowest id) - fills the nodes evenly, the driver's node first."""
        return min(range(self.n_nodes), key=lambda n: (len(self.actors_on[n]), n))

    def new_index(self, node: int) -> int:
        slot = self._next_slot[node]
        if slot >= self.pool_size:
            raise RuntimeError(f"actor pool of node {node} exhausted ({self.pool_size} slots)")
        self._next_slot[node] = slot + 1
        index = node * self.pool_size + slot
        self.actors_on[node].append(index)
        return index

    def next_gpu(self, node: int) -> int:
        g = self._gpu_cursor[node]
        self._gpu_cursor[node] = g + 1
        return g

    def spawn_remote(self, node: int, index: int, payload: bytes, env: Dict[str, str], name: str) -> RemoteProc:
        conn, lock = self.agents[node]
        _send(conn, lock, ("spawn", index, payload, env, name))
        return RemoteProc(self, index)

    def kill_remote(self, index: int) -> None:
        node = index // self.pool_size
        agent = self.agents.get(node)
        if agent is not None:
            try:
                _send(agent[0], agent[1], ("kill", index))
            except OSError:
                pass

    def shutdown(self) -> None:
        self._closing = True
        for node_id, (conn, lock) in list(self.agents.items()):
            try:
                _send(conn, lock, ("stop",))
            except OSError:
                pass
        time.sleep(0.1)
        for conn, _lock in self.agents.values():
            try:
                conn.close()
            except OSError:
                pass
        try:
            self._srv.close()
        except OSError:
            pass


def agent_main(address: str, num_gpus: Optional[int] = None, connect_timeout: float = 300.0) -> int:
    """``python -m spacy_ray_b200 ray node --address HOST:PORT``: serve this machine's share of the actors."""
    from . import actors

    host, port = parse_address(address)
    deadline = time.time() + connect_timeout
    while True:
        try:
            sock = socket.create_connection((host, port), timeout=5.0)
            break
        except OSError:
            if time.time() > deadline:
                raise TimeoutError(f"no head at {address} within {connect_timeout:.0f}s")
            time.sleep(0.2)
    sock.settimeout(None)
    sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
    lock = threading.Lock()
    _send(sock, lock, ("hello", {"gpus": num_gpus, "host": socket.gethostname(), "pid": os.getpid()}))
    kind, node_id, n_nodes, pool_size = _recv(sock)
    assert kind == "welcome"
    ctx = actors._CTX
    local = [ctx.Queue() for _ in range(pool_size)]
    uplink = ctx.Queue()
    pool = Pool(node_id, n_nodes, pool_size, local, uplink)
    procs: Dict[int, Any] = {}
    reported: set = set()
    state = {"stop": False}

    def uplink_loop() -> None:
        while not state["stop"]:
            try:
                index, msg = uplink.get(timeout=0.2)
            except queue.Empty:
                continue
            except (OSError, ValueError, EOFError):
                return
            if index // pool_size == node_id:
                local[index % pool_size].put(msg)
            else:
                try:
                    _send(sock, lock, ("fwd", index, msg))
                except OSError:
                    state["stop"] = True

    def monitor() -> None:
        while not state["stop"]:
            for idx, p in list(procs.items()):
                if idx not in reported and not p.is_alive():
                    reported.add(idx)
                    try:
                        _send(sock, lock, ("died", idx, p.exitcode))
                    except OSError:
                        state["stop"] = True
            time.sleep(0.1)

    threading.Thread(target=uplink_loop, daemon=True, name="srb-agent-uplink").start()
    threading.Thread(target=monitor, daemon=True, name="srb-agent-monitor").start()
    print(f"[node {node_id}/{n_nodes}] connected to {address}", flush=True)
    try:
        while not state["stop"]:
            kind, *rest = _recv(sock)
            if kind == "fwd":
                local[rest[0] % pool_size].put(rest[1])
            elif kind == "spawn":
                index, payload, env, name = rest
                p = ctx.Process(target=actors._actor_main, args=(pool, index, payload, env), daemon=True, name=name)
                p.start()
                procs[index] = p
            elif kind == "kill":
                p = procs.get(rest[0])
                if p is not None and p.is_alive():
                    p.terminate()
            elif kind == "stop":
                break
    except (ConnectionError, OSError, EOFError):
        pass
    state["stop"] = True
    for idx, p in procs.items():
        if p.is_alive():
            try:
                local[idx % pool_size].put(("stop",))
            except Exception:
                pass
    for p in procs.values():
        p.join(timeout=5)
        if p.is_alive():
            p.terminate()
            p.join(timeout=2)
    try:
        sock.close()
    except OSError:
        pass
    return 0
