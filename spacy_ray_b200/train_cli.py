"""CLI + driver.

``python -m spacy_ray_b200 ray train CONFIG [-c code.py] [-o out] [-w N]
[-a addr] [-g gpu] [-V] [--section.key value ...]`` - the flags of
``spacy ray train`` (``/root/reference/spacy_ray/train_cli.py:23-36``),
including dotted config overrides from leftover args (``:44``) and loading the
config un-interpolated (``:46``).  Unlike the reference, ``--output`` is wired
(``:41`` is a TODO there).  Extra flags select the data plane:
``--mode sync|async``, ``--comm auto|dist|fused``, ``--quorum K``,
``--shard-balance auto|nodes|bytes|lpt``, ``--grad-transport fp32|bf16``, ``--resume PATH``.

``ray_train(config, *, ray_address, num_workers, use_gpu, code_path)`` keeps the
reference's Python signature (``:56-63``) and control flow (create workers ->
``set_proxy`` -> ``Evaluator`` -> ``train`` -> poll ``is_running``), running on
the built-in actor runtime instead of Ray.  Typer is not available here, so
parsing is argparse.
"""
from __future__ import annotations

import argparse
import logging
import socket
import sys
import time
from pathlib import Path
from typing import Any, Optional, Sequence

from .config import Config, ConfigValidationError, load_config, parse_config_overrides
from .utils.logging import logger
from .worker import Evaluator, Worker

RAY_HELP = (
    "CLI for parallel and distributed training (spacy-ray compatible surface; "
    "runs on the built-in single-node actor runtime + NCCL/NVLink kernels, not Ray)."
)


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return int(s.getsockname()[1])


def setup_gpu(use_gpu: int) -> None:
    """Driver-side device check (the reference calls spaCy's ``setup_gpu`` in the
    driver too, ``train_cli.py:43``); workers select their own device."""
    if use_gpu >= 0:
        import torch

        if not torch.cuda.is_available():
            raise SystemExit(f"--gpu-id {use_gpu} requested but CUDA is not available")
        logger.info("Using GPU: %d", use_gpu)
    else:
        logger.info("Using CPU")


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="spacy_ray_b200", description=RAY_HELP)
    sub = parser.add_subparsers(dest="group")
    ray = sub.add_parser("ray", help=RAY_HELP)
    ray_sub = ray.add_subparsers(dest="command")
    train = ray_sub.add_parser("train", help="Train a pipeline in parallel.")
    _add_train_args(train)
    node = ray_sub.add_parser("node", help="Join a multi-node run as a node agent (the driver's --address is the head).")
    node.add_argument("--address", "-a", required=True, help="HOST:PORT the driver listens on (its --address)")
    node.add_argument("--num-gpus", type=int, default=None, help="GPUs this node offers (informational)")
    plain = sub.add_parser("train", help="Single-process training (no workers).")
    _add_train_args(plain)
    ev = sub.add_parser("evaluate", help="Score a saved pipeline (model-best / model-last) on a corpus, like `spacy evaluate`.")
    ev.add_argument("model", type=Path, help="directory written by --output (model-best / model-last)")
    ev.add_argument("data_path", type=Path, help=".spacy (DocBin) or .jsonl corpus with gold annotations")
    ev.add_argument("--output", "-o", type=Path, default=None, help="write the scores as JSON")
    ev.add_argument("--gpu-id", "-g", dest="use_gpu", type=int, default=-1)
    ev.add_argument("--batch-size", type=int, default=256)
    conv = sub.add_parser("convert", help="Convert a JSONL / CoNLL-U / IOB corpus to a DocBin (.spacy) file, like `spacy convert`.")
    conv.add_argument("input_path", type=Path)
    conv.add_argument("output_path", type=Path, help="output file (.spacy) or directory")
    conv.add_argument("--limit", "-L", type=int, default=0, help="stop after this many documents")
    conv.add_argument("--converter", "-c", choices=["auto", "jsonl", "conllu", "iob"], default="auto")
    conv.add_argument("--n-sents", "-n", type=int, default=1, help="sentences per document (CoNLL-U / IOB)")
    conv.add_argument("--tag-column", choices=["xpos", "upos"], default="xpos", help="CoNLL-U column that becomes Doc.tags")
    conv.add_argument("--lang", "-l", default=None, help="accepted for `spacy convert` compatibility (unused)")
    return parser


def _add_train_args(p: argparse.ArgumentParser) -> None:
    p.add_argument("config_path", type=Path, help="Path to config file")
    p.add_argument("--code", "-c", dest="code_path", type=Path, default=None,
                   help="Path to Python file with additional code (registered functions) to be imported")
    p.add_argument("--output", "--output-path", "-o", dest="output_path", type=Path, default=None,
                   help="Output directory for the trained pipeline (model-best / model-last)")
    p.add_argument("--n-workers", "-w", dest="num_workers", type=int, default=1, help="Number of workers")
    p.add_argument("--address", "-a", dest="ray_address", default=None,
                   help="HOST:PORT this driver listens on as the head of a multi-node run (with --nodes > 1)")
    p.add_argument("--nodes", type=int, default=1,
                   help="machines in the run: the driver's plus N-1 started with `ray node --address HOST:PORT`")
    p.add_argument("--gpu-id", "-g", dest="use_gpu", type=int, default=-1, help="GPU ID or -1 for CPU")
    p.add_argument("--verbose", "-V", "-VV", dest="verbose", action="store_true", help="Debug logging")
    p.add_argument("--mode", choices=["sync", "async"], default="sync",
                   help="sync = flat-bucket reduce-scatter/Adam/all-gather per step; async = reference peer-proxy protocol")
    p.add_argument("--comm", choices=["auto", "dist", "fused", "local"], default="auto")
    p.add_argument("--quorum", type=int, default=None, help="async mode: gradients per update (reference default 2)")
    p.add_argument("--shard-balance", choices=["auto", "nodes", "bytes", "lpt"], default="auto")
    p.add_argument("--grad-transport", choices=["fp32", "bf16"], default="fp32",
                   help="gradient dtype on the wire for --comm dist (library collectives); optimizer state stays fp32")
    p.add_argument("--resume", dest="resume_path", type=Path, default=None, help="Checkpoint dir to resume from")
    p.add_argument("--no-shard-data", dest="shard_data", action="store_false",
                   help="Give every rank the full corpus (reference behaviour)")
    p.add_argument("--inject-fault", default=None, help="rank:step - raise in that worker at that step (tests)")


def main(argv: Optional[Sequence[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    parser = build_parser()
    args, extra = parser.parse_known_args(argv)
    if args.group is None or (args.group == "ray" and args.command is None):
        parser.print_help()
        return 1
    if args.group == "ray" and args.command == "node":
        from .parallel.cluster import agent_main

        return agent_main(args.address, args.num_gpus)
    if args.group == "evaluate":
        return evaluate_cli(args)
    if args.group == "convert":
        from .training.docbin import convert

        out = args.output_path
        if out.suffix != ".spacy":
            out.mkdir(parents=True, exist_ok=True)
            out = out / (args.input_path.stem + ".spacy")
        n = convert(args.input_path, out, converter=args.converter, n_sents=args.n_sents, limit=args.limit,
                    tag_column=args.tag_column)
        print(f"✔ Generated output file ({n} documents): {out}")
        return 0
    try:
        ray_train_cli(args, extra)
    except ConfigValidationError as e:
        print(f"\n✘ Config validation error ({args.config_path})\n{e}", file=sys.stderr)
        return 1
    return 0


def evaluate_cli(args: argparse.Namespace) -> int:
    """``spacy evaluate``: load a saved pipeline, annotate the corpus, print the components' scores."""
    import json

    from .pipeline import load
    from .training.corpus import JsonlCorpus

    if not args.model.exists():
        print(f"\u2718 Model directory not found: {args.model}", file=sys.stderr)
        return 1
    if not args.data_path.exists():
        print(f"\u2718 Evaluation data not found: {args.data_path}", file=sys.stderr)
        return 1
    setup_gpu(args.use_gpu)
    nlp = load(args.model)
    examples = list(JsonlCorpus(str(args.data_path))(nlp))
    scores = nlp.evaluate(examples, batch_size=args.batch_size)
    flat = {k: v for k, v in scores.items() if isinstance(v, (int, float)) and v is not None}
    width = max((len(k) for k in flat), default=5)
    print(f"\n================== Results ({len(examples)} docs) ==================\n")
    for k, v in flat.items():
        shown = f"{v:.0f}" if k == "speed" else f"{100 * v:.2f}"
        print(f"{k.upper():<{width}}   {shown}")
    if args.output is not None:
        args.output.parent.mkdir(parents=True, exist_ok=True)
        args.output.write_text(json.dumps(scores, indent=2, default=str))
        print(f"\n\u2714 Saved results to {args.output}")
    return 0


def ray_train_cli(args: argparse.Namespace, extra: Sequence[str]) -> None:
    logger.setLevel(logging.DEBUG if args.verbose else logging.ERROR)
    if not args.config_path.exists():
        raise ConfigValidationError(f"Config file not found: {args.config_path}")
    setup_gpu(args.use_gpu)
    overrides = parse_config_overrides(extra)
    config = load_config(args.config_path, overrides=overrides, interpolate=False)
    ray_train(
        config,
        ray_address=args.ray_address,
        num_workers=args.num_workers,
        use_gpu=args.use_gpu,
        code_path=args.code_path,
        output_path=args.output_path,
        mode=args.mode,
        comm=args.comm,
        quorum=args.quorum,
        shard_balance=args.shard_balance,
        grad_transport=args.grad_transport,
        nodes=getattr(args, "nodes", 1),
        resume_path=args.resume_path,
        shard_data=args.shard_data,
        inject_fault=args.inject_fault,
    )


def ray_train(
    config: Config,
    *,
    ray_address: Optional[str] = None,
    num_workers: int = 1,
    use_gpu: int = -1,
    code_path: Optional[Path] = None,
    output_path: Optional[Path] = None,
    mode: str = "sync",
    comm: str = "auto",
    quorum: Optional[int] = None,
    shard_balance: str = "auto",
    grad_transport: str = "fp32",
    nodes: int = 1,
    resume_path: Optional[Path] = None,
    shard_data: bool = True,
    inject_fault: Optional[str] = None,
    ray: Any = None,
    poll_interval: float = 0.2,
) -> None:
    """Driver: same sequence as the reference (``train_cli.py:66-91``)."""
    if ray is None:
        from .parallel import actors as ray
    master_addr = "127.0.0.1"
    if ray_address is not None:
        if nodes > 1:
            # this driver is the head: node agents connect to ray_address, torch.distributed meets there too;
            # the peer-memory exchange cannot span machines, so "auto" means the library collectives
            ray.init(address=ray_address, nodes=nodes)
            host = ray_address.rpartition(":")[0]
            if host in ("", "*", "0.0.0.0"):
                import socket

                host = socket.gethostbyname(socket.gethostname())
            master_addr = host
            if comm == "auto":
                comm = "dist"
            elif comm == "fused":
                raise ValueError("--comm fused is single-node (NVLink peer memory); use --comm dist across machines")
        else:
            ray.init(address=ray_address)
    else:
        ray.init(ignore_reinit_error=True)
    dist_init = {"master_addr": master_addr, "master_port": _free_port()}
    RemoteWorker = ray.remote(Worker).options(num_gpus=int(use_gpu >= 0), num_cpus=2)
    workers = [
        RemoteWorker.remote(
            config,
            rank=rank,
            num_workers=num_workers,
            use_gpu=use_gpu,
            code_path=code_path,
            mode=mode,
            comm=comm,
            quorum=quorum,
            output_path=output_path,
            resume_path=resume_path,
            shard_data=shard_data,
            shard_balance=shard_balance,
            grad_transport=grad_transport,
            dist_init=dist_init,
            inject_fault=inject_fault,
        )
        for rank in range(num_workers)
    ]
    try:
        # set_proxy may run a collective (initial weight sync), so launch it on all
        # workers before waiting on any (the reference waits serially, which is
        # fine for its purely local set_proxy).
        ray.get([w.set_proxy.remote(workers) for w in workers])
        evaluator = ray.remote(Evaluator).remote()
        for worker in workers:
            ray.get(worker.train.remote(workers, evaluator))
        todo = list(workers)
        while todo:
            time.sleep(poll_interval)
            todo = [w for w in workers if ray.get(w.is_running.remote())]
            errors = [e for e in ray.get([w.get_error.remote() for w in workers]) if e]
            if errors:
                raise RuntimeError("worker failed:\n" + "\n".join(errors))
        return ray.get([w.get_stats.remote() for w in workers])     # per-rank throughput / message counters
    finally:
        ray.shutdown()


def train_single(config: Config, *, use_gpu: int = -1, output_path: Optional[Path] = None,
                 code_path: Optional[Path] = None) -> Worker:
    """Single-process training through the same ``Worker`` (world size 1)."""
    worker = Worker(config, rank=0, num_workers=1, use_gpu=use_gpu, code_path=code_path,
                    mode="sync", comm="local", output_path=output_path)
    worker.set_proxy(None)
    worker.train(None, None)
    worker.join()
    return worker


# Name the reference registers with spaCy's CLI (``train_cli.py:19``).
ray_cli = main
