import subprocess
import sys
from pathlib import Path

import pytest

from spacy_ray_b200.train_cli import build_parser

ROOT = Path(__file__).resolve().parent.parent


def test_cli_flags_match_reference_surface():
    p = build_parser()
    args, extra = p.parse_known_args(
        ["ray", "train", "cfg.cfg", "-c", "code.py", "-o", "out", "-w", "4", "-a", "auto", "-g", "0", "-V",
         "--training.max_steps", "5"])
    assert args.group == "ray" and args.command == "train"
    assert str(args.config_path) == "cfg.cfg" and str(args.code_path) == "code.py" and str(args.output_path) == "out"
    assert args.num_workers == 4 and args.ray_address == "auto" and args.use_gpu == 0 and args.verbose
    assert extra == ["--training.max_steps", "5"]
    args2, _ = p.parse_known_args(["ray", "train", "cfg.cfg", "--output-path", "o2", "--n-workers", "2", "--gpu-id", "-1"])
    assert str(args2.output_path) == "o2" and args2.num_workers == 2 and args2.use_gpu == -1


@pytest.mark.slow
def test_cli_end_to_end_single_worker(tmp_path):
    cfg = ROOT / "configs" / "tagger_w96.cfg"
    r = subprocess.run(
        [sys.executable, "-m", "spacy_ray_b200", "ray", "train", str(cfg), "-w", "1", "-o", str(tmp_path),
         "--training.max_steps", "3", "--training.eval_frequency", "2", "--corpora.train.n_docs", "60"],
        cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "LOSS TAGGER" in r.stdout and (tmp_path / "model-last" / "config.cfg").exists()


def test_missing_config_is_a_clean_error(tmp_path):
    r = subprocess.run([sys.executable, "-m", "spacy_ray_b200", "ray", "train", str(tmp_path / "nope.cfg")],
                       cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Config" in r.stderr


@pytest.mark.slow
def test_cli_trains_ner_from_jsonl_files(tmp_path):
    """The reference's example workload shape (bin/get-data.sh -> NER from JSONL): make-data.py writes the
    files, `spacy.Corpus.v1` reads them through ${paths.*} overrides, checkpoints land in -o."""
    data = tmp_path / "data"
    r = subprocess.run([sys.executable, str(ROOT / "bin" / "make-data.py"), str(data), "--n-train", "80", "--n-dev", "20"],
                       cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    out = tmp_path / "out"
    r = subprocess.run(
        [sys.executable, "-m", "spacy_ray_b200", "ray", "train", str(ROOT / "configs" / "ner_jsonl.cfg"), "-w", "1",
         "-o", str(out), "--paths.train", str(data / "train.jsonl"), "--paths.dev", str(data / "dev.jsonl"),
         "--training.max_steps", "4", "--training.eval_frequency", "2"],
        cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "LOSS NER" in r.stdout and "ENTS_F" in r.stdout
    assert (out / "model-last" / "ner" / "model").exists()


@pytest.mark.slow
def test_stock_spacy_generated_config_runs_unchanged(tmp_path):
    """A config as `spacy init config` writes it (tokenizer block, [initialize], augmenter = null, listener +
    shared tok2vec, batch_by_words with a compounding size, spacy.ConsoleLogger.v1) must train as is."""
    data = tmp_path / "data"
    r = subprocess.run([sys.executable, str(ROOT / "bin" / "make-data.py"), str(data), "--n-train", "60", "--n-dev", "15"],
                       cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run(
        [sys.executable, "-m", "spacy_ray_b200", "ray", "train", str(ROOT / "configs" / "spacy_default_ner.cfg"), "-w", "2",
         "--paths.train", str(data / "train.jsonl"), "--paths.dev", str(data / "dev.jsonl"),
         "--training.max_steps", "4", "--training.eval_frequency", "2"],
        cwd=ROOT, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "LOSS TOK2VEC" in r.stdout and "LOSS NER" in r.stdout and "ENTS_F" in r.stdout


def test_evaluate_subcommand_scores_a_saved_pipeline(tmp_path, capsys):
    import json

    from spacy_ray_b200.config import Config, resolve_dot_names
    from spacy_ray_b200.train_cli import main
    from spacy_ray_b200.training import init_nlp
    from spacy_ray_b200.training.docbin import DocBin

    cfg = ROOT / "configs" / "tagger_w96.cfg"
    out = tmp_path / "out"
    small = ["--training.max_steps", "6", "--training.eval_frequency", "3", "--corpora.train.n_docs", "60",
             "--corpora.dev.n_docs", "20"]
    assert main(["ray", "train", str(cfg), "-w", "1", "-o", str(out), *small]) == 0
    nlp = init_nlp(Config().from_disk(cfg, overrides={"corpora.train.n_docs": 60, "corpora.dev.n_docs": 20}))
    (dev,) = resolve_dot_names(nlp.config.interpolate(), ["corpora.dev"])
    DocBin(docs=[eg.reference for eg in dev(nlp)]).to_disk(tmp_path / "dev.spacy")
    capsys.readouterr()
    assert main(["evaluate", str(out / "model-best"), str(tmp_path / "dev.spacy"), "-o", str(tmp_path / "s.json")]) == 0
    shown = capsys.readouterr().out
    assert "TAG_ACC" in shown and "SPEED" in shown
    scores = json.loads((tmp_path / "s.json").read_text())
    assert 0.0 <= scores["tag_acc"] <= 1.0
    assert main(["evaluate", str(tmp_path / "nope"), str(tmp_path / "dev.spacy")]) == 1
