"""The parts of ``bench.py``'s driver contract that can be checked without a GPU: the reference arm really
tries to import the reference and reports the exception text; the argument surface the driver uses exists;
the flagship config is the model BASELINE.json names."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(*args):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_real_reason():
    out = _run("--impl", "reference", "--gpus", "1", "--steps", "5", "--warmup", "3")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 1
    if "unavailable" in line:                     # this image: spaCy / thinc / Ray are not installable
        why = line["unavailable"]
        assert "baseline/_ref" in why
        assert "does not exist" in why or "ModuleNotFoundError" in why or "ImportError" in why or "could not be driven" in why
    else:                                         # an image with the wheels: a real measurement
        assert line["value"] > 0 and line["unit"]


def test_cli_surface_the_driver_uses():
    out = _run("--help")
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--config", "--no-own-baseline"):
        assert flag in out.stdout, flag


def test_flagship_config_is_the_named_model():
    sys.path.insert(0, str(ROOT))
    import argparse

    import bench
    from spacy_ray_b200.config import Config

    ns = argparse.Namespace(docs_per_gpu=1024, width=256, depth=8, min_len=8, max_len=40, dropout=0.1)
    cfg = Config().from_str(bench.flagship_config(ns, 0), interpolate=False)
    assert cfg["nlp"]["pipeline"] == ["ner"] or "ner" in cfg["nlp"]["pipeline"]
    t2v = cfg["components"]["ner"]["model"]["tok2vec"]
    assert "MultiHashEmbed" in t2v["embed"]["@architectures"] and t2v["embed"]["width"] == 256
    assert "MaxoutWindowEncoder" in t2v["encode"]["@architectures"]
    assert t2v["encode"]["width"] == 256 and t2v["encode"]["depth"] == 8 and t2v["encode"]["maxout_pieces"] == 3
    base = json.loads((ROOT / "BASELINE.json").read_text())
    assert "tok2vec+NER" in base["metric"]
