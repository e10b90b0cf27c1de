#!/usr/bin/env python
"""Write a synthetic NER/tagging/parsing corpus as JSONL (the reference's bin/get-data.sh
downloads the 'fashion brands' NER set; there is no network here, so we generate).

    python bin/make-data.py out_dir --n-train 20000 --n-dev 1000
"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from spacy_ray_b200.training.corpus import SyntheticCorpus  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("out_dir", type=Path)
ap.add_argument("--n-train", type=int, default=20000)
ap.add_argument("--n-dev", type=int, default=1000)
ap.add_argument("--n-ent-labels", type=int, default=18)
args = ap.parse_args()
args.out_dir.mkdir(parents=True, exist_ok=True)
for name, n, seed in (("train", args.n_train, 1), ("dev", args.n_dev, 2)):
    corpus = SyntheticCorpus(n, seed=seed, n_ent_labels=args.n_ent_labels)
    with (args.out_dir / f"{name}.jsonl").open("w", encoding="utf8") as f:
        for doc in corpus.docs():
            f.write(json.dumps(doc.to_dict()) + "\n")
    print(f"wrote {n} docs to {args.out_dir / (name + '.jsonl')}")
