#!/usr/bin/env bash
# Example workload of the reference (bin/get-data.sh there downloads the "fashion brands" NER
# JSONL files and converts them).  This box has no network, so the default is a synthetic corpus
# of the same shape; pass URLs to fetch real JSONL instead:
#   bin/get-data.sh [TRAIN_URL DEV_URL]
set -euo pipefail
out=${OUT_DIR:-data}
mkdir -p "$out"
# works from any directory, installed or not: put the repository root on the module path
repo="$(cd "$(dirname "$0")/.." && pwd)"
export PYTHONPATH="$repo${PYTHONPATH:+:$PYTHONPATH}"
if [ $# -ge 2 ] && command -v wget >/dev/null; then
  wget -q -O "$out/train.jsonl" "$1"
  wget -q -O "$out/dev.jsonl" "$2"
else
  python "$(dirname "$0")/make-data.py" "$out" --n-train ${TRAIN_DOCS:-20000} --n-dev ${DEV_DOCS:-2000}
fi
# the reference converts with `spacy convert`; same step, same output names (DocBin .spacy files)
python -m spacy_ray_b200 convert "$out/train.jsonl" "$out/train.spacy"
python -m spacy_ray_b200 convert "$out/dev.jsonl" "$out/dev.spacy"
echo "wrote $out/{train,dev}.jsonl and $out/{train,dev}.spacy (use with [corpora.*] @readers = \"spacy.Corpus.v1\")"
