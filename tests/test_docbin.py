"""DocBin (.spacy) container: hash compatibility with spaCy's string store, round trips, the
``spacy.Corpus.v1`` reader and the ``convert`` CLI (reference: bin/get-data.sh:6-13)."""
import json

import numpy as np
import pytest

from spacy_ray_b200.pipeline.doc import Doc
from spacy_ray_b200.training.docbin import ATTR_IDS, DocBin, convert_jsonl, hash_string, murmurhash64a


def test_hash_string_matches_spacys_documented_values():
    # values from spaCy's documentation of the StringStore
    assert hash_string("coffee") == 3197928453018144401
    assert hash_string("apple") == 8566208034543834098
    assert hash_string("") == 0
    assert murmurhash64a(b"12345678abcdefgh", 1) == murmurhash64a(b"12345678abcdefgh", 1)
    assert hash_string("ab") != hash_string("ba") and hash_string("12345678") != hash_string("123456789")


def _docs():
    return [
        Doc(["Apple", "is", "looking", "at", "U.K.", "startups"], [True, True, True, True, True, False],
            tags=["PROPN", "AUX", "VERB", "ADP", "PROPN", "NOUN"],
            ents=[(0, 1, "ORG"), (4, 5, "GPE")],
            heads=[2, 2, 2, 2, 5, 3], deps=["nsubj", "aux", "ROOT", "prep", "compound", "pobj"]),
        Doc(["Hello", "world", "!"], ents=[]),
        Doc(["New", "York", "City", "and", "Rome"], ents=[(0, 3, "GPE"), (4, 5, "GPE")]),
        Doc(["no", "annotation"]),
    ]


def test_docbin_roundtrip_bytes_and_disk(tmp_path):
    db = DocBin(docs=_docs())
    assert db.attrs[0] == ATTR_IDS["ORTH"] and db.attrs[1:] == sorted(db.attrs[1:])
    data = db.to_bytes()
    for back in (DocBin().from_bytes(data), None):
        if back is None:
            db.to_disk(tmp_path / "x.spacy")
            back = DocBin().from_disk(tmp_path / "x.spacy")
        got = list(back.get_docs())
        assert len(got) == 4
        a = got[0]
        assert a.words == ["Apple", "is", "looking", "at", "U.K.", "startups"] and a.spaces[-1] is False
        assert a.tags == ["PROPN", "AUX", "VERB", "ADP", "PROPN", "NOUN"]
        assert a.heads == [2, 2, 2, 2, 5, 3] and a.deps[2] == "ROOT"
        assert a.ents == [(0, 1, "ORG"), (4, 5, "GPE")] and a.has_ents_annotation
        assert got[1].ents == [] and got[1].has_ents_annotation and got[1].heads is None and got[1].tags is None
        assert got[2].ents == [(0, 3, "GPE"), (4, 5, "GPE")]
        assert not got[3].has_ents_annotation
    # negative head offsets are stored as two's complement in the uint64 cells
    col = db.attrs.index(ATTR_IDS["HEAD"])
    assert int(db.tokens[0][5, col]) == (1 << 64) - 2


def test_docbin_layout_is_the_documented_msgpack_map():
    import msgpack
    import zlib

    msg = msgpack.unpackb(zlib.decompress(DocBin(docs=_docs()).to_bytes()), raw=False)
    assert {"version", "attrs", "tokens", "spaces", "lengths", "strings", "cats", "flags"} <= set(msg)
    lengths = np.frombuffer(msg["lengths"], dtype="<i4")
    assert lengths.tolist() == [6, 3, 5, 2]
    assert len(msg["tokens"]) == int(lengths.sum()) * len(msg["attrs"]) * 8
    assert "Apple" in msg["strings"] and "GPE" in msg["strings"] and msg["strings"] == sorted(msg["strings"])


def test_corpus_reader_reads_spacy_files_and_convert_cli(tmp_path):
    from spacy_ray_b200.train_cli import main
    from spacy_ray_b200.training.corpus import create_docbin_reader

    src = tmp_path / "train.jsonl"
    with src.open("w", encoding="utf8") as f:
        for d in _docs():
            f.write(json.dumps(d.to_dict()) + "\n")
        f.write(json.dumps({"text": "Zara opened in Paris", "spans": [{"start": 0, "end": 4, "label": "BRAND"}]}) + "\n")
    assert main(["convert", str(src), str(tmp_path / "out")]) == 0
    out = tmp_path / "out" / "train.spacy"
    assert out.exists()
    egs = list(create_docbin_reader(str(out))(None))
    assert len(egs) == 5 and egs[4].reference.ents == [(0, 1, "BRAND")]
    assert egs[0].reference.heads == [2, 2, 2, 2, 5, 3]
    assert egs[0].predicted.tags is None, "the predicted side must not carry gold"
    # a directory with both kinds of file
    assert len(list(create_docbin_reader(str(tmp_path))(None))) == 10
    assert len(list(create_docbin_reader(str(out), limit=2)(None))) == 2
    assert convert_jsonl(src, tmp_path / "two.spacy", limit=2) == 2
    with pytest.raises(ValueError):
        DocBin().from_bytes(b"not a docbin")
