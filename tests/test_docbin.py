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


CONLLU = """# sent_id = 1
# text = The cat sat.
1\tThe\tthe\tDET\tDT\t_\t2\tdet\t_\t_
2\tcat\tcat\tNOUN\tNN\t_\t3\tnsubj\t_\t_
3\tsat\tsit\tVERB\tVBD\t_\t0\troot\t_\tSpaceAfter=No
4\t.\t.\tPUNCT\t.\t_\t3\tpunct\t_\t_

# sent_id = 2
1-2\tdon't\t_\t_\t_\t_\t_\t_\t_\t_
1\tdo\tdo\tAUX\tVBP\t_\t3\taux\t_\t_
2\tn't\tnot\tPART\tRB\t_\t3\tadvmod\t_\t_
3\tgo\tgo\tVERB\tVB\t_\t0\troot\t_\tNER=O
3.1\tghost\t_\t_\t_\t_\t_\t_\t_\t_
4\tParis\tParis\tPROPN\tNNP\t_\t3\tobl\t_\tNER=B-GPE
"""


def test_convert_conllu_reads_tags_heads_deps_sentences_and_misc_ner(tmp_path):
    from spacy_ray_b200.training.docbin import DocBin, convert, read_conllu

    src = tmp_path / "tb.conllu"
    src.write_text(CONLLU, encoding="utf8")
    docs = read_conllu(src, n_sents=1)
    assert [d.words for d in docs] == [["The", "cat", "sat", "."], ["do", "n't", "go", "Paris"]]
    assert docs[0].tags == ["DT", "NN", "VBD", "."] and docs[0].heads == [1, 2, 2, 2]
    assert docs[0].deps == ["det", "nsubj", "ROOT", "punct"] and docs[0].spaces == [True, True, False, True]
    assert docs[1].ents == [(3, 4, "GPE")] and not docs[0].has_ents_annotation
    both = read_conllu(src, n_sents=2, tag_column="upos")
    assert len(both) == 1 and both[0].heads == [1, 2, 2, 2, 6, 6, 6, 6] and both[0].tags[:2] == ["DET", "NOUN"]
    n = convert(src, tmp_path / "tb.spacy", n_sents=2)
    assert n == 1
    back = list(DocBin().from_disk(tmp_path / "tb.spacy").get_docs())
    assert back[0].words == both[0].words and back[0].heads == both[0].heads and back[0].deps[2] == "ROOT"


def test_convert_iob_formats(tmp_path):
    from spacy_ray_b200.training.docbin import _biluo_or_iob_to_spans, convert, read_iob

    assert _biluo_or_iob_to_spans(["O", "B-PER", "I-PER", "O", "I-LOC", "B-LOC", "U-ORG", "B-X", "L-X"]) == [
        (1, 3, "PER"), (4, 5, "LOC"), (5, 6, "LOC"), (6, 7, "ORG"), (7, 9, "X")]
    conll = tmp_path / "ner.conll"
    conll.write_text("-DOCSTART- -X- O O\n\nJohn NNP B-NP B-PER\nSmith NNP I-NP I-PER\nsleeps VBZ B-VP O\n\nParis NNP B-NP B-LOC\n",
                     encoding="utf8")
    docs = read_iob(conll)
    assert [d.words for d in docs] == [["John", "Smith", "sleeps"], ["Paris"]]
    assert docs[0].ents == [(0, 2, "PER")] and docs[0].tags == ["NNP", "NNP", "VBZ"] and docs[1].ents == [(0, 1, "LOC")]
    iob = tmp_path / "ner.iob"
    iob.write_text("I|PRP|O like|VBP|O London|NNP|B-GPE .|.|O\nBerlin|B-GPE rocks|O\n", encoding="utf8")
    docs = read_iob(iob, n_sents=2)
    assert len(docs) == 1 and docs[0].words[2] == "London" and docs[0].ents == [(2, 3, "GPE"), (4, 5, "GPE")]
    assert docs[0].tags is None                       # the second sentence has no tag column
    assert convert(iob, tmp_path / "ner.spacy") == 2
    with pytest.raises(ValueError):
        convert(tmp_path / "x.unknown", tmp_path / "x.spacy")
