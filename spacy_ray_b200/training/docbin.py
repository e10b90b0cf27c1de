"""spaCy ``DocBin`` (``.spacy``) reader / writer without spaCy.

The reference's only example workload converts JSONL to DocBin with ``spacy convert`` and trains
from it through ``spacy.Corpus.v1`` (``/root/reference/bin/get-data.sh:6-13``).  spaCy is not
installable here, so the container format is implemented directly (written from knowledge of
spaCy 3.x ``spacy/tokens/_serialize.py``; not byte-compared against a real file on this box):

    zlib( msgpack( {"version", "attrs": [attr ids, ORTH first], "tokens": uint64[n_tokens, n_attrs],
                    "spaces": bool[n_tokens, 1], "lengths": int32[n_docs], "strings": [...],
                    "cats": [...], "flags": [{"has_unknown_spaces": bool}], ...} ) )

String-valued attributes (ORTH, TAG, DEP, ENT_TYPE, ...) hold the 64-bit MurmurHash64A (seed 1) of
the string - spaCy's ``hash_string`` - and are resolved through the ``strings`` list; ``HEAD`` is
the head offset relative to the token (two's complement in the uint64 cell); ``ENT_IOB`` is
0 = missing, 1 = I, 2 = O, 3 = B; ``SENT_START`` is 1 / -1 / 0 (start / not a start / unknown).

``hash_string`` is also exported for code that needs spaCy-compatible ids (vectors tables keyed
by ORTH hash, ``nn.staticvectors``).
"""
from __future__ import annotations

import zlib
from pathlib import Path
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np

from ..pipeline.doc import Doc

# spacy/attrs.pxd ``attr_id_t`` (the ids are part of the file format)
ATTR_IDS: Dict[str, int] = {
    "ORTH": 65, "LOWER": 66, "NORM": 67, "SHAPE": 68, "PREFIX": 69, "SUFFIX": 70, "LENGTH": 71, "CLUSTER": 72,
    "LEMMA": 73, "POS": 74, "TAG": 75, "DEP": 76, "ENT_IOB": 77, "ENT_TYPE": 78, "HEAD": 79, "SENT_START": 80,
    "SPACY": 81, "PROB": 82, "LANG": 83, "ENT_KB_ID": 84, "MORPH": 85, "ENT_ID": 86, "IDX": 87,
}
ATTR_NAMES = {v: k for k, v in ATTR_IDS.items()}
DEFAULT_ATTRS = ("ORTH", "TAG", "HEAD", "DEP", "ENT_IOB", "ENT_TYPE", "ENT_KB_ID", "LEMMA", "MORPH", "POS", "SENT_START")

_M64 = (1 << 64) - 1
_MM = 0xC6A4A7935BD1E995


def murmurhash64a(data: bytes, seed: int = 1) -> int:
    """Austin Appleby's MurmurHash64A (the 64-bit variant for 64-bit platforms)."""
    n = len(data)
    h = (seed ^ ((n * _MM) & _M64)) & _M64
    nblocks = n // 8
    if nblocks:
        for k in np.frombuffer(data, dtype="<u8", count=nblocks).tolist():
            k = (k * _MM) & _M64
            k ^= k >> 47
            k = (k * _MM) & _M64
            h ^= k
            h = (h * _MM) & _M64
    tail = data[nblocks * 8:]
    if tail:
        h ^= int.from_bytes(tail, "little")
        h = (h * _MM) & _M64
    h ^= h >> 47
    h = (h * _MM) & _M64
    h ^= h >> 47
    return h


def hash_string(s: str) -> int:
    """spaCy's ``strings.hash_string``: MurmurHash64A of the UTF-8 bytes, seed 1; "" -> 0."""
    return murmurhash64a(s.encode("utf8"), 1) if s else 0


def _u64(v: int) -> int:
    return v & _M64


def _i64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


class DocBin:
    """Container of serialised docs (subset of spaCy's API: ``add``, ``get_docs``,
    ``to_bytes`` / ``from_bytes``, ``to_disk`` / ``from_disk``, ``merge``, ``len``)."""

    def __init__(self, attrs: Iterable[str] = DEFAULT_ATTRS, store_user_data: bool = False,
                 docs: Iterable[Doc] = ()):
        ids = [ATTR_IDS[a] if isinstance(a, str) else int(a) for a in attrs]
        ids = sorted(a for a in set(ids) if a not in (ATTR_IDS["ORTH"], ATTR_IDS["SPACY"]))
        self.version = "0.1"
        self.attrs: List[int] = [ATTR_IDS["ORTH"]] + ids            # ORTH is always column 0
        self.tokens: List[np.ndarray] = []
        self.spaces: List[np.ndarray] = []
        self.cats: List[dict] = []
        self.span_groups: List[bytes] = []
        self.flags: List[dict] = []
        self.strings: set = set()
        self.store_user_data = store_user_data
        for doc in docs:
            self.add(doc)

    def __len__(self) -> int:
        return len(self.tokens)

    # ---- writing ---------------------------------------------------------------
    def add(self, doc: Doc) -> None:
        n = len(doc)
        col = {a: i for i, a in enumerate(self.attrs)}
        arr = np.zeros((n, len(self.attrs)), dtype=np.uint64)

        def put(name: str, values: Sequence[int]) -> None:
            j = col.get(ATTR_IDS[name])
            if j is not None:
                arr[:, j] = np.asarray([_u64(int(v)) for v in values], dtype=np.uint64)

        def strs(name: str, values: Sequence[Optional[str]]) -> None:
            for v in values:
                if v:
                    self.strings.add(v)
            put(name, [hash_string(v) if v else 0 for v in values])

        strs("ORTH", doc.words)
        if doc.tags is not None:
            strs("TAG", doc.tags)
        if doc.pos is not None:
            strs("POS", doc.pos)
        if doc.morphs is not None:
            strs("MORPH", doc.morphs)
        if doc.lemmas is not None:
            strs("LEMMA", doc.lemmas)
        if doc.heads is not None:
            put("HEAD", [(h - i) if (h is not None and h >= 0) else 0 for i, h in enumerate(doc.heads)])
            strs("DEP", doc.deps if doc.deps is not None else [None] * n)
        if doc.sent_starts is not None:
            put("SENT_START", [0 if v is None else (1 if v else -1) for v in doc.sent_starts])
        elif doc.heads is not None:
            # one sentence per doc unless the tree has several roots
            roots = [i for i, h in enumerate(doc.heads) if h == i]
            starts = set(_sentence_starts(doc.heads)) if len(roots) > 1 else {0}
            put("SENT_START", [1 if i in starts else -1 for i in range(n)])
        if doc.has_ents_annotation:
            iob = [2] * n
            etype: List[Optional[str]] = [None] * n
            for s, e, lab in doc.ents:
                for t in range(s, e):
                    iob[t] = 3 if t == s else 1
                    etype[t] = lab
            put("ENT_IOB", iob)
            strs("ENT_TYPE", etype)
        self.tokens.append(arr)
        self.spaces.append(np.asarray(doc.spaces, dtype=bool).reshape(n, 1))
        self.cats.append({str(k): float(v) for k, v in (doc.cats or {}).items()})
        self.span_groups.append(_pack_spans(doc.spans))
        self.flags.append({"has_unknown_spaces": False})

    def to_bytes(self) -> bytes:
        import msgpack

        lengths = np.asarray([len(t) for t in self.tokens], dtype="int32")
        tokens = np.vstack(self.tokens) if self.tokens else np.zeros((0, len(self.attrs)), dtype=np.uint64)
        spaces = np.vstack(self.spaces) if self.spaces else np.zeros((0, 1), dtype=bool)
        msg = {
            "version": self.version, "attrs": list(self.attrs),
            "tokens": tokens.astype("<u8").tobytes("C"), "spaces": spaces.tobytes("C"),
            "lengths": lengths.astype("<i4").tobytes("C"), "strings": sorted(self.strings),
            "cats": self.cats, "flags": self.flags,
            "span_groups": list(self.span_groups) + [b""] * (len(self.tokens) - len(self.span_groups)),
        }
        return zlib.compress(msgpack.packb(msg, use_bin_type=True))

    def to_disk(self, path: Union[str, Path]) -> None:
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_bytes(self.to_bytes())

    # ---- reading ---------------------------------------------------------------
    def from_bytes(self, data: bytes) -> "DocBin":
        import msgpack

        try:
            msg = msgpack.unpackb(zlib.decompress(data), raw=False, strict_map_key=False)
        except (zlib.error, ValueError) as e:
            raise ValueError(f"not a DocBin (.spacy) file: {e}") from e
        self.version = msg.get("version", "0.1")
        self.attrs = [int(a) for a in msg["attrs"]]
        self.strings = set(msg["strings"])
        lengths = np.frombuffer(msg["lengths"], dtype="<i4")
        n_attr = len(self.attrs)
        flat = np.frombuffer(msg["tokens"], dtype="<u8")
        total = int(lengths.sum())
        if flat.size != total * n_attr:
            raise ValueError(f"DocBin: token array has {flat.size} cells, expected {total} x {n_attr}")
        flat = flat.reshape(total, n_attr)
        spaces = np.frombuffer(msg["spaces"], dtype=bool)
        spaces = spaces.reshape(total, 1) if spaces.size == total else np.ones((total, 1), dtype=bool)
        self.tokens, self.spaces = [], []
        pos = 0
        for n in lengths.tolist():
            self.tokens.append(flat[pos:pos + n])
            self.spaces.append(spaces[pos:pos + n])
            pos += n
        self.cats = list(msg.get("cats") or [{} for _ in self.tokens])
        self.span_groups = list(msg.get("span_groups") or [b"" for _ in self.tokens])
        self.flags = list(msg.get("flags") or [{} for _ in self.tokens])
        return self

    def from_disk(self, path: Union[str, Path]) -> "DocBin":
        return self.from_bytes(Path(path).read_bytes())

    def merge(self, other: "DocBin") -> None:
        if self.attrs != other.attrs:
            raise ValueError("DocBin.merge: attribute lists differ")
        self.tokens.extend(other.tokens)
        self.spaces.extend(other.spaces)
        self.cats.extend(other.cats)
        self.span_groups.extend(other.span_groups)
        self.flags.extend(other.flags)
        self.strings.update(other.strings)

    def get_docs(self, vocab=None) -> Iterator[Doc]:
        lookup = {hash_string(s): s for s in self.strings}
        col = {a: i for i, a in enumerate(self.attrs)}

        def column(arr, name):
            j = col.get(ATTR_IDS[name])
            return None if j is None else arr[:, j].tolist()

        def strings_of(values, what):
            out = []
            for v in values:
                if v == 0:
                    out.append(None)
                else:
                    s = lookup.get(v)
                    if s is None:
                        raise ValueError(f"DocBin: {what} hash {v} is not in the string table")
                    out.append(s)
            return out

        cats_list = list(self.cats) + [{}] * (len(self.tokens) - len(self.cats))
        groups_list = list(self.span_groups) + [b""] * (len(self.tokens) - len(self.span_groups))
        for arr, sp, cats, groups in zip(self.tokens, self.spaces, cats_list, groups_list):
            n = len(arr)
            words = strings_of(column(arr, "ORTH"), "ORTH")
            if any(w is None for w in words):
                raise ValueError("DocBin: token without ORTH")
            tags = column(arr, "TAG")
            tag_s = strings_of(tags, "TAG") if tags is not None and any(tags) else None
            heads = deps = None
            hd = column(arr, "HEAD")
            dp = column(arr, "DEP")
            if hd is not None and dp is not None and any(dp):
                heads = [i + _i64(h) for i, h in enumerate(hd)]
                heads = [h if 0 <= h < n else i for i, h in enumerate(heads)]
                deps = [d or "dep" for d in strings_of(dp, "DEP")]
            ents = None
            iob = column(arr, "ENT_IOB")
            if iob is not None and any(iob):
                et = strings_of(column(arr, "ENT_TYPE") or [0] * n, "ENT_TYPE")
                ents, start = [], None
                for t in range(n + 1):
                    tag = iob[t] if t < n else 2
                    if start is not None and (tag != 1 or et[t] != et[start]):
                        ents.append((start, t, et[start] or ""))
                        start = None
                    if t < n and tag == 3 or (t < n and tag == 1 and start is None and et[t]):
                        start = t
            def opt_strings(name):
                vals = column(arr, name)
                return strings_of(vals, name) if vals is not None and any(vals) else None

            ss = column(arr, "SENT_START")
            sent_starts = None
            if ss is not None and any(ss):
                sent_starts = [None if _i64(v) == 0 else (_i64(v) > 0) for v in ss]
            yield Doc(words, [bool(x) for x in sp.reshape(-1).tolist()], tags=tag_s, ents=ents, heads=heads, deps=deps,
                      pos=opt_strings("POS"), morphs=opt_strings("MORPH"), lemmas=opt_strings("LEMMA"),
                      sent_starts=sent_starts, cats=cats or None, spans=_unpack_spans(groups))


_SPAN_MAGIC = b"srb-spans-1"


def _pack_spans(spans: Optional[Dict]) -> bytes:
    """Span groups of one doc in DocBin's per-doc ``span_groups`` bytes slot.  (spaCy stores its own
    ``SpanGroup.to_bytes`` blobs there, which need a Vocab to decode; ours carry a magic prefix + msgpack
    ``{name: [[start, end, label], ...]}`` and anything else is ignored on read.)"""
    if not spans:
        return b""
    import msgpack

    return _SPAN_MAGIC + msgpack.packb({str(k): [[int(a), int(b), str(l)] for a, b, l in v] for k, v in spans.items()},
                                       use_bin_type=True)


def _unpack_spans(blob) -> Optional[Dict]:
    if not blob or not isinstance(blob, (bytes, bytearray)) or not bytes(blob).startswith(_SPAN_MAGIC):
        return None
    import msgpack

    return {k: [tuple(x) for x in v] for k, v in msgpack.unpackb(bytes(blob)[len(_SPAN_MAGIC):], raw=False).items()}


def _sentence_starts(heads: Sequence[int]) -> List[int]:
    """First token of every root's subtree span (projective trees: spans are contiguous)."""
    n = len(heads)
    root_of = list(range(n))
    for i in range(n):
        j, guard = i, 0
        while heads[j] != j and 0 <= heads[j] < n and guard <= n:
            j, guard = heads[j], guard + 1
        root_of[i] = j
    starts, seen = [], set()
    for i in range(n):
        if root_of[i] not in seen:
            seen.add(root_of[i])
            starts.append(i)
    return starts


def convert_jsonl(src: Union[str, Path], dst: Union[str, Path], *, limit: int = 0) -> int:
    """``spacy convert``-style: JSONL (one doc per line, this package's or Prodigy's format) -> DocBin."""
    import json

    db = DocBin()
    with Path(src).open("r", encoding="utf8") as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            doc = Doc.from_dict(json.loads(line))
            if len(doc):
                db.add(doc)
            if limit and len(db) >= limit:
                break
    db.to_disk(dst)
    return len(db)


# ----------------------------------------------------------------------------------------------
# `spacy convert` for the treebank / NER column formats (what real tagger / parser / NER corpora
# ship as): CoNLL-U and IOB / CoNLL-2003.  Output: Docs with words, spaces, tags, heads, deps, ents.
# ----------------------------------------------------------------------------------------------
def _biluo_or_iob_to_spans(tags: Sequence[str]) -> List[tuple]:
    """Entity spans ``(start, end_exclusive, label)`` from IOB / IOB2 / BILUO token tags."""
    spans, start, label = [], None, None

    def close(i):
        nonlocal start, label
        if start is not None:
            spans.append((start, i, label))
        start, label = None, None

    for i, t in enumerate(tags):
        t = t or "O"
        if t in ("O", "-", "_", ""):
            close(i)
            continue
        prefix, _, lab = t.partition("-")
        if not lab:                      # bare label without a prefix: treat as I-
            prefix, lab = "I", t
        if prefix in ("B", "U") or (prefix in ("I", "L") and (start is None or lab != label)):
            close(i)
            start, label = i, lab
        if prefix in ("L", "U"):
            close(i + 1)
    close(len(tags))
    return spans


def read_conllu(src: Union[str, Path], *, n_sents: int = 1, tag_column: str = "xpos", merge_subtokens: bool = False):
    """Docs from a CoNLL-U treebank (ID FORM LEMMA UPOS XPOS FEATS HEAD DEPREL DEPS MISC).  ``n_sents``
    sentences are concatenated per Doc (heads re-based, every sentence root heads itself - the parser's
    sentence boundaries come from that); multi-word-token ranges (``3-4``) and empty nodes (``5.1``) are
    skipped (``merge_subtokens`` is accepted for CLI compatibility and not implemented); NER tags in MISC
    (``NER=B-ORG``/``name=B-ORG``, as `spacy convert` reads them) become ``ents``."""
    col = {"upos": 3, "xpos": 4}[tag_column]
    sents, cur = [], []
    with Path(src).open("r", encoding="utf8") as fh:
        for line in list(fh) + [""]:
            line = line.rstrip("\n")
            if not line.strip():
                if cur:
                    sents.append(cur)
                    cur = []
                continue
            if line.startswith("#"):
                continue
            parts = line.split("\t")
            if len(parts) < 8 or "-" in parts[0] or "." in parts[0]:
                continue
            cur.append(parts)
    docs = []
    for i in range(0, len(sents), max(1, n_sents)):
        words, spaces, tags, heads, deps, ner = [], [], [], [], [], []
        pos, morphs, lemmas, starts = [], [], [], []
        any_ner = False
        for sent in sents[i:i + max(1, n_sents)]:
            base = len(words)
            for parts in sent:
                pos.append(parts[3] if parts[3] != "_" else None)
                morphs.append(parts[5] if parts[5] != "_" else "")
                lemmas.append(parts[2] if parts[2] != "_" else None)
                starts.append(len(words) == base)
                tid = int(parts[0]) - 1
                head = int(parts[6]) if parts[6].isdigit() else 0
                misc = parts[9] if len(parts) > 9 else "_"
                words.append(parts[1])
                spaces.append("SpaceAfter=No" not in misc)
                tag = parts[col] if parts[col] != "_" else parts[3]
                tags.append(tag)
                heads.append(base + (tid if head == 0 else head - 1))
                deps.append("ROOT" if head == 0 else parts[7])
                t = "O"
                for kv in misc.split("|"):
                    k, _, v = kv.partition("=")
                    if k in ("NER", "name") and v:
                        t, any_ner = v, True
                ner.append(t)
        if words:
            docs.append(Doc(words, spaces, tags=tags, heads=heads, deps=deps,
                            ents=_biluo_or_iob_to_spans(ner) if any_ner else None,
                            pos=pos if any(pos) else None, morphs=morphs if any(pos) else None,
                            lemmas=lemmas if any(lemmas) else None, sent_starts=starts))
    return docs


def read_iob(src: Union[str, Path], *, n_sents: int = 1):
    """Docs from the NER column formats `spacy convert` takes: CoNLL-2003 style (one token per line,
    whitespace-separated columns, word first, NER tag last, POS tag second if there are >= 3 columns,
    blank line = sentence, ``-DOCSTART-`` = document) or spaCy's ``.iob`` (one sentence per line,
    ``word|tag|ner`` or ``word|ner`` tokens)."""
    sents, cur = [], []
    with Path(src).open("r", encoding="utf8") as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    pipe_format = any("|" in ln and len(ln.split()) > 1 and all("|" in t for t in ln.split()) for ln in lines if ln.strip())
    if pipe_format:
        for ln in lines:
            toks = ln.split()
            if not toks:
                continue
            rows = []
            for t in toks:
                f = t.split("|")
                rows.append((f[0], f[1] if len(f) >= 3 else None, f[-1]))
            sents.append(rows)
    else:
        for ln in lines + [""]:
            if not ln.strip() or ln.startswith("-DOCSTART-"):
                if cur:
                    sents.append(cur)
                    cur = []
                continue
            f = ln.split()
            cur.append((f[0], f[1] if len(f) >= 3 else None, f[-1]))
    docs = []
    for i in range(0, len(sents), max(1, n_sents)):
        words, tags, ner = [], [], []
        for sent in sents[i:i + max(1, n_sents)]:
            for w, tag, t in sent:
                words.append(w)
                tags.append(tag)
                ner.append(t)
        if words:
            docs.append(Doc(words, tags=tags if all(t is not None for t in tags) else None,
                            ents=_biluo_or_iob_to_spans(ner)))
    return docs


def convert(src: Union[str, Path], dst: Union[str, Path], *, converter: str = "auto", n_sents: int = 1,
            limit: int = 0, tag_column: str = "xpos") -> int:
    """``spacy convert``: ``.jsonl`` / ``.conllu`` (``.conll``) / ``.iob`` (``.ner``) -> DocBin."""
    src = Path(src)
    kind = converter
    if kind == "auto":
        ext = src.suffix.lower().lstrip(".")
        kind = {"jsonl": "jsonl", "json": "jsonl", "conllu": "conllu", "conll": "conllu", "iob": "iob", "ner": "iob"}.get(ext)
        if kind is None:
            raise ValueError(f"cannot guess the converter for {src.name!r}; pass --converter jsonl|conllu|iob")
    if kind == "jsonl":
        return convert_jsonl(src, dst, limit=limit)
    docs = read_conllu(src, n_sents=n_sents, tag_column=tag_column) if kind == "conllu" else read_iob(src, n_sents=n_sents)
    db = DocBin()
    for d in docs[: limit or None]:
        db.add(d)
    db.to_disk(dst)
    return len(db)

