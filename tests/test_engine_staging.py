"""Host side of the device-resident engine: the packed staging buffer must carry exactly the
batch / gold arrays the generic ``nlp.update`` path would build (engine/trainer.py)."""
import numpy as np
import torch

from conftest import multi_cfg
from spacy_ray_b200.config import Config
from spacy_ray_b200.engine.trainer import ExampleStore, _kind, _Layout, _Views, fill_stage, make_stage
from spacy_ray_b200.training.initialize import init_nlp


def _nlp_and_examples():
    cfg = Config().from_str(multi_cfg(["tagger", "parser", "ner"], width=32, depth=1, n_docs=40, max_len=9, hidden=32),
                            interpolate=False)
    nlp = init_nlp(cfg)
    from spacy_ray_b200.config import registry, resolve_dot_names

    corpus = resolve_dot_names(cfg.interpolate(), ["corpora.train"])[0]
    return nlp, list(corpus(nlp))


def test_packed_buffer_matches_generic_batch_and_gold():
    nlp, examples = _nlp_and_examples()
    heads = [(n, c, _kind(c)) for n, c in nlp.pipeline if getattr(c, "is_trainable", False)]
    assert [k for _, _, k in heads] == ["tagger", "parser", "ner"]
    store = ExampleStore(examples, heads)
    B = 8
    lay = _Layout(rows=256, docs=B, lmax=64, slots=tuple(store.slots))
    stage = make_stage(lay, store, pin=False)
    for ids in (np.array([3, 5, 6, 10, 11, 20, 21, 39]), np.array([0, 1, 2])):   # full, then partial batch
        fill_stage(store, lay, stage, ids)
        chosen = [examples[i] for i in ids]
        batch = nlp.make_batch([eg.predicted for eg in chosen])
        rows = stage["rows"]
        assert rows == batch.n_rows
        a = stage["np"]
        np.testing.assert_array_equal(a["attrs"][:rows].view(np.uint64), batch.attrs.numpy().view(np.uint64))
        np.testing.assert_array_equal(a["mask"][:rows], batch.mask.numpy().reshape(-1))
        assert a["mask"][rows:].sum() == 0
        np.testing.assert_array_equal(a["starts"][: len(ids)], batch.doc_starts.numpy())
        np.testing.assert_array_equal(a["lens"][: len(ids)], batch.doc_lens.numpy())
        assert (a["lens"][len(ids):] == 0).all()
        tagger, parser, ner = nlp.get_pipe("tagger"), nlp.get_pipe("parser"), nlp.get_pipe("ner")
        want_tags = tagger._gold_labels(chosen, batch).numpy()
        np.testing.assert_array_equal(a["gold"]["tagger"][:rows], want_tags)
        assert (a["gold"]["tagger"][rows:] == -1).all()
        words = stage["words"]
        np.testing.assert_array_equal(a["gold"]["ner"][:words],
                                      np.concatenate([ner.gold_actions(eg.reference) for eg in chosen]))
        np.testing.assert_array_equal(a["gold"]["parser.heads"][:words],
                                      np.concatenate([parser._gold_np(eg.reference)[0] for eg in chosen]))
        np.testing.assert_array_equal(a["gold"]["parser.labels"][:words],
                                      np.concatenate([parser._gold_np(eg.reference)[1] for eg in chosen]))
        # HashEmbed-backward grouping: per attribute column, every token row exactly once with
        # equal ids adjacent; the tail points at pad row 0
        rb = min(lay.rows, -(-rows // lay.bucket) * lay.bucket)
        token_rows = np.nonzero(a["mask"][:rows])[0]
        for c in range(4):
            perm = a["perm"][c * rb: (c + 1) * rb]
            assert sorted(perm[:words].tolist()) == token_rows.tolist()
            assert (perm[words:] == 0).all()
            col_ids = a["attrs"][perm[:words], c]
            n_runs = int((col_ids[1:] != col_ids[:-1]).sum()) + 1
            assert n_runs == len(np.unique(col_ids))
        lens = np.array([len(eg) for eg in chosen])
        np.testing.assert_array_equal(a["tok_off"][: len(ids)], np.cumsum(lens) - lens)
    # typed views of the same bytes on the "device" side
    dv = _Views(lay, stage["buf"].clone())
    assert dv.attrs.shape == (256, 4) and dv.gold["tagger"].dtype == torch.int32


def test_unsupported_reason_lists_the_blocking_component():
    from spacy_ray_b200.engine import Trainer

    nlp, examples = _nlp_and_examples()
    assert Trainer.unsupported_reason(nlp, max_len=9) is None
    assert "128" in Trainer.unsupported_reason(nlp, max_len=500)
