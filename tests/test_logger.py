import io

from spacy_ray_b200.loggers import ray_console_logger


class FakeNLP:
    pipe_names = ["tagger", "ner"]
    config = {"training": {"score_weights": {"tag_acc": 0.5, "ents_f": 0.5}}}


def test_console_logger_matches_reference_table_layout():
    out = io.StringIO()
    log_step, finalize = ray_console_logger(stream=out)(FakeNLP())
    log_step({"seconds": 65, "epoch": 1, "step": 200, "words": 12345, "score": 0.87,
              "losses": {"tagger": 12.3456, "ner": 0.5}, "other_scores": {"tag_acc": 0.9, "ents_f": 0.8}})
    finalize()
    lines = out.getvalue().splitlines()
    assert lines[0].split() == ["T", "E", "#", "W", "LOSS", "TAGGER", "LOSS", "NER", "TAG_ACC", "ENTS_F", "SCORE"]
    widths = [8, 3, 6, 6, 11, 8, 7, 6, 6]
    assert lines[1] == "   ".join("-" * w for w in widths)
    cells = lines[2].split()
    assert cells == ["0:01:05", "1", "200", "12345", "12.35", "0.50", "90.00", "80.00", "0.87"]
    assert len(lines[2]) == len(lines[1])


def test_missing_loss_key_raises_keyerror():
    import pytest

    log_step, _ = ray_console_logger(stream=io.StringIO())(FakeNLP())
    with pytest.raises(KeyError):
        log_step({"seconds": 1, "epoch": 0, "step": 0, "words": 1, "score": 0.0, "losses": {"tagger": 1.0},
                  "other_scores": {}})
