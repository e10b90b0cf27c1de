import pytest

from spacy_ray_b200.config import Config, ConfigValidationError, parse_config_overrides, registry, resolve_dot_names


CFG = """
[paths]
train = "x.jsonl"
dev = null

[a]
x = 1
y = [1, 2, 3]
z = ${paths.train}
name = "v-${a.x}"

[a.b]
@schedules = "compounding.v1"
start = 1.0
stop = 8.0
compound = 2.0

[c]
ref = ${a.y}
"""


def test_parse_types_and_nesting():
    c = Config().from_str(CFG, interpolate=False)
    assert c["a"]["x"] == 1 and c["a"]["y"] == [1, 2, 3]
    assert c["paths"]["dev"] is None
    assert c["a"]["b"]["@schedules"] == "compounding.v1"
    assert c["a"]["z"] == "${paths.train}" and not c.is_interpolated


def test_interpolation_values_sections_and_strings():
    c = Config().from_str(CFG, interpolate=True)
    assert c["a"]["z"] == "x.jsonl"
    assert c["a"]["name"] == "v-1"
    assert c["c"]["ref"] == [1, 2, 3]


def test_roundtrip_text():
    c = Config().from_str(CFG, interpolate=False)
    again = Config().from_str(c.to_str(), interpolate=False)
    assert again == c


def test_overrides_cli_style():
    ov = parse_config_overrides(["--a.x", "5", "--paths.train=y.jsonl", "--a.flag"])
    assert ov == {"a.x": 5, "paths.train": "y.jsonl", "a.flag": True}
    c = Config().from_str(CFG, interpolate=True, overrides=ov)
    assert c["a"]["x"] == 5 and c["a"]["z"] == "y.jsonl" and c["a"]["flag"] is True
    with pytest.raises(ConfigValidationError):
        parse_config_overrides(["--nodots", "1"])
    with pytest.raises(ConfigValidationError):
        parse_config_overrides(["positional"])


def test_registry_resolution_and_errors():
    c = Config().from_str(CFG, interpolate=True)
    resolved = registry.resolve(c["a"])
    sched = resolved["b"]
    assert [next(sched) for _ in range(5)] == [1.0, 2.0, 4.0, 8.0, 8.0]
    with pytest.raises(ConfigValidationError):
        registry.resolve({"q": {"@schedules": "nope.v1"}})
    with pytest.raises(ConfigValidationError):
        registry.resolve({"q": {"@schedules": "compounding.v1", "start": 1.0}})   # missing args


def test_resolve_dot_names():
    c = Config({"corpora": {"train": {"@readers": "spacy_ray_b200.SyntheticCorpus.v1", "n_docs": 3}}})
    (train, none) = resolve_dot_names(c, ["corpora.train", None])
    assert none is None and len(list(train())) == 3


def test_merge_switching_registry_function_drops_old_args():
    base = Config({"o": {"@optimizers": "Adam.v1", "learn_rate": 0.1, "beta1": 0.8}})
    merged = base.merge({"o": {"@optimizers": "SGD.v1", "learn_rate": 0.5}})
    assert merged["o"] == {"@optimizers": "SGD.v1", "learn_rate": 0.5}
