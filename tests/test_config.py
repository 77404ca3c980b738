"""CPU: the config loader keeps the reference's YAML schema (reference config/config.yaml:31-108)."""
import textwrap

import pytest

from rr_b200.config import build_config, load_config, strip_provider
from rr_b200.models import resolve_spec

REF_STYLE = textwrap.dedent('''
    aws:
      profile_name: &aws-profile "genai"
      region_name: &aws-region "us-east-1"
    x-aws-defaults: &aws-defaults
      aws_region_name: *aws-region
      aws_profile_name: *aws-profile
    litellm:
      port: 4100
    model_list:
    - model_name: demo
      litellm_params:
        model: bedrock/us.anthropic.claude-sonnet-4-20250514-v1:0
        <<: *aws-defaults
      rpm: 3
      tpm: 100000
    - model_name: demo
      litellm_params:
        model: bedrock/us.anthropic.claude-3-7-sonnet-20250219-v1:0
        <<: *aws-defaults
      rpm: 3
    - model_name: overflow
      litellm_params:
        model: bedrock/us.anthropic.claude-3-5-sonnet-20241022-v2:0
      rpm: 25
    cris:
      model_id: "x"
    router_settings:
      routing_strategy: "simple-shuffle"
      enable_pre_call_checks: true
      allowed_fails: 2
      cooldown_time: 15
      fallbacks: [
        {"demo": ["overflow"]}
      ]
''')


def test_reference_style_yaml_with_anchors_and_merge_keys(tmp_path):
    p = tmp_path / "c.yaml"
    p.write_text(REF_STYLE)
    cfg = load_config(str(p))
    assert cfg.port == 4100 and cfg.groups == ["demo", "overflow"]
    assert [(d.group, d.rpm, d.tpm, d.weight) for d in cfg.deployments] == [(0, 3, 100000, -1), (0, 3, -1, -1), (1, 25, -1, -1)]
    assert cfg.deployments[0].params["aws_region_name"] == "us-east-1"       # merge key resolved, then ignored
    assert cfg.fallbacks == {0: [1]} and cfg.routing_strategy == "simple-shuffle"
    assert cfg.enable_pre_call_checks and cfg.allowed_fails == 2 and cfg.cooldown_time == 15
    assert cfg.deployments[0].response_model == "us.anthropic.claude-sonnet-4-20250514-v1:0"


def test_repo_config_loads_and_resolves_specs():
    cfg = load_config("config/config.yaml")
    assert len(cfg.deployments) == 8 and len(cfg.groups) == 7
    assert cfg.group_index("claude-sonnet-loadbalance-demo") == 1 and cfg.group_index("nope") == -1
    assert {d.gpu for d in cfg.deployments} == {0, 1}
    assert all(resolve_spec(d.model).name == "llama-3-8b" for d in cfg.deployments)
    assert "claude-3-5-sonnet" in cfg.deployments[3].response_model       # what the demos grep for


def test_weights_from_litellm_params_and_validation():
    ml = [{"model_name": "g", "litellm_params": {"model": "b200/tiny", "weight": 3}},
          {"model_name": "g", "litellm_params": {"model": "b200/tiny", "weight": 1}},
          {"model_name": "h", "litellm_params": {"model": "b200/tiny", "rpm": 7}}]
    cfg = build_config(ml, {"routing_strategy": "least-busy"})
    assert [d.weight for d in cfg.deployments] == [3, 1, 7] and cfg.deployments[2].rpm == 7
    assert cfg.strategy_id == 1
    with pytest.raises(ValueError):
        build_config(ml, {"routing_strategy": "usage-based-routing"})
    with pytest.raises(ValueError):
        build_config(ml, {"fallbacks": [{"g": ["missing"]}]})
    with pytest.raises(ValueError):
        build_config([{"model_name": "g", "litellm_params": {"model": "b200/tiny"}, "rpm": 2.5}])
    assert strip_provider("b200/llama-3-8b@x") == "llama-3-8b@x" and strip_provider("plain") == "plain"
