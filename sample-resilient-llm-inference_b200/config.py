"""Loader for the reference's gateway configuration schema.

Keeps the keys the reference's config/config.yaml uses (reference config/config.yaml:31-33 `litellm.port`,
:35-94 `model_list[].{model_name, litellm_params.model, rpm, tpm}`, :100-108
`router_settings.{routing_strategy, enable_pre_call_checks, allowed_fails, cooldown_time, fallbacks}`);
YAML anchors / merge keys (`<<: *aws-defaults`, :40) are resolved by PyYAML.  AWS-only keys
(`aws_region_name`, `aws_profile_name`, the `aws:` / `cris:` sections) are accepted and ignored.
Local extension: `litellm_params.gpu: N` pins a deployment to a replica; `litellm_params.model:
b200/<spec>` names an on-box model (models.SPECS) instead of `bedrock/<id>`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import yaml

STRATEGIES = {"simple-shuffle": 0, "least-busy": 1, "round-robin": 2, "split": 3, "random": 4}
_PROVIDER_PREFIXES = ("bedrock/", "b200/", "openai/", "azure/")


def strip_provider(model: str) -> str:
    """`bedrock/us.anthropic...` -> `us.anthropic...` — what callers see in `response.model`
    (reference README.md:153-162; demos do the same for display, src/demo_load_balancing.py:59-60)."""
    for p in _PROVIDER_PREFIXES:
        if model.startswith(p):
            return model[len(p):]
    return model


@dataclass
class DeploymentCfg:
    index: int
    model_name: str            # model group
    model: str                 # litellm_params.model
    group: int
    rpm: int = -1
    tpm: int = -1
    weight: int = -1           # simple-shuffle weight (litellm_params.weight / rpm / tpm), -1 = unset
    gpu: int = 0
    params: Dict[str, Any] = field(default_factory=dict)

    @property
    def response_model(self) -> str:
        return strip_provider(self.model)


@dataclass
class RouterConfig:
    groups: List[str]
    deployments: List[DeploymentCfg]
    fallbacks: Dict[int, List[int]]
    routing_strategy: str = "simple-shuffle"
    enable_pre_call_checks: bool = False
    allowed_fails: int = 3
    cooldown_time: float = 5.0
    port: int = 4000

    @property
    def strategy_id(self) -> int:
        return STRATEGIES[self.routing_strategy]

    def group_index(self, name: str) -> int:
        try:
            return self.groups.index(name)
        except ValueError:
            return -1


def _int_or(v, default=-1) -> int:
    if v is None:
        return default
    if isinstance(v, bool) or not isinstance(v, (int, float)) or int(v) != v:
        raise ValueError(f"rpm/tpm/weight must be integers, got {v!r}")
    return int(v)


def build_config(model_list: List[dict], router_settings: Optional[dict] = None, port: int = 4000) -> RouterConfig:
    rs = dict(router_settings or {})
    groups: List[str] = []
    deps: List[DeploymentCfg] = []
    for i, entry in enumerate(model_list or []):
        name = entry["model_name"]
        lp = dict(entry.get("litellm_params") or {})
        if "model" not in lp:
            raise ValueError(f"model_list[{i}] has no litellm_params.model")
        if name not in groups:
            groups.append(name)
        # limits: top level of the entry (reference config.yaml:41-42) or inside litellm_params
        rpm = _int_or(entry.get("rpm", lp.get("rpm")))
        tpm = _int_or(entry.get("tpm", lp.get("tpm")))
        deps.append(DeploymentCfg(index=i, model_name=name, model=lp["model"], group=groups.index(name),
                                  rpm=rpm, tpm=tpm, gpu=int(lp.get("gpu", 0)), params=lp))
    # simple-shuffle weights: the first of weight / rpm / tpm present in the litellm_params of the
    # group's first deployment decides the key for the whole group (oracle/router.py header).
    for g in range(len(groups)):
        members = [d for d in deps if d.group == g]
        key = next((k for k in ("weight", "rpm", "tpm") if members[0].params.get(k) is not None), None)
        for d in members:
            d.weight = _int_or(d.params.get(key), 0) if key else -1
    strategy = rs.get("routing_strategy", "simple-shuffle")
    if strategy not in STRATEGIES:
        raise ValueError(f"unsupported routing_strategy {strategy!r}; supported: {sorted(STRATEGIES)}")
    fbs: Dict[int, List[int]] = {}
    for item in rs.get("fallbacks") or []:
        for src, dst in item.items():
            if src not in groups:
                raise ValueError(f"fallbacks: unknown model group {src!r}")
            lst = []
            for d in dst:
                if d not in groups:
                    raise ValueError(f"fallbacks: unknown model group {d!r}")
                lst.append(groups.index(d))
            fbs[groups.index(src)] = lst
    return RouterConfig(groups=groups, deployments=deps, fallbacks=fbs, routing_strategy=strategy,
                        enable_pre_call_checks=bool(rs.get("enable_pre_call_checks", False)),
                        allowed_fails=int(rs.get("allowed_fails", 3)),
                        cooldown_time=float(rs.get("cooldown_time", 5.0)), port=port)


def load_config(path: str) -> RouterConfig:
    with open(path, "r") as f:
        raw = yaml.safe_load(f) or {}
    port = (raw.get("litellm") or {}).get("port", 4000)
    if not isinstance(port, int) or port < 1024 or port > 65535:   # reference bin/start-gateway.sh:18
        port = 4000
    return build_config(raw.get("model_list") or [], raw.get("router_settings") or {}, port)
