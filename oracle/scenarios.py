"""Reference topology (reference config/config.yaml:35-108) expressed for the oracle and for the
parity tests — TEST INFRASTRUCTURE.  Numbers and names are the reference's configuration data:
model groups, per-deployment rpm/tpm (top level of each model_list entry, so simple-shuffle has
no weights and picks uniformly), router_settings and the two single-hop fallback chains."""
from __future__ import annotations

from . import router as O

_S4 = "bedrock/us.anthropic.claude-sonnet-4-20250514-v1:0"
_S37 = "bedrock/us.anthropic.claude-3-7-sonnet-20250219-v1:0"
_S35 = "bedrock/us.anthropic.claude-3-5-sonnet-20241022-v2:0"

# (model_name, litellm_params.model, rpm, tpm)            reference config.yaml line
REF_MODEL_LIST = [
    ("claude-sonnet-fallback-demo", _S4, 3, 100000),         # :36-42
    ("claude-sonnet-loadbalance-demo", _S4, 3, 100000),      # :45-50
    ("claude-sonnet-loadbalance-demo", _S37, 3, 100000),     # :52-57
    ("claude-sonnet-fallback-loadbalance", _S35, 25, 250000),  # :60-65
    ("claude-sonnet-fallback-quota", _S35, 25, 250000),      # :67-72
    ("consumer-a-model", _S37, 3, 30000),                    # :75-80
    ("consumer-b-model", _S37, 10, 100000),                  # :82-87
    ("consumer-c-model", _S37, 10, 100000),                  # :89-94
]
REF_FALLBACKS = {                                            # :105-108
    "claude-sonnet-fallback-demo": ["claude-sonnet-fallback-quota"],
    "claude-sonnet-loadbalance-demo": ["claude-sonnet-fallback-loadbalance"],
}
REF_SETTINGS = dict(strategy=O.STRATEGY_SIMPLE_SHUFFLE,      # :101
                    enable_pre_call_checks=True,             # :102
                    allowed_fails=2,                         # :103
                    cooldown_ms=15000)                       # :104

REF_GROUP_NAMES = []
for _n, *_ in REF_MODEL_LIST:
    if _n not in REF_GROUP_NAMES:
        REF_GROUP_NAMES.append(_n)
REF_GROUPS = {n: i for i, n in enumerate(REF_GROUP_NAMES)}


def reference_topology():
    """-> (deployments [O.Deployment], n_groups, fallbacks {g: [g]}, settings, info [dict])."""
    deps = [O.Deployment(group=REF_GROUPS[n], rpm=rpm, tpm=tpm, weight=-1, replica=i)
            for i, (n, _m, rpm, tpm) in enumerate(REF_MODEL_LIST)]
    info = [dict(model_name=n, model=m, rpm=rpm, tpm=tpm) for n, m, rpm, tpm in REF_MODEL_LIST]
    fbs = {REF_GROUPS[k]: [REF_GROUPS[x] for x in v] for k, v in REF_FALLBACKS.items()}
    return deps, len(REF_GROUP_NAMES), fbs, O.Settings(**REF_SETTINGS), info


def reference_router(seed: int = 0):
    deps, ng, fbs, st, info = reference_topology()
    return O.OracleRouter(deps, ng, fbs, st, seed=seed), info
