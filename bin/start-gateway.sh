#!/bin/bash
# Start the on-box gateway (the analogue of the reference's LiteLLM gateway launcher): one process,
# port from config/config.yaml unless overridden.  Extra flags are passed through (--stub, --spec tiny, ...).
cd "$(dirname "$0")/.." || exit 1
exec python rr_b200_server.py --config ./config/config.yaml "$@"
