"""Test infrastructure only: CPU restatements used as the parity checker (never shipped,
never on the product path).  See oracle/router.py and oracle/llama_ref.py headers."""
