"""CPU checks of the C-ABI boundary: the library loads without a GPU, exports every function
declared in include/rr_b200.h, the ctypes table matches the header, and host-only entry points
(MT19937 seeding, tokenizer) give reference answers."""
import ctypes as C
import random
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "rr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from rr_b200 import _lib
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(_lib.lib, n)]
    assert not missing, missing
    assert not _lib.MISSING
    assert sorted(_lib.PROTOTYPES) == names


def test_struct_sizes_match_header():
    from rr_b200 import _lib
    assert C.sizeof(_lib.Event) == 24 and C.sizeof(_lib.Decision) == 16
    assert C.sizeof(_lib.DeploymentDesc) == 24 and C.sizeof(_lib.RouterSettings) == 32
    assert C.sizeof(_lib.DeploymentState) == 48
    assert C.sizeof(_lib.ModelDesc) == 36 and C.sizeof(_lib.Completion) == 48


def test_mt19937_seeding_matches_cpython():
    from rr_b200 import _lib
    for seed in [0, 1, 42, 1234, 2**31 - 1, 2**32 - 1, 2**32, 2**40 + 12345, 2**64 - 1]:
        buf = (C.c_uint32 * 625)()
        assert _lib.lib.rr_mt_seed_state(seed, buf) == 0
        want = random.Random(seed).getstate()[1]
        assert tuple(buf) == tuple(want), seed


def test_tokenizer_counts():
    from rr_b200 import _lib
    n = C.c_int32()
    s = "What is machine learning?".encode()
    assert _lib.lib.rr_count_tokens(s, len(s), C.byref(n)) == 0 and n.value == len(s) + 1
    ids = (C.c_int32 * 64)()
    assert _lib.lib.rr_tokenize(s, len(s), 128256, ids, 64, C.byref(n)) == 0
    assert n.value == len(s) + 1 and ids[0] == 1 and ids[1] == 3 + ord("W")
    assert _lib.lib.rr_tokenize(s, len(s), 128256, ids, 4, C.byref(n)) != 0   # buffer too small


def test_version_and_errors():
    from rr_b200 import _lib
    assert b"sm_100a" in _lib.lib.rr_version()
    assert b"429" in _lib.lib.rr_strerror(1)
