"""GPU: K2, the prompt token count / tokenizer kernel (csrc/rr_tokenizer.cu), against the host form of the same tokenizer
(rr_tokenize) and its definition (count = UTF-8 bytes + 1 BOS; id = 3 + byte, folded into small vocabularies): bit-exact
ids, counts and offsets on ragged batches -- empty messages, lengths around the 16-byte vector and the 4 KB chunk edges,
multi-byte UTF-8, one long message.  (No reference tokenizer is reachable here: SURVEY.md 8c; the count feeds the tpm
check, reference config/config.yaml:42.)"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_tokenizer_matches_host_tokenizer_on_ragged_batches():
    from rr_b200.router import count_tokens, tokenize, tokenize_batch, tokenize_host
    rng = np.random.RandomState(0)
    lens = [0, 1, 2, 15, 16, 17, 31, 32, 33, 255, 256, 4095, 4096, 4097, 8191, 8192, 8193, 12345, 0, 100003]
    texts = ["".join(chr(int(c)) for c in rng.randint(32, 127, size=n)) for n in lens]
    texts += ["héllo wörld — ✓ 你好 🙂", "", "a" * 5000 + "é" * 777, "user: What is machine learning?"]
    for vocab in (128256, 32000, 300, 259, 128, 7):
        ids, counts = tokenize_batch(texts, vocab)
        assert len(ids) == len(texts)
        for t, got, c in zip(texts, ids, counts):
            want = tokenize_host(t, vocab)
            nb = len(t.encode("utf-8"))
            assert c == nb + 1 == len(got)
            assert np.array_equal(got, want), (vocab, nb)
            assert got[0] == 1 and (nb == 0 or (got[1:] >= 3).all()) and (got < max(vocab, 4)).all()
    # single-message forms
    assert count_tokens("What is machine learning?") == len("What is machine learning?") + 1
    assert np.array_equal(tokenize("abc", 128256), np.array([1, 3 + 97, 3 + 98, 3 + 99], dtype=np.int32))


def test_device_resident_form_is_asynchronous_on_the_callers_stream():
    import torch
    from rr_b200 import _lib
    texts = [b"alpha", b"", b"x" * 9000, b"omega!"]
    blob = b"".join(texts)
    off = np.zeros(len(texts) + 1, dtype=np.int64); np.cumsum([len(t) for t in texts], out=off[1:])
    d_text = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    d_off = torch.from_numpy(off).cuda()
    d_counts = torch.zeros(len(texts), dtype=torch.int32, device="cuda")
    d_ids = torch.zeros(len(blob) + len(texts), dtype=torch.int32, device="cuda")
    d_ids_off = torch.zeros(len(texts) + 1, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        _lib.check(_lib.lib.rr_tokenize_batch_device(C.c_void_p(d_text.data_ptr()), C.c_void_p(d_off.data_ptr()), len(texts), 9000,
                                                     128256, C.c_void_p(d_counts.data_ptr()), C.c_void_p(d_ids.data_ptr()),
                                                     C.c_void_p(d_ids_off.data_ptr()), C.c_void_p(st.cuda_stream)))
    st.synchronize()
    assert d_counts.tolist() == [len(t) + 1 for t in texts]
    io = d_ids_off.tolist()
    assert io == [int(off[i]) + i for i in range(len(texts) + 1)]
    ids = d_ids.cpu().numpy()
    for i, t in enumerate(texts):
        assert ids[io[i]] == 1 and np.array_equal(ids[io[i] + 1: io[i + 1]], 3 + np.frombuffer(t, dtype=np.uint8).astype(np.int32))
