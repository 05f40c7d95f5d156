"""CPU: the numpy Philox twin against the published Random123 known-answer vectors for philox4x32-10."""
import numpy as np

from oracle import philox


def _run(counter, key):
    out = philox.philox4x32_10(np.array([counter], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
    return [int(x) for x in out]


def test_random123_known_answers():
    assert _run([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _run([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_streams_are_keyed_by_env_step_and_stream():
    ids = np.arange(16)
    m0, n0 = philox.reset_tables(5, 3, ids)
    m1, n1 = philox.reset_tables(5, 4, ids)
    assert m0.dtype == np.float32 and n0.shape == (16, 21)
    assert (m0 != m1).any() and (n0 != n1).any()
    # a draw depends on the GLOBAL env id only: any subset / shard sees the same numbers
    sub = np.array([3, 9, 11])
    ms, ns = philox.reset_tables(5, 3, sub)
    assert np.array_equal(ms, m0[sub]) and np.array_equal(ns, n0[sub])
    s = philox.stone_tables(5, 3, ids)
    assert s.shape == (5, 16, 20) and float(s.min()) >= 0.0 and float(s.max()) < 1.0
    assert not np.array_equal(s[0, :, :20].ravel()[:21], np.concatenate([[m0[0]], n0[0]])[:21])  # other stream
    u = philox.stream_uniforms(1, 0, 0, np.arange(4096), 22)
    assert 0.45 < float(u.mean()) < 0.55
