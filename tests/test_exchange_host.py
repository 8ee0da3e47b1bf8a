"""Host side of the hit exchange: the log format (include/anchored_fusion.h) and its parser."""
import numpy as np
import pytest

from anchored_fusion_b200 import dist as afdist
from anchored_fusion_b200._lib import HIT_DTYPE


def _marker(base, count):
    m = np.zeros(1, HIT_DTYPE)
    raw = m.view(np.uint32).reshape(-1, 4)
    raw[0] = [afdist.LOG_MARKER, base & 0xFFFFFFFF, base >> 32, count]
    return m


def _hits(ids):
    h = np.zeros(len(ids), HIT_DTYPE)
    h["read_id"] = ids
    h["pos"] = np.arange(len(ids)) + 1
    h["m_len"] = 100
    h["score_strand"] = 200
    return h


def test_parse_log_splits_batches_and_keeps_64_bit_bases():
    big = (1 << 33) + 12345 * 32
    log = np.concatenate([_marker(0, 3), _hits([0, 5, 9]), _marker(64, 0), _marker(big, 2), _hits([1, 2])])
    got = afdist.parse_log(log)
    assert [(b, len(h)) for b, h in got] == [(0, 3), (64, 0), (big, 2)]
    assert got[0][1]["read_id"].tolist() == [0, 5, 9]
    merged = afdist.globalise(got)
    assert merged["read_id"].tolist() == [0, 5, 9, 2 * big + 1, 2 * big + 2]
    assert merged.dtype["read_id"] == np.int64 and merged["m_len"].tolist() == [100] * 5
    assert afdist.parse_log(np.zeros(0, HIT_DTYPE)) == []
    assert len(afdist.globalise([])) == 0


def test_parse_log_rejects_broken_logs():
    with pytest.raises(ValueError):
        afdist.parse_log(_hits([3]))                                         # no marker first
    with pytest.raises(ValueError):
        afdist.parse_log(np.concatenate([_marker(0, 5), _hits([1, 2])]))     # truncated batch
