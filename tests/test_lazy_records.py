"""compat/lazy.py: the column-backed stand-in for the reference's list-of-dict stage outputs."""
import io
import os
import pickle

import numpy as np
import pytest

from radar_slam_b200.compat.lazy import LazyRecords, column_of, records_of


def _sample(n=7):
    rs = np.random.RandomState(3)
    cols = {'antenna': rs.randint(0, 8, n), 'range_bin': rs.randint(0, 256, n), 'range_m': rs.rand(n) * 30,
            'spatial_signature': rs.randn(n, 4) + 1j * rs.randn(n, 4), 'spectrum': None}
    dicts = [{k: (None if v is None else v[i]) for k, v in cols.items()} for i in range(n)]
    return cols, dicts


def test_behaves_like_the_list_of_dicts():
    cols, dicts = _sample()
    rec = LazyRecords(cols)
    assert len(rec) == len(dicts) and bool(rec) and rec == dicts and not (rec == dicts[:-1])
    assert list(rec[2].keys()) == list(dicts[2].keys())
    assert rec[-1]['range_m'] == dicts[-1]['range_m'] and rec[0]['spectrum'] is None
    assert np.array_equal(rec[3]['spatial_signature'], dicts[3]['spatial_signature'])
    assert [r['antenna'] for r in rec] == [d['antenna'] for d in dicts]
    assert rec[1:4] == dicts[1:4] and isinstance(rec[1:4], LazyRecords)
    assert rec[np.array([4, 0])] == [dicts[4], dicts[0]]
    with pytest.raises(IndexError):
        rec[len(dicts)]
    assert not LazyRecords({'a': np.zeros(0)}) and len(LazyRecords({})) == 0
    with pytest.raises(ValueError):
        LazyRecords({'a': np.zeros(2), 'b': np.zeros(3)})
    assert sorted(rec, key=lambda r: r['range_m'])[0]['range_m'] == cols['range_m'].min()


def test_column_access_and_fallbacks():
    cols, dicts = _sample()
    rec = LazyRecords(cols)
    assert column_of(rec, 'range_bin') is cols['range_bin']                     # no copy, no dict
    assert np.array_equal(column_of(dicts, 'range_bin', np.int64), cols['range_bin'])
    assert np.array_equal(column_of(np.array(dicts, dtype=object).tolist(), 'range_m', float), cols['range_m'])
    assert records_of(rec) is rec and rec == records_of(dicts)
    assert rec == records_of(np.array(dicts, dtype=object))                    # what np.load gives for the reference's files


@pytest.mark.parametrize("eager", [False, True])
def test_survives_the_pipeline_scripts_npz_round_trip(eager, monkeypatch):
    """run_ego_motion_pipeline.py:164-169, 211-212: np.savez(file, **peak_info); dict(np.load(file, allow_pickle=True))."""
    cols, dicts = _sample(50)
    rec = LazyRecords(cols)
    if eager:
        monkeypatch.setenv("RADAR_SLAM_B200_EAGER_FILES", "1")
    buf = io.BytesIO()
    np.savez(buf, peaks=rec, range_bins_m=np.arange(4.0))
    size = buf.tell()
    buf.seek(0)
    back = dict(np.load(buf, allow_pickle=True))
    got = back['peaks']
    assert got.dtype == object and got.ndim == (1 if eager else 0)
    assert rec == records_of(got)
    if not eager:
        assert isinstance(records_of(got), LazyRecords)
        assert size < len(pickle.dumps(np.array(dicts, dtype=object)))           # columns, not 50 pickled dicts
    assert pickle.loads(pickle.dumps(rec)) == dicts
