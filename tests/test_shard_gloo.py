"""world_size-2 gloo test of the multi-GPU plumbing (jda_b200/shard.py) on CPU: contiguous frame
sharding + the single all-gather of detection records.  Per-rank detections come from the oracle
(the CUDA path needs a GPU; the exchange logic under test is backend-agnostic)."""
import os
import socket

import numpy as np
import pytest

from jda_b200 import shard, synth
from tests.conftest import SHIPPED_F32


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 2845, 100000):
        for world in (1, 2, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    res = []
    for n in (0, 3, 1):
        res.append((rng.integers(0, 2000, (n, 3)).astype(np.int32), rng.normal(size=n).astype(np.float32),
                    rng.normal(size=(n, 54)).astype(np.float32)))
    rec = shard.pack_records(res, frame0=2840)
    fid, boxes, scores, shapes = shard.unpack_records(rec)
    assert fid.tolist() == [2841, 2841, 2841, 2842]
    np.testing.assert_array_equal(boxes, np.concatenate([r[0] for r in res]))
    np.testing.assert_array_equal(scores, np.concatenate([r[1] for r in res]))
    np.testing.assert_array_equal(shapes, np.concatenate([r[2] for r in res]))
    # the flat batch result (jdaB200DetectBatchFlat) packs to the same records
    flat = shard.pack_records_flat(np.array([len(r[1]) for r in res]), boxes, scores, shapes, frame0=2840)
    np.testing.assert_array_equal(flat.view(np.uint32), rec.view(np.uint32))


def _frames():
    # 5 frames: rank 0 gets 3 (ceil split), rank 1 gets 2; frame 3 has no face
    return [synth.face_canvas(), synth.facemix_frame(5), synth.facemix_frame(7), synth.blur_frame(1),
            synth.face_canvas()[:, ::-1].copy()]


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = _frames()
    lo, hi = shard.shard_range(len(frames), rank, world)
    o = pyoracle.Oracle()
    h = o.load(SHIPPED_F32, double=False)
    local = [o.detect(h, frames[i], th=-0.5) for i in range(lo, hi)]
    rec = shard.pack_records(local, frame0=lo)
    table = shard.all_gather_records(rec)
    np.save(os.path.join(out_dir, "table_%d.npy" % rank), table)
    # the one-collective exchange: same table; a block that is too small is repeated with a larger one
    g = shard.RecordGather(rec.shape[1], cap=256)
    g.start(rec)
    t2 = g.finish()
    assert g.exchanges == 1 and g.finish() is None
    np.testing.assert_array_equal(t2.view(np.uint32), table.view(np.uint32))
    small = shard.RecordGather(rec.shape[1], cap=1)
    small.start(rec)
    t3 = small.finish()
    assert small.exchanges == 2 and small.cap >= 2
    np.testing.assert_array_equal(t3.view(np.uint32), table.view(np.uint32))
    small.start(rec[:0])                       # an empty contribution from every rank
    assert small.finish().shape == (0, rec.shape[1])
    o.release(h)
    dist.destroy_process_group()


def test_all_gather_detections_world2(oracle, oracle_shipped, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0 = np.load(tmp_path / "table_0.npy")
    t1 = np.load(tmp_path / "table_1.npy")
    np.testing.assert_array_equal(t0.view(np.uint32), t1.view(np.uint32))  # every rank holds the same table
    frames = _frames()
    want = shard.pack_records([oracle.detect(oracle_shipped, f, th=-0.5) for f in frames])
    np.testing.assert_array_equal(t0.view(np.uint32), want.view(np.uint32))  # = the single-process answer
    fid = shard.unpack_records(t0)[0]
    assert (np.diff(fid) >= 0).all() and len(fid) >= 4


def test_record_gather_block_capacity_rule():
    """detections (hundreds per batch): twice the largest count, a power of two; mining (tens of thousands per batch):
    a quarter above it, rounded to 4096 -- every padded row crosses NVLink and PCIe to every rank"""
    from jda_b200.shard import RecordGather
    assert RecordGather._next_cap(0) == 1 and RecordGather._next_cap(3) == 8 and RecordGather._next_cap(200) == 512
    assert RecordGather._next_cap(4095) == 8192
    for n in (4096, 27785, 100000):
        cap = RecordGather._next_cap(n)
        assert cap % 4096 == 0 and 1.25 * n <= cap < 1.25 * n + 4096
