"""Two ranks on two B200s: the multi-GPU path of SURVEY.md 8(e) on hardware.  Each rank scans its contiguous block of
frames through the C ABI on its own GPU, the detection records are exchanged with ONE NCCL all-gather
(shard.RecordGather) and every rank must end with the job-wide table of the single-process CPU oracle, bit for bit.
Skipped unless two CUDA devices are visible (the 1-GPU box the driver uses for `-m gpu` skips it; run it with
`gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

from jda_b200 import api, shard, synth
from tests.conftest import SHIPPED_F32

pytestmark = pytest.mark.gpu

N_FRAMES = 7


def _frames():
    fr = [synth.face_canvas(), synth.facemix_frame(5), synth.facemix_frame(7), synth.blur_frame(1),
          synth.face_canvas()[:, ::-1].copy(), synth.facemix_frame(9), synth.noise_frame(3)]
    return np.stack(fr)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    frames = _frames()
    lo, hi = shard.shard_range(len(frames), rank, world)
    c = api.Cascador(SHIPPED_F32, double=False, device=rank)
    res = c.detect_batch(frames[lo:hi], th=-0.5, flat=True)
    assert c.last_stats["scan_launches"] >= 1
    rec = shard.pack_records_flat(*res, frame0=lo)
    g = shard.RecordGather(rec.shape[1], device="cuda", cap=2)        # cap 2: the first block overflows, the repeat path runs
    g.start(rec)
    table = g.finish()
    assert g.exchanges == 2
    g.start(rec)                                                     # steady state: one collective
    table2 = g.finish()
    assert g.exchanges == 3
    np.testing.assert_array_equal(table.view(np.uint32), table2.view(np.uint32))
    t3 = shard.all_gather_records(rec, device="cuda")                # the two-collective form agrees
    np.testing.assert_array_equal(table.view(np.uint32), t3.view(np.uint32))
    # mining records (truncated cascade, every survivor) take the same route
    raw = c.detect_batch(frames[lo:hi], t_limit=1, k_limit=0, flags=api.RAW_HITS | api.NO_FINAL_TH, flat=True)
    g.start(shard.pack_records_flat(*raw, frame0=lo))
    np.save(os.path.join(out_dir, "mine_%d.npy" % rank), g.finish())
    np.save(os.path.join(out_dir, "table_%d.npy" % rank), table)
    c.close()
    dist.destroy_process_group()


@pytest.mark.skipif(api.device_count() < 2, reason="needs two CUDA devices (gpurun --gpus 2)")
def test_two_gpus_end_with_the_oracle_table(oracle, oracle_shipped, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    frames = _frames()
    want = shard.pack_records([oracle.detect(oracle_shipped, f, th=-0.5) for f in frames])
    mine = []
    for f in frames:
        ob, osc, osh, _ = oracle.detect_raw(oracle_shipped, f, t_limit=1, use_th=False)
        mine.append((ob, osc, osh))
    want_mine = shard.pack_records(mine)
    assert len(want) >= 6 and len(want_mine) > len(want)
    for r in range(2):
        got = np.load(tmp_path / ("table_%d.npy" % r))
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
        got = np.load(tmp_path / ("mine_%d.npy" % r))
        np.testing.assert_array_equal(got.view(np.uint32), want_mine.view(np.uint32))
